"""Deterministic synthetic inputs (replaces the reference's unseeded generator).

The reference fills matrices with i.i.d. U[min,max) drawn from a fresh ``random_device``-seeded
``mt19937`` per element (matrix.h:350-363, matrix_gpu.h:336-349; benchmark range [0,5),
svd_cuda_2.cu:1361-1362) and is therefore not reproducible.  We keep the distribution and make the
stream reproducible and position-addressable: element ``i`` (row-major) of the matrix with seed
``s`` is ``min + (max-min) * (splitmix64(s + i) >> 11) * 2**-53`` evaluated in double and then
rounded to the element type.  The same formula is implemented in C++ (csrc/host_util.h) and in the
device fill kernel, so host, device and oracle all see identical inputs.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def uniform_matrix(n_rows: int, n_cols: int, seed: int, lo: float = 0.0, hi: float = 5.0, dtype=np.float64):
    """Row-major (n_rows, n_cols) matrix of U[lo,hi) values; ``seed`` selects the stream."""
    idx = np.arange(n_rows * n_cols, dtype=np.uint64)
    with np.errstate(over="ignore"):
        bits = splitmix64(idx + np.uint64(seed)) >> np.uint64(11)
    u = bits.astype(np.float64) * (1.0 / 9007199254740992.0)
    return (lo + (hi - lo) * u).astype(dtype).reshape(n_rows, n_cols)


def default_seed(n: int) -> int:
    """Benchmark convention: seed = 586 + n."""
    return 586 + n
