"""Host-side plumbing of the multi-GPU stage 1 (one process per GPU, torch.distributed for the
rendezvous only): 1-D block-cyclic column layout helpers and the NCCL unique-id exchange.

Layout (include/svdb200.h): global block column j (width = band) lives on rank j % P as local block
j // P; a rank stores its blocks side by side, row-major n x ncols_local.
"""
import ctypes

import numpy as np

from . import capi


def local_cols(n, band, rank, nranks):
    return int(capi.lib().svdb200_dist_local_cols(ctypes.c_size_t(n), ctypes.c_size_t(band), ctypes.c_int(rank), ctypes.c_int(nranks)))


def owned_blocks(n, band, rank, nranks):
    return list(range(rank, n // band, nranks))


def scatter_block_cyclic(a, band, rank, nranks):
    """Columns of the global matrix `a` owned by `rank`, as a contiguous (n, ncols_local) array."""
    n = a.shape[1]
    blocks = owned_blocks(n, band, rank, nranks)
    if not blocks:
        return np.zeros((a.shape[0], 0), a.dtype)
    return np.ascontiguousarray(np.concatenate([a[:, j * band:(j + 1) * band] for j in blocks], axis=1))


def gather_block_cyclic(parts, band, n):
    """Inverse of scatter_block_cyclic: parts[r] is rank r's (n, ncols_local) array."""
    nranks = len(parts)
    out = np.zeros((parts[0].shape[0], n), parts[0].dtype)
    for r, p in enumerate(parts):
        for lb, j in enumerate(owned_blocks(n, band, r, nranks)):
            out[:, j * band:(j + 1) * band] = p[:, lb * band:(lb + 1) * band]
    return out


def unpack_band(packed, n, band):
    """packed band storage (n x (band+1): packed[gc, t] = A[gc - band + t, gc]) -> dense n x n numpy array"""
    out = np.zeros((n, n), packed.dtype)
    for t in range(band + 1):
        k = band - t                      # super-diagonal index
        rows = np.arange(0, n - k)
        out[rows, rows + k] = packed[k:, t]
    return out


def exchange_unique_id(rank, nranks):
    """Rank 0 creates the ncclUniqueId (through the C ABI) and broadcasts it with torch.distributed."""
    import torch.distributed as dist
    buf = (ctypes.c_ubyte * 128)()
    if nranks > 1:
        if rank == 0:
            st = capi.lib().svdb200_dist_unique_id(buf)
            if st != 0:
                raise capi.SvdB200Error(st, "svdb200_dist_unique_id")
        obj = [bytes(buf)]
        dist.broadcast_object_list(obj, src=0)
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(obj[0])
    return buf


class DistHandle:
    """svdb200_dist_handle: stage 1 of an n x n matrix distributed over `nranks` GPUs."""

    def __init__(self, n, band, dtype, rank, nranks, unique_id, device=0):
        self.suf, code = capi._suf(dtype)
        self.h = ctypes.c_void_p()
        st = capi.lib().svdb200_dist_create(ctypes.byref(self.h), ctypes.c_int(device), ctypes.c_int(rank), ctypes.c_int(nranks),
                                            unique_id, ctypes.c_size_t(n), ctypes.c_size_t(band), ctypes.c_int(code))
        if st != 0:
            self.h = ctypes.c_void_p()
            raise capi.SvdB200Error(st, capi.lib().svdb200_strerror(st).decode())
        self.n, self.band = n, band

    def set_stream(self, ptr):
        st = capi.lib().svdb200_dist_set_stream(self.h, ctypes.c_void_p(ptr))
        if st != 0:
            raise capi.SvdB200Error(st, "set_stream")

    def dense_to_band_dev(self, a_local_ptr):
        st = getattr(capi.lib(), f"svdb200_dist_dense_to_band_dev_{self.suf}")(self.h, ctypes.c_void_p(a_local_ptr), ctypes.c_size_t(self.n),
                                                                                ctypes.c_size_t(self.band))
        if st != 0:
            raise capi.SvdB200Error(st, capi.lib().svdb200_strerror(st).decode())

    def gather_band_dev(self, a_local_ptr, packed_ptr):
        st = getattr(capi.lib(), f"svdb200_dist_gather_band_dev_{self.suf}")(self.h, ctypes.c_void_p(a_local_ptr), ctypes.c_void_p(packed_ptr))
        if st != 0:
            raise capi.SvdB200Error(st, capi.lib().svdb200_strerror(st).decode())

    def svdvals_dev(self, a_local_ptr, sigma_ptr):
        """stage 1 on all ranks, band gathered, stage 2 + singular values on rank 0 (sigma_ptr: device, n; rank 0 only)"""
        st = getattr(capi.lib(), f"svdb200_dist_svdvals_dev_{self.suf}")(self.h, ctypes.c_void_p(a_local_ptr), ctypes.c_void_p(sigma_ptr or 0))
        if st != 0:
            raise capi.SvdB200Error(st, capi.lib().svdb200_strerror(st).decode())

    def configure(self, stage2_schedule=-1, qr_method=-1, tc05_mode=-1):
        st = capi.lib().svdb200_dist_configure(self.h, ctypes.c_int(stage2_schedule), ctypes.c_int(qr_method), ctypes.c_int(tc05_mode))
        if st != 0:
            raise capi.SvdB200Error(st, "svdb200_dist_configure")

    def configure_panels(self, lq_distributed=1):
        """1 (default): LQ panels by local Gram matrix + one all-reduce; 0: all-gather of the row panel, redundant factorisation"""
        st = capi.lib().svdb200_dist_configure_panels(self.h, ctypes.c_int(lq_distributed))
        if st != 0:
            raise capi.SvdB200Error(st, "svdb200_dist_configure_panels")

    def launch_count(self):
        return int(capi.lib().svdb200_dist_launch_count(self.h))

    def close(self):
        if self.h:
            capi.lib().svdb200_dist_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
