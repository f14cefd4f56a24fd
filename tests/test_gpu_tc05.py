"""FP32 trailing update on tcgen05 / TMEM / TMA (svdsolver_b200/csrc/gemm_tc05.cu), forced on through
svdb200_set_tc05(h, 2): building-block self-test and the three GEMM shapes of qr_apply / lq_apply
(svd_parallel.h:243-281) against an fp64 reference, including ragged edges and split-K shapes."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from svdsolver_b200 import capi as m
    m.lib()
    return m


def handle(capi, n, band, suf):
    return capi.Handle(n, band, np.float32 if suf == "f32" else np.float64)


@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
def test_tc05_selftest_operand_layouts(capi, a_mn, b_mn):
    """TMA box -> swizzled smem -> UMMA descriptor -> TMEM -> tcgen05.ld, exact on TF32-representable inputs."""
    import torch
    rng = np.random.default_rng(3)
    A = (rng.integers(-8, 9, size=(128, 32)) / 8.0).astype(np.float32)
    B = (rng.integers(-8, 9, size=(32, 64)) / 8.0).astype(np.float32)
    a_host = A if a_mn == 0 else np.ascontiguousarray(A.T)
    b_host = np.ascontiguousarray(B.T) if b_mn == 0 else B
    a = torch.from_numpy(a_host.copy()).cuda()
    b = torch.from_numpy(b_host.copy()).cuda()
    out = torch.zeros(128 * 64 + 1, device="cuda")
    dump = torch.zeros(6144, device="cuda")
    torch.cuda.synchronize()
    with handle(capi, 256, 64, "f32") as h:
        st = capi.lib().svdb200_tc05_selftest(h.h, ctypes.c_int(a_mn), ctypes.c_int(b_mn), ctypes.c_void_p(a.data_ptr()),
                                              ctypes.c_void_p(b.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                              ctypes.c_void_p(dump.data_ptr()))
        assert st == 0
    D = out.cpu().numpy()[:128 * 64].reshape(128, 64)
    assert np.array_equal(D, (A.astype(np.float64) @ B.astype(np.float64)).astype(np.float32))


@pytest.mark.parametrize("m,n,b", [(256, 192, 32), (300, 132, 64), (128, 128, 64), (1000, 776, 32), (2048, 1536, 64),
                                   (1500, 2048, 32), (4096, 4160, 64), (64, 3008, 64), (3008, 64, 64), (5000, 36, 32)])
def test_tc05_trailing_update_gemms(capi, m, n, b):
    import torch
    g = torch.Generator(device="cuda").manual_seed(5)
    ld = n + 8
    C = torch.rand(m, ld, device="cuda", generator=g) * 5
    V = torch.rand(m, b, device="cuda", generator=g) - 0.5
    Ut = torch.rand(n, b, device="cuda", generator=g) - 0.5
    Q = torch.rand(b, n, device="cuda", generator=g) - 0.5
    tol = 4e-6
    torch.cuda.synchronize()
    with handle(capi, max(m, n) + 64, b, "f32") as h:
        h.set_tc05(2)
        n0 = h.launch_count()
        W = torch.zeros(b, n, device="cuda")
        torch.cuda.synchronize()
        h.gemm_tn_dev(V.data_ptr(), C.data_ptr(), ld, m, n, b, W.data_ptr())
        h.synchronize()
        ref = V.double().T @ C[:, :n].double()
        assert (W.double() - ref).abs().max().item() <= tol * ref.abs().max().item()
        W2 = torch.zeros(m, b, device="cuda")
        torch.cuda.synchronize()
        h.gemm_nn_dev(C.data_ptr(), ld, m, n, b, Ut.data_ptr(), W2.data_ptr())
        h.synchronize()
        ref2 = C[:, :n].double() @ Ut.double()
        assert (W2.double() - ref2).abs().max().item() <= tol * ref2.abs().max().item()
        C2 = C.clone()
        torch.cuda.synchronize()
        h.rank_update_dev(C2.data_ptr(), ld, m, n, b, V.data_ptr(), Q.data_ptr(), n)
        h.synchronize()
        ref3 = C[:, :n].double() + V.double() @ Q.double()
        assert (C2[:, :n].double() - ref3).abs().max().item() <= tol * ref3.abs().max().item()
        assert torch.equal(C2[:, n:], C[:, n:])          # padding columns untouched
        assert h.launch_count() > n0


def test_tc05_matches_mma_sync_path_in_stage1(capi):
    """stage 1 (panel order) with the tcgen05 update == with the mma.sync update, to fp32 round-off."""
    from svdsolver_b200.synth import uniform_matrix
    n, b = 2048, 64
    a = uniform_matrix(n, n, 586 + n, 0.0, 5.0, np.float32)
    outs = []
    for mode in (0, 2):
        with handle(capi, n, b, "f32") as h:
            h.set_tc05(mode)
            outs.append(h.dense_to_band(a.copy(), b))
    band = [np.triu(np.tril(o, b)) for o in outs]
    scale = np.abs(band[0]).max()
    assert np.abs(np.abs(band[0]) - np.abs(band[1])).max() <= 1e-4 * scale
    s0 = np.linalg.svd(a.astype(np.float64), compute_uv=False)
    s1 = np.linalg.svd(band[1].astype(np.float64), compute_uv=False)
    assert np.abs(s0 - s1).max() <= 1e-4 * s0[0]
