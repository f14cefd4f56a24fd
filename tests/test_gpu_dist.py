"""Multi-GPU parity under pytest (-m gpu): spawns tests/dist_check.py with torchrun on all visible GPUs (2, 4 or 8);
skipped on a single-GPU box.  The single-rank path of the same driver is covered in test_gpu_parity.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("which", ["small", "large"])
def test_dist_stage1_under_torchrun(which):
    n = _ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs on one box")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dist_check.py"), which]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    sys.stdout.write(out.stdout[-4000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "FAIL" not in out.stdout


def test_dist_band_gather_and_svdvals_single_rank():
    """the band hand-off and the distributed svdvals entry on ONE rank (no NCCL): same code path minus the collectives"""
    import ctypes
    import torch
    from svdsolver_b200 import capi, distributed as D
    from svdsolver_b200.synth import uniform_matrix
    for n, b, dt, tdt in ((768, 32, np.float64, torch.float64), (512, 64, np.float32, torch.float32)):
        a = uniform_matrix(n, n, 586 + n, 0.0, 5.0, dt)
        loc = torch.from_numpy(a.copy()).cuda()
        packed = torch.zeros(n, b + 1, dtype=tdt, device="cuda")
        with D.DistHandle(n, b, dt, 0, 1, (ctypes.c_ubyte * 128)()) as dh:
            dh.dense_to_band_dev(loc.data_ptr())
            dh.gather_band_dev(loc.data_ptr(), packed.data_ptr())
            torch.cuda.synchronize()
        full = loc.cpu().numpy()
        assert np.array_equal(D.unpack_band(packed.cpu().numpy(), n, b), np.triu(np.tril(full, b)))
        loc2 = torch.from_numpy(a.copy()).cuda()
        sigma = torch.zeros(n, dtype=tdt, device="cuda")
        with D.DistHandle(n, b, dt, 0, 1, (ctypes.c_ubyte * 128)()) as dh:
            dh.configure(stage2_schedule=1, qr_method=2)
            dh.svdvals_dev(loc2.data_ptr(), sigma.data_ptr())
            torch.cuda.synchronize()
        s_ref = np.linalg.svd(a.astype(np.float64), compute_uv=False)
        assert np.abs(sigma.cpu().numpy().astype(np.float64) - s_ref).max() <= (2e-5 if dt == np.float32 else 1e-11) * s_ref[0]
