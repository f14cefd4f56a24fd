"""Very tall panels: stage 1 with the register / cluster panel kernels vs the shared-memory L2-transport kernels
(SVDB200_PANEL_REG=0) on the same input, plus timing.   python tools/tall_check.py <n> <band> <f32|f64>"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svdsolver_b200 import capi  # noqa: E402

n, b = int(sys.argv[1]), int(sys.argv[2])
dt = np.float32 if sys.argv[3] == "f32" else np.float64
tdt = torch.float32 if dt == np.float32 else torch.float64
outs = []
for reg in ("1", "0"):
    os.environ["SVDB200_PANEL_REG"] = reg
    h = capi.Handle(n, b, dt)
    s = torch.cuda.Stream()
    h.set_stream(s.cuda_stream)
    a = torch.empty(n, n, device="cuda", dtype=tdt)
    best = None
    for rep in range(2):
        h.fill_uniform_dev(a.data_ptr(), n * n, 586 + n, 0.0, 5.0)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        h.dense_to_band_dev(a.data_ptr(), n, b)
        e1.record(s)
        s.synchronize()
        t = e0.elapsed_time(e1)
        best = t if best is None else min(best, t)
    print(f"panel_reg={reg}: stage1 {best:.1f} ms", flush=True)
    # keep only the band (diagonals 0..b) as a compact array
    band = torch.stack([torch.diagonal(a, k)[: n - b] for k in range(b + 1)])
    outs.append(band.double().cpu())
    h.close()
    del a
    torch.cuda.empty_cache()
d = (outs[0].abs() - outs[1].abs()).abs().max().item() / outs[1].abs().max().item()
print(f"max | |band_reg| - |band_smem| | / max|band| = {d:.3e}", flush=True)
sys.exit(0 if d < (1e-3 if dt == np.float32 else 1e-9) else 1)
