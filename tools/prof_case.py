"""One small invocation of the hot path for ncu: dense -> band -> bidiagonal at (n, band, dtype)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from svdsolver_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
band = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dt = {"f64": (np.float64, torch.float64), "f32": (np.float32, torch.float32)}[sys.argv[3] if len(sys.argv) > 3 else "f64"]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
stage = sys.argv[5] if len(sys.argv) > 5 else "both"
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
with capi.Handle(n, band, dt[0]) as h:
    h.set_stream(s.cuda_stream)
    a = torch.empty(n, n, device="cuda", dtype=dt[1])
    d = torch.empty(n, device="cuda", dtype=dt[1]); e = torch.empty(n, device="cuda", dtype=dt[1])
    for r in range(reps):
        h.fill_uniform_dev(a.data_ptr(), n * n, 586 + n, 0.0, 5.0)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        if stage in ("both", "s1"):
            h.dense_to_band_dev(a.data_ptr(), n, band)
        e1.record(s)
        if stage in ("both", "s2"):
            h.band_to_bidiag_dev(a.data_ptr(), n, band, d.data_ptr(), e.data_ptr())
        e2.record(s)
        torch.cuda.synchronize()
        print(f"rep {r}: stage1 {e0.elapsed_time(e1):.3f} ms  stage2 {e1.elapsed_time(e2):.3f} ms  launches {h.launch_count()}")
