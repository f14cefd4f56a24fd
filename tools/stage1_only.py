"""Stage 1 only (dense -> band) on a synthetic matrix: used under ncu to capture individual kernels.
    python tools/stage1_only.py <n> <band> <f32|f64>"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svdsolver_b200 import capi  # noqa: E402

n, b = int(sys.argv[1]), int(sys.argv[2])
dt = np.float32 if sys.argv[3] == "f32" else np.float64
tdt = torch.float32 if dt == np.float32 else torch.float64
h = capi.Handle(n, b, dt)
a = torch.empty(n, n, device="cuda", dtype=tdt)
h.fill_uniform_dev(a.data_ptr(), n * n, 586 + n, 0.0, 5.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s = torch.cuda.Stream()
h.set_stream(s.cuda_stream)
if len(sys.argv) > 4 and sys.argv[4] == 'serial':
    h.set_profile(True)        # brackets every launch with events: kernels run one at a time
reps = int(os.environ.get("REPS", "3"))       # first repetition carries one-time costs (module load, lazy workspaces)
times = []
for rep in range(reps):
    if rep:
        h.fill_uniform_dev(a.data_ptr(), n * n, 586 + n, 0.0, 5.0)
        s.synchronize()
    e0.record(s)
    h.dense_to_band_dev(a.data_ptr(), n, b)
    e1.record(s)
    s.synchronize()
    times.append(e0.elapsed_time(e1))
print("stage1 ms", min(times), [round(t, 1) for t in times], flush=True)
import ctypes
out = (ctypes.c_longlong * 16)()
capi.lib().svdb200_debug_panel_timing(out)
if any(out):
    tot = sum(out)
    names = ["loop top/fused pass tail", "psum+owner publish+sync", "xwarp reduce+cluster.sync", "level-1 DSMEM reduce", "level-2 publish", "level-2 poll+sync",
             "level-2 read+sync", "scalars+fused pass+sync", "after loop", "epilogue"]
    for i, nme in enumerate(names):
        print(f"  phase {i} {nme:28s} {out[i]:12d} cycles {100.0*out[i]/tot:5.1f}%")
