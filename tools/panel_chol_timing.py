"""Per-phase cycle counters of the Cholesky-QR panel's algebra kernel (thread 0) from a -DSVDB_PANEL_TIMING=1 side build, plus
per-kernel times of the three launches:   python tools/panel_chol_timing.py <f32|f64> <band> <m> [m ...]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VAR = os.path.join(ROOT, "svdsolver_b200", "libsvdb200_timing.so")
if "SVDB200_LIB" not in os.environ:
    from svdsolver_b200 import build as B
    B.build()
    src = os.path.join(ROOT, "svdsolver_b200", "csrc", "stage1_panel_chol.cu")
    if not os.path.exists(VAR) or os.path.getmtime(VAR) < os.path.getmtime(src):
        B.build(out=VAR, extra=("-DSVDB_PANEL_TIMING=1",), only=("stage1_panel_chol.cu",))
    os.environ["SVDB200_LIB"] = VAR
    os.execv(sys.executable, [sys.executable] + sys.argv)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from svdsolver_b200 import capi  # noqa: E402

suf, b = sys.argv[1], int(sys.argv[2])
dt = np.float32 if suf == "f32" else np.float64
tdt = torch.float32 if suf == "f32" else torch.float64
names = ["sum partials + load top block", "G = G2 + A1^T A1", "elimination loop (Cholesky + LU)", "M1 = U~^-1", "X = G2 M1", "Y^T Y", "T = (T^-1)^-1",
         "M2, V2 top, stores"]
for m in [int(x) for x in sys.argv[3:]]:
    with capi.Handle(m, b, dt) as h:
        a0 = torch.rand(m, b, device="cuda", dtype=tdt) * 5
        v = torch.empty(m, b, device="cuda", dtype=tdt)
        v2 = torch.empty(m * b, device="cuda", dtype=tdt)
        out = (ctypes.c_longlong * 16)()
        reps = 5
        for rep in range(reps + 1):
            a = a0.clone()
            h.synchronize()
            if rep == 1:
                capi.lib().svdb200_debug_panel_chol_timing(out)
            h.panel_factor_dev(a.data_ptr(), b, m, b, 0, v.data_ptr(), v2.data_ptr())
        h.synchronize()
        capi.lib().svdb200_debug_panel_chol_timing(out)
        cnt = max(out[15], 1)
        tot = sum(out[i] for i in range(8))
        print(f"{suf} band {b} m={m}: algebra kernel {tot / cnt:.0f} cycles per panel ({tot / cnt / 1.965e3:.1f} us at 1965 MHz), {cnt} panels")
        for i, nm in enumerate(names):
            print(f"    {nm:36s} {100.0 * out[i] / tot:5.1f}%  {out[i] / cnt:9.0f} cycles")
