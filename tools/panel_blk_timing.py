"""Per-phase cycle counters of the blocked panel kernel (CTA 0, thread 0) from a -DSVDB_PANEL_TIMING=1 side build:
    python tools/panel_blk_timing.py <n> <band> <f32|f64> [n band dtype ...]
builds svdsolver_b200/libsvdb200_timing.so on first use (only stage1_panel_blk.cu is recompiled)."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VAR = os.path.join(ROOT, "svdsolver_b200", "libsvdb200_timing.so")
if "SVDB200_LIB" not in os.environ:
    from svdsolver_b200 import build as B
    B.build()
    if not os.path.exists(VAR) or os.path.getmtime(VAR) < os.path.getmtime(os.path.join(ROOT, "svdsolver_b200", "csrc", "stage1_panel_blk.cu")):
        B.build(out=VAR, extra=("-DSVDB_PANEL_TIMING=1",), only=("stage1_panel_blk.cu",))
    os.environ["SVDB200_LIB"] = VAR
    os.execv(sys.executable, [sys.executable] + sys.argv)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from svdsolver_b200 import capi  # noqa: E402

args = sys.argv[1:]
for k in range(0, len(args), 3):
    n, b, suf = int(args[k]), int(args[k + 1]), args[k + 2]
    dt = np.float32 if suf == "f32" else np.float64
    tdt = torch.float32 if suf == "f32" else torch.float64
    with capi.Handle(n, b, dt) as h:
        a = torch.empty(n, n, device="cuda", dtype=tdt)
        out = (ctypes.c_longlong * 16)()
        for rep in range(2):
            h.fill_uniform_dev(a.data_ptr(), n * n, 586 + n, 0.0, 5.0)
            h.synchronize()
            capi.lib().svdb200_debug_panel_blk_timing(out)          # clear
            h.reset_profile(); h.set_profile(True)                  # serialised: one kernel at a time
            h.dense_to_band_dev(a.data_ptr(), n, b)
            h.synchronize(); h.set_profile(False)
        p = h.get_profile()["panel"]
        capi.lib().svdb200_debug_panel_blk_timing(out)
        rounds, panels = max(out[8], 1), max(out[9], 1)
        names = ["load + first dot products", "publish psum + sync", "all-reduce", "algebra (8 Householder steps)", "pass (update + next dots)", "epilogue: V2 stores"]
        tot = sum(out[i] for i in range(6)) + sum(out[i] for i in (10, 11, 12))
        print(f"n={n} band={b} {suf}: {panels} panels, {rounds} rounds ({rounds / panels:.2f} per panel; band/8 = {b // 8}), "
              f"panel class {p['ms']:.2f} ms = {p['ms'] / panels * 1e3:.1f} us per panel; CTA-0 cycles per panel {tot / panels:.0f}")
        for i, nm in enumerate(names):
            per = out[i] / (panels if i in (0, 5) else rounds)
            print(f"    {nm:32s} {100.0 * out[i] / tot:5.1f}%   {per:9.0f} cycles per {'panel' if i in (0, 5) else 'round'}")
        for i, nm in ((10, "epilogue: last T block + cluster sync"), (11, "epilogue: stage rows, V / R stores"), (12, "epilogue: V2 = V S^T")):
            print(f"    {nm:40s} {100.0 * out[i] / tot:5.1f}%   {out[i] / panels:9.0f} cycles per panel")
    del a
    torch.cuda.empty_cache()
