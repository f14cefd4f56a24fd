"""One stage-1 panel at a time (svdb200_panel_factor_dev_*): Cholesky-QR panel (kind 2) vs blocked kernel (1) vs per-column
kernels (0) -- time per panel, agreement of R and V with the per-column kernels, Q^T A = [R; 0] with the compact-WY factors.
python tools/panel_only.py <f32|f64> <band> <m> [m ...]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svdsolver_b200 import capi  # noqa: E402

suf, b = sys.argv[1], int(sys.argv[2])
dt = np.float32 if suf == "f32" else np.float64
tdt = torch.float32 if suf == "f32" else torch.float64
for m in [int(x) for x in sys.argv[3:]]:
    with capi.Handle(m, b, dt) as h:
        s = torch.cuda.Stream()
        h.set_stream(s.cuda_stream)
        g = torch.Generator(device="cuda").manual_seed(m)
        for trans in (0, 1):
            a0 = torch.rand((b, m) if trans else (m, b), device="cuda", dtype=tdt, generator=g) * 5
            v = torch.empty(m, b, device="cuda", dtype=tdt)
            v2 = torch.empty(m * b, device="cuda", dtype=tdt)
            res = {}
            for blocked in (2, 1, 0):
                assert capi.lib().svdb200_set_panel_kernel(h.h, ctypes.c_int(blocked)) == 0
                ts = []
                for rep in range(6):
                    a = a0.clone()
                    s.synchronize()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(s)
                    h.panel_factor_dev(a.data_ptr(), a.shape[1], m, b, trans, v.data_ptr(), v2.data_ptr())
                    e1.record(s)
                    s.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                R = (a.T if trans else a)[:b, :b].double().cpu().numpy()
                V = v.double()
                V2 = (v2.view(b, m).T if trans else v2.view(m, b)).double()
                # Q^T A0 = [R; 0] with Q = I + V S V^T, V2 = V S^T  =>  Q^T X = X + V2 (V^T X)
                X = (a0.T if trans else a0).double()
                QtX = X + V2 @ (V.T @ X)
                # orthogonality of Q = I + V S V^T:  Q Q^T - I = V (S + S^T + S (V^T V) S^T) V^T, with S^T = V[:b]^-1 V2[:b]
                St = torch.linalg.solve_triangular(V[:b], V2[:b], upper=False, unitriangular=True)
                orth = float((St + St.T + St.T @ (V.T @ V) @ St).abs().max())
                res[blocked] = (min(ts[1:]), R, float((QtX[:b] - torch.from_numpy(np.triu(R)).cuda()).abs().max() / X.abs().max()),
                                float(QtX[b:].abs().max() / X.abs().max()), V.cpu().numpy(), orth)
            fb = ctypes.c_longlong(0)
            capi.lib().svdb200_chol_fallback_count(h.h, ctypes.byref(fb))
            rdiff = float(np.abs(res[2][1] - res[0][1]).max() / np.abs(res[0][1]).max())
            vdiff = float(np.abs(res[2][4] - res[0][4]).max())
            print(f"{suf} band {b} m={m:6d} {'LQ' if trans else 'QR'}: chol {res[2][0]:7.1f} us  blocked {res[1][0]:7.1f} us  per-column {res[0][0]:7.1f} us | "
                  f"chol vs per-column R {rdiff:.1e} V {vdiff:.1e} fallbacks {fb.value} | Q^T A = [R;0]: chol {res[2][2]:.1e} / {res[2][3]:.1e}  "
                  f"blocked {res[1][2]:.1e} / {res[1][3]:.1e}  per-column {res[0][2]:.1e} / {res[0][3]:.1e} | Q^T Q - I: {res[2][5]:.1e} {res[1][5]:.1e} {res[0][5]:.1e}", flush=True)
