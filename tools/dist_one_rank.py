"""The distributed stage-1 driver on ONE rank (no NCCL): same code path as N ranks minus the collectives -- used to debug the
distributed LQ panel.   python tools/dist_one_rank.py <n> <band> <f32|f64>"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svdsolver_b200 import capi, distributed as D  # noqa: E402

n, b = int(sys.argv[1]), int(sys.argv[2])
dt = np.float32 if sys.argv[3] == "f32" else np.float64
tdt = torch.float32 if dt == np.float32 else torch.float64
a = torch.empty(n, n, device="cuda", dtype=tdt)
with capi.Handle(64, 32, dt) as hf:
    hf.fill_uniform_dev(a.data_ptr(), n * n, 586 + n, 0.0, 5.0)
    hf.synchronize()
a0 = a.clone()
for lqd in (1, 0):
    a.copy_(a0)
    with D.DistHandle(n, b, dt, 0, 1, (ctypes.c_ubyte * 128)()) as dh:
        dh.configure(tc05_mode=int(os.environ.get("TC05", "1")))
        dh.configure_panels(lqd)
        try:
            dh.dense_to_band_dev(a.data_ptr())
            torch.cuda.synchronize()
            fro = float(torch.linalg.norm(a.double()) / torch.linalg.norm(a0.double()) - 1)
            below = float(a.tril(-1).abs().max())
            print(f"n={n} band={b} lq_dist={lqd}: ok, |A|_F drift {fro:.2e}, below diag {below:.1e}", flush=True)
        except capi.SvdB200Error as ex:
            print(f"n={n} band={b} lq_dist={lqd}: {ex}", flush=True)
