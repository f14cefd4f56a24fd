"""GPU diagnostic: panel-order stage 1 with the blocked panel kernel (svdb200_set_panel_kernel 1) and with the per-column
kernels (0) -- band vs the CPU oracle on small sizes, agreement between the two and timing on large ones."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from svdsolver_b200 import capi  # noqa: E402
from svdsolver_b200.synth import uniform_matrix  # noqa: E402
from conftest import Oracle, band_rel  # noqa: E402

orc = Oracle(os.path.join(ROOT, "oracle", "libsvd_oracle.so"))
DT = {"f32": (np.float32, torch.float32), "f64": (np.float64, torch.float64)}


def set_blk(h, on):
    st = capi.lib().svdb200_set_panel_kernel(h.h, ctypes.c_int(on))
    assert st == 0


def small():
    for n, b, suf in [(64, 8, "f64"), (128, 16, "f64"), (256, 32, "f64"), (256, 64, "f64"), (320, 32, "f32"), (512, 8, "f32"), (512, 8, "f64"),
                      (640, 32, "f64"), (640, 32, "f32"), (768, 64, "f32"), (768, 64, "f64"), (512, 16, "f32"), (1024, 32, "f64")]:
        dt = DT[suf][0]
        a = uniform_matrix(n, n, 586 + n + b, 0.0, 5.0, dt)
        ref = orc.brd_p1_panel(a, b)
        res = []
        for blk in (1, 0):
            with capi.Handle(n, b, dt) as h:
                set_blk(h, blk)
                out = h.dense_to_band(a, b, capi.ORDER_PANEL)
            res.append(band_rel(out, ref, b))
            res.append(band_rel(np.abs(out), np.abs(ref), b))
        print(f"small n={n} b={b} {suf}: rel vs oracle blocked {res[0]:.3e} (|.| {res[1]:.3e})  per-column {res[2]:.3e} (|.| {res[3]:.3e})", flush=True)


def large():
    for n, b, suf in [(1920, 32, "f64"), (3840, 32, "f64"), (3840, 32, "f32"), (4096, 64, "f64"), (8192, 64, "f32"), (16384, 64, "f64"), (16384, 64, "f32")]:
        dt, tdt = DT[suf]
        a = torch.empty(n, n, device="cuda", dtype=tdt)
        outs, times = [], []
        with capi.Handle(n, b, dt) as h:
            for blk in (1, 0):
                set_blk(h, blk)
                best = None
                for rep in range(2):
                    h.fill_uniform_dev(a.data_ptr(), n * n, 586 + n, 0.0, 5.0)
                    h.synchronize()
                    t0 = time.perf_counter()
                    h.dense_to_band_dev(a.data_ptr(), n, b)
                    h.synchronize()
                    t = (time.perf_counter() - t0) * 1e3
                    best = t if best is None else min(best, t)
                times.append(best)
                outs.append(torch.triu(torch.tril(a, b)).clone() if n <= 8192 else torch.stack([torch.diagonal(a, k)[: n - b] for k in range(b + 1)]))
                if blk == 1:
                    h.reset_profile(); h.set_profile(True)
                    h.fill_uniform_dev(a.data_ptr(), n * n, 586 + n, 0.0, 5.0)
                    h.dense_to_band_dev(a.data_ptr(), n, b)
                    h.synchronize(); h.set_profile(False)
                    p = h.get_profile()["panel"]
                    pstr = f"panels {p['ms']:.1f} ms / {p['launches']}"
        diff = float((outs[0] - outs[1]).abs().max() / outs[1].abs().max())
        adiff = float((outs[0].abs() - outs[1].abs()).abs().max() / outs[1].abs().max())
        print(f"large n={n} b={b} {suf}: stage1 blocked {times[0]:.1f} ms ({pstr})  per-column {times[1]:.1f} ms   band diff {diff:.3e} (|.| {adiff:.3e})", flush=True)
        del a, outs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    small()
    large()
