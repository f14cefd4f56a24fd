"""BASELINE.json configs[2] (n=16384 double full SVD, band 64) and configs[4] (batched 256x256 double SVDs,
one GPU's share) on one B200: device times per stage and accuracy checks (test-only torch/cuSOLVER reference).

    python tools/config_runs.py c3 [n] [band]      python tools/config_runs.py c5 [count] [n] [band]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svdsolver_b200 import capi  # noqa: E402


def flops(n):
    return 8.0 * n ** 3 / 3.0


def c3(n=16384, band=64, dt=np.float64):
    tdt = torch.float64 if dt == np.float64 else torch.float32
    h = capi.Handle(n, band, dt)
    h.set_stage2_schedule(int(os.environ.get("S2_COMPLETE", "0")))
    s = torch.cuda.Stream()
    h.set_stream(s.cuda_stream)
    a = torch.empty(n, n, device="cuda", dtype=tdt)
    h.fill_uniform_dev(a.data_ptr(), n * n, 586 + n, 0.0, 5.0)
    torch.cuda.synchronize()
    a0 = a.clone()
    d = torch.empty(n, device="cuda", dtype=tdt)
    e = torch.empty(n, device="cuda", dtype=tdt)
    sig = torch.empty(n, device="cuda", dtype=tdt)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.cuda.stream(s):
        ev[0].record(s)
        h.dense_to_band_dev(a.data_ptr(), n, band)
        ev[1].record(s)
        h.band_to_bidiag_dev(a.data_ptr(), n, band, d.data_ptr(), e.data_ptr())
        ev[2].record(s)
        st = None
        try:
            h.bidiag_qr_dev(d.data_ptr(), e.data_ptr(), n, sig.data_ptr())
        except Exception as ex:   # NOCONV is reported, the timings stay valid
            st = str(ex)
        ev[3].record(s)
    s.synchronize()
    t1, t2, t3 = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
    out = {"config": f"{n}x{n} {np.dtype(dt).name} full SVD, band {band}", "stage1_ms": round(t1, 1), "stage2_ms": round(t2, 1),
           "qr_ms": round(t3, 1), "qr_status": st, "stage1_tflops": round(flops(n) / t1 * 1e-9, 2),
           "reduction_gflops": round(flops(n) / (t1 + t2) * 1e-6, 1), "total_ms": round(t1 + t2 + t3, 1)}
    print(json.dumps(out), flush=True)
    # accuracy: sigma vs cuSOLVER (test-only reference) on the same input; invariants
    fro = float(torch.linalg.norm(a0.double()))
    sg = sig.double()
    out["fro_rel_err"] = abs(float(torch.sqrt((sg * sg).sum())) - fro) / fro
    t0 = time.time()
    ref = torch.linalg.svdvals(a0)
    torch.cuda.synchronize()
    out["cusolver_svdvals_s"] = round(time.time() - t0, 2)
    out["sigma_rel_err_vs_cusolver"] = float((sg - ref.double()).abs().max() / ref.double()[0])
    # the reference's stage-2 schedule loses orthogonality at the matrix boundary: report the band's sigma too
    print(json.dumps(out), flush=True)
    h.close()


def c5(count=1024, n=256, band=32, dt=np.float64):
    tdt = torch.float64 if dt == np.float64 else torch.float32
    h = capi.Handle(n, band, dt)
    s = torch.cuda.Stream()
    h.set_stream(s.cuda_stream)
    a = torch.empty(count, n, n, device="cuda", dtype=tdt)
    for i in range(count):
        h.fill_uniform_dev(a[i].data_ptr(), n * n, 586 + i, 0.0, 5.0)
    torch.cuda.synchronize()
    a0 = a.clone()
    sig = torch.empty(count, n, device="cuda", dtype=tdt)
    res = []
    for rep in range(3):
        a.copy_(a0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record(s)
            h.svdvals_batched_dev(a.data_ptr(), count, n, band, sig.data_ptr())
            e1.record(s)
        s.synchronize()
        res.append(e0.elapsed_time(e1))
    ms = min(res)
    ref = torch.linalg.svdvals(a0[:64])
    err = float(((sig[:64].double() - ref.double()).abs().amax(dim=1) / ref.double()[:, 0]).max())
    print(json.dumps({"config": f"batched {count} x {n}x{n} {np.dtype(dt).name} SVDs, band {band} (one GPU)", "ms": round(ms, 1),
                      "matrices_per_s": round(count / ms * 1e3, 1), "gflops_reduction_equiv": round(count * flops(n) / ms * 1e-6, 1),
                      "sigma_rel_err_vs_cusolver_first64": err, "all_ms": [round(x, 1) for x in res]}), flush=True)
    h.close()


if __name__ == "__main__":
    mode = sys.argv[1]
    args = [int(x) for x in sys.argv[2:]]
    if mode == "c3":
        c3(*args)
    elif mode == "c3f":
        c3(*(args or [16384, 64]), dt=np.float32)
    elif mode == "c5":
        c5(*args)
