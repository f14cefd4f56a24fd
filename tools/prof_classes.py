"""Per-kernel-class device times of one reduction (svdb200_set_profile): python tools/prof_classes.py n band dtype [stages]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from svdsolver_b200 import capi

n = int(sys.argv[1]); band = int(sys.argv[2]); suf = sys.argv[3]; stages = sys.argv[4] if len(sys.argv) > 4 else "s1"
dt = {"f64": (np.float64, torch.float64), "f32": (np.float32, torch.float32)}[suf]
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
with capi.Handle(n, band, dt[0]) as h:
    h.set_stream(s.cuda_stream)
    a = torch.empty(n, n, device="cuda", dtype=dt[1])
    d = torch.empty(n, device="cuda", dtype=dt[1]); e = torch.empty(n, device="cuda", dtype=dt[1])
    for rep in range(2):
        h.fill_uniform_dev(a.data_ptr(), n * n, 586 + n, 0.0, 5.0)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        h.set_profile(rep == 1); h.reset_profile()
        e0.record(s)
        h.dense_to_band_dev(a.data_ptr(), n, band)
        e1.record(s)
        if "s2" in stages:
            h.band_to_bidiag_dev(a.data_ptr(), n, band, d.data_ptr(), e.data_ptr())
        e2.record(s)
        torch.cuda.synchronize()
        t1 = e0.elapsed_time(e1); t2 = e1.elapsed_time(e2)
        print(f"rep {rep} (profile={'on' if rep else 'off'}): stage1 {t1:.2f} ms = {8*n**3/3/t1*1e-9:.2f} TFLOP/s   stage2 {t2:.2f} ms")
    for k, v in h.get_profile().items():
        if v["launches"]:
            extra = f"{v['work']/(v['ms']*1e-3)*1e-12:.2f} TFLOP/s" if k not in ("stage2", "qr") else f"{v['work']/(v['ms']*1e-3)*1e-9:.1f} GB/s"
            print(f"  {k:12s} {v['ms']:10.2f} ms  {v['launches']:6d} launches  avg {v['ms']/v['launches']*1e3:9.1f} us   {extra}")
    band_m = a.cpu().numpy() if n <= 4096 else None
