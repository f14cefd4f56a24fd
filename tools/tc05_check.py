"""Diagnostic for the tcgen05 FP32 kernels (run on the GPU box): each case compares one kernel against an
fp64 torch reference and prints an error map coarse enough to spot descriptor / swizzle mistakes.

    python tools/tc05_check.py <ru|nn|tn|probe|chain> [m n b]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svdsolver_b200 import capi  # noqa: E402


def errmap(got, ref, br, bc, name):
    err = (got.double() - ref).abs()
    scale = ref.abs().max().item()
    rel = err.max().item() / scale
    print(f"{name}: shape {tuple(ref.shape)} max rel err {rel:.3e}", flush=True)
    if rel > 1e-4:
        m, n = ref.shape
        R, Cc = (m + br - 1) // br, (n + bc - 1) // bc
        print(f"  error map, blocks of {br} x {bc} (log10 of max rel err; '.' < 1e-5):")
        for i in range(min(R, 40)):
            row = ""
            for j in range(min(Cc, 64)):
                e = err[i * br:(i + 1) * br, j * bc:(j + 1) * bc].max().item() / scale
                row += "." if e < 1e-5 else str(min(9, max(0, int(-np.log10(max(e, 1e-9))))))
            print("  " + row)
    return rel


def main():
    mode = sys.argv[1]
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 640
    b = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    force = int(os.environ.get("TC05_MODE", "2"))
    torch.manual_seed(0)
    dev = "cuda"
    h = capi.Handle(max(m, n) + 64, b, np.float32)
    h.set_tc05(force)
    if mode == "probe":
        for k, nm in ((3, "tf32 mma.sync"), (4, "tf32 tcgen05")):
            print(nm, h.probe_peak(k), "TFLOP/s", flush=True)
        return 0
    ld = n + 8
    C = torch.rand(m, ld, device=dev) * 5
    V = torch.rand(m, b, device=dev) - 0.5
    Ut = torch.rand(n, b, device=dev) - 0.5
    Q = torch.rand(b, n, device=dev) - 0.5
    torch.cuda.synchronize()
    worst = 0.0
    reps = int(os.environ.get("TC05_REPS", "1"))
    if mode in ("ru", "all"):
        C2 = C.clone()
        torch.cuda.synchronize()
        h.rank_update_dev(C2.data_ptr(), ld, m, n, b, V.data_ptr(), Q.data_ptr(), n)
        h.synchronize()
        ref = C[:, :n].double() + V.double() @ Q.double()
        worst = max(worst, errmap(C2[:, :n], ref, 8, 32, "rank_update"))
        print("  padding untouched:", bool(torch.equal(C2[:, n:], C[:, n:])), flush=True)
        if reps > 1:
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            s = torch.cuda.Stream()
            h.set_stream(s.cuda_stream)
            with torch.cuda.stream(s):
                t0.record(s)
                for _ in range(reps):
                    h.rank_update_dev(C2.data_ptr(), ld, m, n, b, V.data_ptr(), Q.data_ptr(), n)
                t1.record(s)
            s.synchronize()
            ms = t0.elapsed_time(t1) / reps
            print(f"  rank_update {ms*1e3:.1f} us  {2.0*m*n*b/ms/1e9:.2f} TFLOP/s  {8.0*m*n/ms/1e6:.0f} GB/s", flush=True)
            h.set_stream(0)
    if mode in ("nn", "all"):
        W2 = torch.zeros(m, b, device=dev)
        torch.cuda.synchronize()
        h.gemm_nn_dev(C.data_ptr(), ld, m, n, b, Ut.data_ptr(), W2.data_ptr())
        h.synchronize()
        ref2 = C[:, :n].double() @ Ut.double()
        worst = max(worst, errmap(W2, ref2, 8, 8, "gemm_nn"))
        if reps > 1:
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            s = torch.cuda.Stream()
            h.set_stream(s.cuda_stream)
            with torch.cuda.stream(s):
                t0.record(s)
                for _ in range(reps):
                    h.gemm_nn_dev(C.data_ptr(), ld, m, n, b, Ut.data_ptr(), W2.data_ptr())
                t1.record(s)
            s.synchronize()
            ms = t0.elapsed_time(t1) / reps
            print(f"  gemm_nn {ms*1e3:.1f} us  {2.0*m*n*b/ms/1e9:.2f} TFLOP/s  {4.0*m*n/ms/1e6:.0f} GB/s", flush=True)
            h.set_stream(0)
    if mode in ("tn", "all"):
        W = torch.zeros(b, n, device=dev)
        torch.cuda.synchronize()
        h.gemm_tn_dev(V.data_ptr(), C.data_ptr(), ld, m, n, b, W.data_ptr())
        h.synchronize()
        ref = V.double().T @ C[:, :n].double()
        worst = max(worst, errmap(W, ref, 8, 32, "gemm_tn"))
        if reps > 1:
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            s = torch.cuda.Stream()
            h.set_stream(s.cuda_stream)
            with torch.cuda.stream(s):
                t0.record(s)
                for _ in range(reps):
                    h.gemm_tn_dev(V.data_ptr(), C.data_ptr(), ld, m, n, b, W.data_ptr())
                t1.record(s)
            s.synchronize()
            ms = t0.elapsed_time(t1) / reps
            print(f"  gemm_tn {ms*1e3:.1f} us  {2.0*m*n*b/ms/1e9:.2f} TFLOP/s  {4.0*m*n/ms/1e6:.0f} GB/s", flush=True)
            h.set_stream(0)
    print("WORST", worst, flush=True)
    return 0 if worst < 1e-4 else 1


if __name__ == "__main__":
    sys.exit(main())
