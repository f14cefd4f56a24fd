"""Runs svdb200_tc05_selftest for the four operand-major combinations and checks D = A B and the smem layout."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svdsolver_b200 import capi  # noqa: E402


def swz(byte_off):
    """address of logical byte offset inside a 128B-swizzled region (1024 B atoms)"""
    chunk = (byte_off >> 4) & 7
    row = (byte_off >> 7) & 7
    return (byte_off & ~0x70) | ((chunk ^ row) << 4)


def swz32(byte_off):
    """128B span, 32-byte atoms: chunk32 ^= row % 4 (Swizzle<2,5,2>)"""
    chunk = (byte_off >> 5) & 3
    row = (byte_off >> 7) & 3
    return (byte_off & ~0x60) | ((chunk ^ row) << 5)


def main():
    h = capi.Handle(256, 64, np.float32)
    rng = np.random.default_rng(0)
    A = (rng.integers(-8, 9, size=(128, 32)) / 8.0).astype(np.float32)     # exact in TF32
    B = (rng.integers(-8, 9, size=(32, 64)) / 8.0).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    ok = True
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            a_host = A if a_mn == 0 else np.ascontiguousarray(A.T)
            b_host = np.ascontiguousarray(B.T) if b_mn == 0 else B
            a = torch.from_numpy(a_host.copy()).cuda()
            b = torch.from_numpy(b_host.copy()).cuda()
            out = torch.full((128 * 64 + 1,), -777.0, device="cuda")
            dump = torch.full((6144,), -777.0, device="cuda")
            torch.cuda.synchronize()
            st = capi.lib().svdb200_tc05_selftest(h.h, ctypes.c_int(a_mn), ctypes.c_int(b_mn), ctypes.c_void_p(a.data_ptr()),
                                                  ctypes.c_void_p(b.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                                  ctypes.c_void_p(dump.data_ptr()))
            torch.cuda.synchronize()
            o = out.cpu().numpy()
            d = dump.cpu().numpy()
            D = o[:128 * 64].reshape(128, 64)
            err = np.abs(D - ref).max()
            print(f"a_mn={a_mn} b_mn={b_mn} status={st} tmem=0x{np.float32(o[-1]).view(np.uint32):08x} max err {err:.3e}"
                  f"  D[0,:4]={D[0,:4]} ref[0,:4]={ref[0,:4]}", flush=True)
            # expected smem images
            expA = np.zeros(4096, np.float32)
            if a_mn == 0:
                for r in range(128):
                    for k in range(32):
                        expA[swz(r * 128 + k * 4) // 4] = A[r, k]
            else:
                for sl in range(4):
                    for k in range(32):
                        for i in range(32):
                            expA[(sl * 4096 + swz32(k * 128 + i * 4)) // 4] = A[sl * 32 + i, k]
            expB = np.zeros(2048, np.float32)
            if b_mn == 0:
                for nn in range(64):
                    for k in range(32):
                        expB[swz(nn * 128 + k * 4) // 4] = B[k, nn]
            else:
                for sl in range(2):
                    for k in range(32):
                        for i in range(32):
                            expB[(sl * 4096 + swz32(k * 128 + i * 4)) // 4] = B[k, sl * 32 + i]
            okA = np.array_equal(d[:4096], expA)
            okB = np.array_equal(d[4096:], expB)
            print(f"   smem A image matches expectation: {okA}; B: {okB}", flush=True)
            if not okA:
                print("   A dump[:16]", d[:16], "exp", expA[:16])
            if not okB:
                print("   B dump[:16]", d[4096:4112], "exp", expB[:16])
            if err > 1e-5:
                ok = False
                # which rows / cols are wrong
                bad = np.abs(D - ref) > 1e-5
                print("   bad rows:", np.where(bad.any(axis=1))[0][:40], " bad cols:", np.where(bad.any(axis=0))[0][:64])
                nz = np.count_nonzero(D)
                print("   nonzeros in D:", nz, " D==-777:", np.count_nonzero(D == -777.0))
    return 0 if ok else 1




def probe_rounding():
    """How does kind::tf32 treat the low 13 mantissa bits of an fp32 operand in shared memory: truncation or rounding?"""
    h = capi.Handle(256, 64, np.float32)
    rng = np.random.default_rng(5)
    A = rng.standard_normal((128, 32)).astype(np.float32)
    B = rng.standard_normal((32, 64)).astype(np.float32)

    def trunc(x):
        return (x.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)

    def rna(x):
        u = x.view(np.uint32).astype(np.uint64) + 0x1000
        return (u & 0xffffe000).astype(np.uint32).view(np.float32)

    a = torch.from_numpy(A.copy()).cuda()
    b = torch.from_numpy(np.ascontiguousarray(B.T).copy()).cuda()
    out = torch.zeros(128 * 64 + 1, device="cuda")
    dump = torch.zeros(6144, device="cuda")
    capi.lib().svdb200_tc05_selftest(h.h, ctypes.c_int(0), ctypes.c_int(0), ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()),
                                     ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(dump.data_ptr()))
    torch.cuda.synchronize()
    D = out.cpu().numpy()[:128 * 64].reshape(128, 64).astype(np.float64)
    for name, f in (("truncate", trunc), ("round-nearest", rna), ("exact fp32", lambda x: x)):
        ref = f(A).astype(np.float64) @ f(B).astype(np.float64)
        print(f"model {name:14s}: max |D - ref| = {np.abs(D - ref).max():.3e}", flush=True)


if len(sys.argv) > 1 and sys.argv[1] == "rounding":
    probe_rounding()
    sys.exit(0)

if __name__ == "__main__":
    sys.exit(main())
