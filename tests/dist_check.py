"""Multi-GPU parity check, run under torchrun on N GPUs of one box (collected by pytest through
tests/test_gpu_dist.py::test_dist_stage1_under_torchrun when >= 2 GPUs are visible):

  * block-cyclic stage 1 on N ranks == single-GPU panel-order stage 1 (tolerance), f64 and f32, at sizes that reach the
    tcgen05 trailing update (forced with set_tc05(2)) and the multi-cluster panel kernel;
  * the band gathered through svdb200_dist_gather_band_dev_* == the band of the gathered dense result;
  * svdb200_dist_svdvals_dev_* (stage 1 distributed, stage 2 + sigma on rank 0) == singular values of the input
    (complete stage-2 schedule => orthogonally equivalent).

    torchrun --nproc-per-node N tests/dist_check.py [small|large]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svdsolver_b200 import capi, distributed as D  # noqa: E402
from svdsolver_b200.synth import uniform_matrix  # noqa: E402

# (n, band, dtype, tolerance, lq_distributed): 1 = LQ panels through the local Gram matrix + one all-reduce (default),
# 0 = all-gather of the row panel + redundant factorisation (the fallback path)
CASES = {
    "small": [(1024, 32, np.float64, 1e-10, 1), (768, 64, np.float32, 1e-4, 1), (512, 4, np.float64, 1e-10, 1), (1024, 32, np.float64, 1e-10, 0),
              (768, 16, np.float32, 1e-4, 1)],
    "large": [(8192, 64, np.float32, 1e-4, 1), (8192, 64, np.float64, 1e-10, 1), (6144, 32, np.float32, 1e-4, 1), (4096, 64, np.float32, 1e-4, 0)],
}


def gather_dense(loc, n, band, rank, world, tdt):
    parts = [D.local_cols(n, band, r, world) for r in range(world)]
    wmax = max(parts)
    pad = torch.zeros(n, wmax, dtype=tdt, device="cuda")
    pad[:, : loc.shape[1]] = loc
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    if rank != 0:
        return None
    return D.gather_block_cyclic([o[:, : parts[r]].cpu().numpy() for r, o in enumerate(outs)], band, n)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ok = True
    for n, band, dt, tol, lqd in CASES[which]:
        tdt = torch.float32 if dt == np.float32 else torch.float64
        a = uniform_matrix(n, n, 586 + n, 0.0, 5.0, dt)
        s = torch.cuda.Stream()
        # ---- stage 1 + band gather -----------------------------------------------------------------------------
        uid = D.exchange_unique_id(rank, world)      # a ncclUniqueId may seed exactly one communicator
        loc = torch.from_numpy(D.scatter_block_cyclic(a, band, rank, world)).cuda()
        packed = torch.zeros(n, band + 1, dtype=tdt, device="cuda")
        with D.DistHandle(n, band, dt, rank, world, uid, device=lr) as h:
            h.set_stream(s.cuda_stream)
            h.configure(tc05_mode=2)
            h.configure_panels(lqd)
            torch.cuda.synchronize()
            h.dense_to_band_dev(loc.data_ptr())
            h.gather_band_dev(loc.data_ptr(), packed.data_ptr())
            s.synchronize()
        full = gather_dense(loc, n, band, rank, world, tdt)
        # ---- distributed svdvals --------------------------------------------------------------------------------
        uid2 = D.exchange_unique_id(rank, world)
        loc2 = torch.from_numpy(D.scatter_block_cyclic(a, band, rank, world)).cuda()
        sigma = torch.zeros(n, dtype=tdt, device="cuda")
        with D.DistHandle(n, band, dt, rank, world, uid2, device=lr) as h:
            h.set_stream(s.cuda_stream)
            h.configure(stage2_schedule=1, qr_method=2)
            h.configure_panels(lqd)
            torch.cuda.synchronize()
            h.svdvals_dev(loc2.data_ptr(), sigma.data_ptr())
            s.synchronize()
        if rank == 0:
            with capi.Handle(n, band, dt, device=lr) as h1:
                h1.set_tc05(2)
                ref = h1.dense_to_band(a, band, capi.ORDER_PANEL)
            num = max(float(np.abs(np.diagonal(full, k).astype(np.float64) - np.diagonal(ref, k).astype(np.float64)).max()) for k in range(band + 1))
            rel = num / float(np.abs(ref).max())
            if dt == np.float32 and rel > tol:       # a tiny float pivot may flip a row / column sign (tests/conftest.py)
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                from conftest import band_rel_mod_signs
                rel, flips = band_rel_mod_signs(full, ref, band)
            below = float(np.abs(np.tril(full, -1)).max())
            pk = D.unpack_band(packed.cpu().numpy(), n, band)
            band_ok = np.array_equal(pk, np.triu(np.tril(full, band)))
            fa = np.linalg.norm(a.astype(np.float64))
            fro = abs(np.linalg.norm(full.astype(np.float64)) - fa) / fa
            s_ref = torch.linalg.svdvals(torch.from_numpy(a.astype(np.float64)).cuda()).cpu().numpy()    # test-only reference
            serr = float(np.abs(sigma.cpu().numpy().astype(np.float64) - s_ref).max() / s_ref[0])
            stol = 2e-5 if dt == np.float32 else 1e-11
            good = rel <= tol and below == 0.0 and band_ok and fro <= (1e-5 if dt == np.float32 else 1e-12) and serr <= stol
            ok &= good
            print(f"dist n={n} band={band} {np.dtype(dt).name} ranks={world} lq_dist={lqd}: band rel diff vs 1-GPU {rel:.3e}, below-diag {below:.1e}, "
                  f"gathered band == dense band: {band_ok}, |A|_F drift {fro:.2e}, dist_svdvals vs LAPACK {serr:.2e}  {'OK' if good else 'FAIL'}", flush=True)
        del loc, loc2
        torch.cuda.empty_cache()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
