"""Multi-GPU parity check (run under torchrun on N GPUs of one box; not collected by pytest):
block-cyclic stage 1 on N ranks == single-GPU panel-order stage 1 (tolerance), for f64 and f32."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svdsolver_b200 import capi, distributed as D  # noqa: E402
from svdsolver_b200.synth import uniform_matrix  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ok = True
    for n, band, dt, tdt, tol in ((1024, 32, np.float64, torch.float64, 1e-10), (768, 64, np.float32, torch.float32, 1e-4), (512, 4, np.float64, torch.float64, 1e-10)):
        uid = D.exchange_unique_id(rank, world)      # a ncclUniqueId may seed exactly one communicator
        a = uniform_matrix(n, n, 586 + n, 0.0, 5.0, dt)
        loc = torch.from_numpy(D.scatter_block_cyclic(a, band, rank, world)).cuda()
        s = torch.cuda.Stream()
        with D.DistHandle(n, band, dt, rank, world, uid, device=lr) as h:
            h.set_stream(s.cuda_stream)
            torch.cuda.synchronize()
            h.dense_to_band_dev(loc.data_ptr())
            s.synchronize()
        parts = [torch.empty(n, D.local_cols(n, band, r, world), dtype=tdt, device="cuda") for r in range(world)]
        # all_gather needs equal shapes: pad to the widest part
        wmax = max(p.shape[1] for p in parts)
        pad = torch.zeros(n, wmax, dtype=tdt, device="cuda")
        pad[:, : loc.shape[1]] = loc
        outs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(outs, pad)
        if rank == 0:
            full = D.gather_block_cyclic([o[:, : parts[r].shape[1]].cpu().numpy() for r, o in enumerate(outs)], band, n)
            with capi.Handle(n, band, dt, device=lr) as h1:
                ref = h1.dense_to_band(a, band, capi.ORDER_PANEL)
            num = max(float(np.abs(np.diagonal(full, k).astype(np.float64) - np.diagonal(ref, k).astype(np.float64)).max()) for k in range(band + 1))
            rel = num / float(np.abs(ref).max())
            good = rel <= tol and float(np.abs(np.tril(full, -1)).max()) == 0.0
            ok &= good
            print(f"dist stage1 n={n} band={band} {np.dtype(dt).name} ranks={world}: rel diff vs 1-GPU {rel:.3e} {'OK' if good else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
