"""torchrun worker: N-axis sharded run vs a single-GPU run of the same task (used by test_multi_gpu.py).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_worker.py
"""
import json
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from coda_b200 import CODA, SyntheticDataset  # noqa: E402
from coda_b200.dist import LocalComm, TorchComm  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    H, N, C, steps = 48, 30011, 14, 10
    out = {}
    for mode in ("incremental", "recompute"):
        ds = SyntheticDataset(H, N, C, seed=4, device=dev, rank=rank, world=world)
        random.seed(0)
        sh = CODA(ds, mode=mode, comm=TorchComm())
        one = None
        if rank == 0:
            full = SyntheticDataset(H, N, C, seed=4, device=dev)
            random.seed(0)
            one = CODA(full, mode=mode, comm=LocalComm())
        picks, picks1, errs = [], [], {"pi_hat_equal": True, "pbest": 0.0, "eig": 0.0, "D_equal": True}
        random.seed(1)
        st = random.getstate()
        for k in range(steps):
            random.setstate(st)
            idx, q = sh.get_next_item_to_label()
            st_after = random.getstate()
            eig_all = [torch.empty(0)] * world
            parts = [None] * world
            dist.all_gather_object(parts, sh.engine.eig.cpu())
            if rank == 0:
                random.setstate(st)
                i1, q1 = one.get_next_item_to_label()
                assert random.getstate() == st_after
                picks1.append(i1)
                errs["eig"] = max(errs["eig"], float((torch.cat(parts) - one.engine.eig.cpu()).abs().max()))
            st = st_after
            picks.append(idx)
            t = int(ds.labels_host[idx])
            sh.add_label(idx, t, q)
            b = int(sh.get_best_model_prediction())
            if rank == 0:
                one.add_label(idx, t, q)
                b1 = int(one.get_best_model_prediction())
                assert b == b1
                errs["pi_hat_equal"] &= bool(torch.equal(sh.pi_hat, one.pi_hat))
                errs["D_equal"] &= bool(torch.equal(sh.dirichlets, one.dirichlets))
                errs["pbest"] = max(errs["pbest"], float((sh.get_pbest() - one.get_pbest()).abs().max()))
        allp = [None] * world
        dist.all_gather_object(allp, picks)
        out[mode] = dict(picks=picks, picks_single=picks1, same_on_all_ranks=all(p == picks for p in allp), **errs)
        # host-free CUDA-graph loop, continuing from the API steps: exchanges inside the kernels over peer memory
        sh.run_steps(6, ds.labels)
        hi = sh.history()[0].tolist()
        allh = [None] * world
        dist.all_gather_object(allh, hi)
        out[mode]["loop_same_on_all_ranks"] = all(h == hi for h in allh)
        if rank == 0:
            one.run_steps(6, full.labels)
            out[mode]["loop_picks"], out[mode]["loop_picks_single"] = hi, one.history()[0].tolist()
            out[mode]["loop_D_equal"] = bool(torch.equal(sh.dirichlets, one.dirichlets))
            out[mode]["loop_pi_hat_equal"] = bool(torch.equal(sh.pi_hat, one.pi_hat))
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
