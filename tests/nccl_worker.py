"""Worker for tests/test_gpu_nccl.py: one rank per GPU, NCCL transport.
Rank 0 also runs the oracle over all parts and gathers the parity errors."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import make_case, make_oracle, rel_l2  # noqa: E402
from phasta_b200.solver import PhastaGPU, nccl_unique_id  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    case = make_case(4 * world, 5, 4, nparts=world, bc="channel", etol=1e-7, Kspace=30, max_seg=9)
    params, tables, parts, states = case
    g = PhastaGPU(parts[rank], params, tables, device=lr)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    g.comm_init(bytes(idt.cpu().tolist()))
    y, ac = states[rank]
    res, Dy = g.SolGMRe(y, ac)
    o = make_oracle(case)
    iKs, lG = o.SolGMRe()
    op = o.parts[rank]
    errs = torch.tensor([rel_l2(res, op.res), rel_l2(g.rmes, op.rmes), rel_l2(Dy, op.Dy),
                         float(abs(g.iKs - iKs))], dtype=torch.float64, device="cuda")
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    if rank == 0:
        e = errs.cpu().numpy()
        print("NCCL_PARITY world=%d res=%.2e rmes=%.2e Dy=%.2e dIKs=%d iKs=%d" % (world, e[0], e[1], e[2], int(e[3]), iKs),
              flush=True)
        assert e[0] < 1e-10 and e[1] < 1e-10 and e[2] < 1e-8 and e[3] == 0
    g.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
