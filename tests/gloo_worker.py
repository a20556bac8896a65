"""Worker for tests/test_gloo_halo.py: world_size ranks on CPU over torch.distributed/gloo.

Each rank builds ITS OWN mesh part (as one MPI rank of the reference reads its own geombc file), then runs the
halo exchange of common/commu.f:95-297 over real inter-process messages, driven only by its own ilwork:
  'in '  slaves (iacc=0) send their segment values, masters (iacc=1) receive and ADD        (commu.f:185-294)
  'out'  masters send, slaves receive and OVERWRITE                                          (commu.f:145-183)
and checks the result against the oracle's in-process commu over all parts (built redundantly on every rank).
This covers the host-side partition description -- task order, tags, peer ids, segment lists, master/slave flags --
that phb200_init hands to the NCCL transport on the GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import make_case, make_oracle  # noqa: E402
from phasta_b200 import make_box  # noqa: E402


def tasks_of(ilwork):
    il, pos, out = ilwork, 1, []
    for _ in range(int(il[0])):
        tag, iacc, iother, nseg = (int(v) for v in il[pos:pos + 4])
        segs = [(int(il[pos + 4 + 2 * s]), int(il[pos + 5 + 2 * s])) for s in range(nseg)]
        out.append((tag, iacc, iother, segs))
        pos += 4 + 2 * nseg
    return out


def commu_gloo(v, ilwork, code):
    """commu(global, ilwork, n, code) on v(nshg, n) (column-major), over torch.distributed send/recv"""
    n = v.shape[1]
    reqs, recvs = [], []
    for tag, iacc, iother, segs in tasks_of(ilwork):
        idx = np.concatenate([np.arange(a - 1, a - 1 + ln) for a, ln in segs])
        sending = (iacc == 0) if code == "in" else (iacc == 1)
        if sending:
            buf = torch.from_numpy(np.ascontiguousarray(v[idx, :]))
            reqs.append(dist.isend(buf, dst=iother, tag=tag))
        else:
            buf = torch.empty((idx.size, n), dtype=torch.float64)
            reqs.append(dist.irecv(buf, src=iother, tag=tag))
            recvs.append((idx, buf))
    for r in reqs:
        r.wait()
    for idx, buf in recvs:      # in ilwork order (commu.f:268-294)
        if code == "in":
            v[idx, :] += buf.numpy()
        else:
            v[idx, :] = buf.numpy()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    nx, ny, nz = 4 * world, 3, 3
    mine = make_box(nx, ny, nz, nparts=world, bc="channel", only_rank=rank, max_seg=7)[0]
    assert mine.rank == rank and mine.numpe == world
    # the same vector on every rank's copy of a shared node would hide a missing exchange: make it rank-dependent
    n = 5
    rng = np.random.default_rng(100 + rank)
    v = np.asfortranarray(rng.standard_normal((mine.nshg, n)))
    v0 = v.copy(order="F")
    commu_gloo(v, mine.ilwork, "in")
    v_in = v.copy(order="F")
    commu_gloo(v, mine.ilwork, "out")
    # expected: the oracle's in-process commu over ALL parts, with the same per-rank vectors
    case = make_case(nx, ny, nz, nparts=world, bc="channel", max_seg=7)
    o = make_oracle(case)
    assert np.array_equal(case[2][rank].ilwork, mine.ilwork) and np.array_equal(case[2][rank].x, mine.x)
    vs = [np.asfortranarray(np.random.default_rng(100 + r).standard_normal((case[2][r].nshg, n))) for r in range(world)]
    assert np.array_equal(vs[rank], v0)
    o.commu(vs, n, "in")
    ok_in = np.array_equal(vs[rank], v_in)
    o.commu(vs, n, "out")
    ok_out = np.array_equal(vs[rank], v)
    # after 'in' + 'out' every copy of a shared node holds the same value: check across ranks by global id
    gsum = torch.zeros(((nx + 1) * (ny + 1) * (nz + 1), n), dtype=torch.float64)
    gcnt = torch.zeros((nx + 1) * (ny + 1) * (nz + 1), dtype=torch.float64)
    gsum[torch.from_numpy(mine.gnode)] = torch.from_numpy(np.ascontiguousarray(v))
    gcnt[torch.from_numpy(mine.gnode)] = 1.0
    dist.all_reduce(gsum)
    dist.all_reduce(gcnt)
    mean = (gsum / gcnt[:, None]).numpy()
    consistent = np.allclose(mean[mine.gnode], v, rtol=0, atol=1e-15)
    flags = torch.tensor([float(ok_in), float(ok_out), float(consistent)])
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("GLOO_HALO world=%d in=%d out=%d consistent=%d shared_nodes=%d" %
              (world, int(flags[0]), int(flags[1]), int(flags[2]), int((gcnt > 1).sum())), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flags.min().item() == 1.0 else 1)


if __name__ == "__main__":
    main()
