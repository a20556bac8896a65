"""diagnostic: c3_plate_mixed_4M block-CSR flavour at N ranks: kernel-class times of ElmGMRs, iteration counts of both solves"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from phasta_b200 import SolverParams, make_tables
from phasta_b200.solver import PhastaGPU, nccl_unique_id

world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
name = sys.argv[1] if len(sys.argv) > 1 else "c3_plate_mixed_4M"
params = SolverParams(ibksiz=1024, etol=1e-3, Kspace=50)
tables = make_tables(2, 2)
part, y, ac = bench.build_part(name, rank, world)
g = PhastaGPU(part, params, tables, device=lr)
if world > 1:
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    g.comm_init(bytes(idt.cpu().tolist()))
g.set_state(y, ac)
st = g.step()
g.dev_elmgmre(st)
it_e = g.dev_solve(st)
dy_e = g.get("Dy")
g.genadj()
for rep in range(3):
    g.profile(True); g.profile_reset()
    t0 = time.perf_counter()
    g.dev_elmgmrs(st)
    g.sync()
    dt = time.perf_counter() - t0
    pk = g.profile_get(); g.profile(False)
    if rank == 0:
        print("rep %d ElmGMRs wall %.2f ms classes %s" % (rep, dt * 1e3, {k: round(v[0], 3) for k, v in pk.items()}), flush=True)
t0 = time.perf_counter(); g.dev_elmgmrs(st); g.sync()
if rank == 0: print("unprofiled ElmGMRs wall %.2f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
it_s = g.dev_solve_sparse(st)
dy_s = g.get("Dy")
rel = np.linalg.norm(dy_e - dy_s) / np.linalg.norm(dy_e)
print("rank %d: SolGMRe %d its, SolGMRs %d its, rel diff of Dy %.2e (both at etol 1e-3)" % (rank, it_e, it_s, rel), flush=True)
g.close()
if world > 1:
    dist.destroy_process_group()
