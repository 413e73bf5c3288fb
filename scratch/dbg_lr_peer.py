"""debug: sample-sharded logreg, peer vs NCCL exchange paths, blocked / unblocked, HZ / LBFGS"""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cgoptim_b200 as cg
from helpers import make_pair
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ctx = cg.Context(rank); ctx.comm_init_torch()
tag = "peer" if ctx.peer_memory else "nccl"
N, d, lam = 20_000, 2002, 1e-4
out = {}
for flavour in ("LBFGS", "HagerZhang"):
    for blk in (0, 8 * 600):
        for rep in range(2):
            ctx.set_gather_block_bytes(blk)
            _, cfg, ls = make_pair(flavour, eps=1e-6, max_iters=60, c1=1e-4, c2=0.9, lbfgs_m=10)
            obj = cg.LogRegGPU(N, d, 20, 24, lam, ctx)
            ret = cg.minimizeobjective(obj, np.zeros(obj.n_local), cfg, ls)
            out[f"{flavour}_{blk}_{rep}"] = ret.trace.objective.copy()
            obj.close()
if rank == 0:
    np.savez(os.path.join(ROOT, "gpurun_out", f"dbg_{tag}.npz"), **out)
    for k, v in out.items():
        print(tag, k, len(v), repr(v[-1]))
dist.destroy_process_group()
