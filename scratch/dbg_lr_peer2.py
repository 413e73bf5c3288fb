import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cgoptim_b200 as cg
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ctx = cg.Context(rank); ctx.comm_init_torch()
tag = "peer" if ctx.peer_memory else "nccl"
N, d, lam = 20_000, 2002, 1e-4
ctx.set_gather_block_bytes(0)
obj = cg.LogRegGPU(N, d, 20, 24, lam, ctx)
res = {}
for fuse in (True, False):
    ws = obj.make_workspace(np.zeros(obj.n_local), fuse_direction=fuse)
    ws.reset_direction()
    ws.eval_trial(100.0); res[f"{fuse}_p1"] = ws.pack.copy()
    ws.accept()
    ws.update_dir(0.3)
    ws.hint_first_trial(50.0)
    res[f"{fuse}_gu"] = np.array([ws.dot_g_u(), ws.dot_u_u()])
    ws.eval_trial(50.0); res[f"{fuse}_p2"] = ws.pack.copy()
    for nm in ("x", "df_x", "u", "xp", "df_xp"):
        res[f"{fuse}_{nm}"] = ws.download_vector(nm)
    ws.close()
np.savez(os.path.join(ROOT, "gpurun_out", f"dbg2_{tag}_{rank}.npz"), **res)
dist.barrier()
dist.destroy_process_group()
