"""Where the end-to-end call spends its time (cfg 3, n = 2e8): raw pinned H2D / D2H rate, then minimizeobjective
with K = 1 and K = 20 iterations from a pinned x0."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import cgoptim_b200 as cg
from cgoptim_b200 import _capi as capi
import bench

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
out = {"n": n}
h = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device="cuda")
for name, (a, b) in {"h2d": (d, h), "d2h": (h, d)}.items():
    a.copy_(b, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    out[name + "_GBs"] = round(3 * 8 * n / (time.perf_counter() - t0) / 1e9, 2)
p = capi.pinned_empty(n)
pt = torch.from_numpy(p)
torch.cuda.synchronize(); t0 = time.perf_counter(); pt.copy_(d, non_blocking=True); torch.cuda.synchronize()
out["d2h_GBs_cgo_pinned"] = round(8 * n / (time.perf_counter() - t0) / 1e9, 2)
del h, d, p, pt
ctx = cg.default_context()
obj = cg.SparseLSGPU(n, 10, None, 24, 0, ctx)
x0 = np.zeros(n)
x0p = torch.from_numpy(x0).pin_memory().numpy()
for K in (1, 1, 20, 20, 1):
    cfg, ls = bench.solver_configs(cg, K, "sparse_ls")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ret = cg.minimizeobjective(obj, x0p, cfg, ls)
    torch.cuda.synchronize()
    out.setdefault("call_ms_K%d" % K, []).append(round(1e3 * (time.perf_counter() - t0), 1))
print(json.dumps(out))
