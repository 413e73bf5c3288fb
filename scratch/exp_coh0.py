"""coh-0 experiment (r2): K_b / K_c launch times of the cfg-3 matrix under the lockstep window of the CSR sweep."""
import json
import sys
import os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cgoptim_b200 as cg

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
coh = int(sys.argv[2]) if len(sys.argv) > 2 else 0
windows = [int(w) for w in sys.argv[3].split(",")] if len(sys.argv) > 3 else [8, 0, 4, 16]
ctx = cg.Context(0)
obj = cg.SparseLSGPU(n, 10, None, 24, coh, ctx)
ws = obj.make_workspace(np.zeros(n))
ws.reset_direction()
M = 12.0 * 10 * n + 8.0 * (n + 1)
print(json.dumps({"n": n, "coh": coh}), flush=True)
for w in windows:
    ws.eval_trial(1e-3)      # warm
    ctx.timing(True); ctx.timing_read(reset=True)
    for i in range(3):
        ws.eval_trial(1e-3 * (i + 2))
    t = ctx.timing_read(reset=True)
    ctx.timing(False)
    kb, kc = t["spmv"][0] / max(t["spmv"][1], 1), t["spmvT"][0] / max(t["spmvT"][1], 1)
    print(json.dumps({"window": w, "K_b_ms": round(kb, 3), "K_c_ms": round(kc, 3),
                      "K_a_ms": round(t["axpy"][0] / max(t["axpy"][1], 1), 3),
                      "K_b_GBs": round((M + 24.0 * n) / kb * 1e-6, 1), "K_c_GBs": round((M + 32.0 * n) / kc * 1e-6, 1)}), flush=True)
