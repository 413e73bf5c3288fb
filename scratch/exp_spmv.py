import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cgoptim_b200 as cg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
coh = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ctx = cg.Context(0)
ctx.set_reduction_ctas(296)
obj = cg.SparseLSGPU(n, 10, None, 24, coh, ctx)
x = np.ones(n)
ctx.timing(True)
for rep in range(3):
    ctx.timing_read(reset=True)
    obj.spmv(x)
    obj.spmv(x, transposed=True)
    t = ctx.timing_read(reset=True)
    print("EpiStore  A: %.3f ms   AT: %.3f ms" % (t["spmv"][0], t["spmvT"][0]))
ws = obj.make_workspace(np.zeros(n), fuse_direction=False)
ws.reset_direction()
for rep in range(3):
    ctx.timing_read(reset=True)
    ws.eval_trial(0.01)
    t = ctx.timing_read(reset=True)
    print("trial  K_b: %.3f ms   K_c: %.3f ms  K_a: %.3f" % (t["spmv"][0], t["spmvT"][0], t["axpy"][0]))
