#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1 || { cat gpurun_out/r2_build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_user_objective.py tests/test_gpu_sparse_ls.py -k "user or quadratic or hessian" -m gpu -x -q 2>&1 | tail -4
B="--no-e2e --no-cpu-baseline --no-secondary --steps 20 --warmup 5 --coh 30"
for w in 8 0; do for t in 2.0 0.2; do
  echo "== coh30 window $w min-timed $t"
  CGO_SWEEP_WINDOW=$w timeout 300 python bench.py $B --min-timed-s $t | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['config']['repetitions'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
echo "== logreg timers"
timeout 600 python - <<'PY'
import numpy as np, cgoptim_b200 as cg, sys
sys.argv=['x']
import bench
ctx = cg.Context(0)
n=20_000_000
obj = cg.LogRegGPU(int(n*2.5), n, 20, 24, 1e-6, ctx)
cfg, ls = bench.solver_configs(cg, 30, "logreg")
run = cg.MinimizerRun(obj, np.zeros(n), cfg, ls)
for _ in range(3): run.step()
ctx.timing(True); ctx.timing_read(reset=True)
ev=0
import time
t0=time.perf_counter(); l0=ctx.kernel_launches
for _ in range(5):
    run.step(); ev += int(run.fdf_evals_ran)
import torch; torch.cuda.synchronize()
dt=time.perf_counter()-t0
t=ctx.timing_read(reset=True)
print("5 iterations", round(dt*1e3,1), "ms; evals", ev, "launches", ctx.kernel_launches-l0)
print({k:(round(v[0],1), v[1]) for k,v in t.items() if v[1]})
PY
