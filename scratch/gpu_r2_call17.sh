#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1 || { cat gpurun_out/r2_build.log; exit 1; }
for cfg in 3 4 5 6; do
  echo "== logreg direct cfg $cfg"
  CGO_DIRECT_CFG=$cfg timeout 600 python bench.py --workload logreg --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --min-timed-s 1.0 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(d['value'], r['fdf_evals_per_s'], r['frac'], r['avg_launch_ms'])"
done
for cfg in 0 3 6; do echo "== LS coh0 direct cfg $cfg"; CGO_DIRECT_CFG=$cfg timeout 300 python scratch/exp_coh0.py 2e8 0 8; done
