#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1 || { cat gpurun_out/r2_build.log; exit 1; }
echo "== LS coh0 auto"; timeout 300 python scratch/exp_coh0.py 2e8 0 8
echo "== LS coh0 cfg2"; CGO_DIRECT_CFG=2 timeout 300 python scratch/exp_coh0.py 2e8 0 8
timeout 900 python -m pytest tests/test_gpu_sparse_ls.py tests/test_gpu_logreg.py tests/test_gpu_c_host.py -x -q 2>&1 | tail -3
echo "== logreg auto"
timeout 600 python bench.py --workload logreg --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --min-timed-s 2.0 | tee gpurun_out/r2_bench_logreg_auto.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(d['value'], r['fdf_evals_per_s'], r['frac'], r['avg_launch_ms'])"
