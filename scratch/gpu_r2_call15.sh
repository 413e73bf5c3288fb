#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1 || { cat gpurun_out/r2_build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_user_objective.py tests/test_gpu_sparse_ls.py -k "user or quadratic or hessian" -m gpu -q 2>&1 | grep -E "^E |passed|failed|Error" | head -30
for coh in 30 5 10; do
  echo "== coh$coh"
  timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-secondary --steps 20 --warmup 5 --coh $coh | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['config']['repetitions'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['clocks']['sm_mhz'])"
done
echo "== coh5 window 0"
CGO_SWEEP_WINDOW=0 timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-secondary --steps 20 --warmup 5 --coh 5 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['config']['repetitions'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['clocks']['sm_mhz'])"
