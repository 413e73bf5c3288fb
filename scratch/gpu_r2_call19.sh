#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for mib in 32 40 48 56 64 80; do
  echo "== logreg gather block $mib MiB"
  timeout 600 python bench.py --workload logreg --gather-block-mib $mib --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --min-timed-s 1.0 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(d['value'], r['fdf_evals_per_s'], r['frac'], r['avg_launch_ms'], r.get('timers_ms'))"
done
