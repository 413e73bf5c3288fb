#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
rm -f gpurun_out/north_star_gates.jsonl
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1 || { cat gpurun_out/r2_build.log; exit 1; }
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2_pytest_gpu.log 2>&1 ) 2>&1 | grep real; echo "pytest rc=$?"
tail -16 gpurun_out/r2_pytest_gpu.log
cat gpurun_out/north_star_gates.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
