#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1 || { cat gpurun_out/r2_build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --workload logreg --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_logreg_1gpu.json 2> gpurun_out/r2_bench_logreg_1gpu.err; echo "bench logreg rc=$?"
tail -c 300 gpurun_out/r2_bench_logreg_1gpu.err
