#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
rm -f gpurun_out/north_star_gates.jsonl
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1 || { cat gpurun_out/r2_build.log; exit 1; }
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
t0=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"
timeout 900 python bench.py --workload logreg --steps 20 --warmup 3 > gpurun_out/r2_bench_logreg_1gpu.json 2> gpurun_out/r2_bench_logreg_1gpu.err; echo "logreg rc=$?"
timeout 900 python bench.py --workload rosenbrock --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_rosenbrock.json 2> gpurun_out/r2_bench_rosenbrock.err; echo "rosen rc=$?"
