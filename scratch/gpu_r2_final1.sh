#!/bin/bash
# final 1-GPU evidence of round 2: full GPU test suite, smoke, the driver's bench command, the reference arm,
# the other workloads, then the ncu launch list of the bench command (after its plain run exited 0)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
rm -f gpurun_out/north_star_gates.jsonl
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1 || { cat gpurun_out/r2_build.log; exit 1; }
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
t0=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"
t0=$(date +%s)
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference rc=$? wall=$(( $(date +%s) - t0 ))s"
timeout 900 python bench.py --workload rosenbrock --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_rosenbrock.json 2> gpurun_out/r2_bench_rosenbrock.err; echo "rosen rc=$?"
timeout 900 python bench.py --workload batched --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_batched.json 2> gpurun_out/r2_bench_batched.err; echo "batched rc=$?"
timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-secondary --steps 3 --warmup 3 --min-timed-s 0 > gpurun_out/r2_bench_plain_for_ncu.json 2>&1; rc=$?; echo "plain-for-ncu rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_sparse_ls_n2e8_coh0.csv python bench.py --no-e2e --no-cpu-baseline --no-secondary --steps 3 --warmup 3 --min-timed-s 0 > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu rc=$?"
fi
