#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1 || { cat gpurun_out/r2_build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_gpu.log
A="--no-e2e --no-cpu-baseline --steps 10 --warmup 3 --coh 0"
timeout 600 python bench.py $A > gpurun_out/r2_bench_coh0_direct.json 2> gpurun_out/r2_bench_coh0_direct.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_direct -s 8 -c 2 -f -o gpurun_out/prof_ls_coh0_direct_r2 python bench.py $A > gpurun_out/r2_ncu_coh0_direct.log 2>&1
cat gpurun_out/r2_bench_coh0_direct.json
