#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1 || { cat gpurun_out/r2_build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_c_host.py -x -q > gpurun_out/r2_pytest_chost.log 2>&1; echo "c_host pytest rc=$?"; tail -3 gpurun_out/r2_pytest_chost.log
nproc; free -g | head -2
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --write-fixture > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err ) 2>&1 | tail -3
cp tests/golden/bench_trace_n1.json gpurun_out/bench_trace_n1.json
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err ) 2>&1 | tail -3
tail -c 600 gpurun_out/r2_bench_default.err
