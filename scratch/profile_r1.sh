#!/bin/bash
# r1 profiling pass (one gpurun call, one GPU): launch lists + one `ncu --set full` capture per kernel family
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
A="--no-e2e --no-cpu-baseline --steps 3 --warmup 3"
python bench.py $A > gpurun_out/plain_ls.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ls.csv python bench.py $A > gpurun_out/ncu_ls_launch.log 2>&1
python bench.py $A > gpurun_out/plain_ls2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_csr_rows -s 8 -c 2 -f -o gpurun_out/prof_ls_r1b python bench.py $A > gpurun_out/ncu_ls_full.log 2>&1
python bench.py --workload logreg $A > gpurun_out/plain_lr.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_csr_rows -s 40 -c 3 -f -o gpurun_out/prof_logreg_r1 python bench.py --workload logreg $A > gpurun_out/ncu_lr_full.log 2>&1
python bench.py --workload rosenbrock $A > gpurun_out/plain_ro.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_blas1 -s 4 -c 2 -f -o gpurun_out/prof_rosen_r1 python bench.py --workload rosenbrock $A > gpurun_out/ncu_ro_full.log 2>&1
python bench.py --workload batched $A > gpurun_out/plain_ba.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_batched -s 1 -c 1 -f -o gpurun_out/prof_batched_r1 python bench.py --workload batched $A > gpurun_out/ncu_ba_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -n 2 gpurun_out/ncu_*_full.log
