#!/bin/bash
# one ncu --set full capture per call, after the same command ran plain with exit 0:  $1 = ls | logreg
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1 || { cat gpurun_out/r2_build.log; exit 1; }
if [ "$1" = "logreg" ]; then
  A="--workload logreg --no-e2e --no-cpu-baseline --steps 3 --warmup 3 --min-timed-s 0"; SKIP=60; CNT=14; OUT=prof_logreg_final_r2
else
  A="--no-e2e --no-cpu-baseline --no-secondary --steps 6 --warmup 3 --min-timed-s 0"; SKIP=8; CNT=2; OUT=prof_ls_coh0_final_r2
fi
timeout 600 python bench.py $A > gpurun_out/r2_ncu_plain_$1.json 2> gpurun_out/r2_ncu_plain_$1.err; rc=$?; echo "plain rc=$rc"
[ $rc -eq 0 ] || { tail -5 gpurun_out/r2_ncu_plain_$1.err; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_spmv_direct -s $SKIP -c $CNT -f -o gpurun_out/$OUT python bench.py $A > gpurun_out/r2_ncu_$1.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/$OUT.ncu-rep
