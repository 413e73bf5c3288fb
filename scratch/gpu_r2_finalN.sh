#!/bin/bash
# final N-GPU evidence: the multi-rank parity worker on both exchange paths, then the bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
N=${1:-2}
EXTRA=${2:-}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1 || { cat gpurun_out/r2_build.log; exit 1; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for nopeer in 0 1; do
  CGO_NO_PEER=$nopeer timeout 900 $TR --master-port $((29600 + N + 10 * nopeer)) tests/multirank_worker.py > gpurun_out/r2_multirank_n${N}_nopeer${nopeer}.log 2>&1
  echo "multirank N=$N CGO_NO_PEER=$nopeer rc=$?"; grep -E "PEER_MEMORY|MISMATCH|MULTIRANK|SOURCES_SHA|Error|error" gpurun_out/r2_multirank_n${N}_nopeer${nopeer}.log | head -8
done
t0=$(date +%s)
timeout 600 $TR --master-port $((29700 + N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_default_n${N}.json 2> gpurun_out/r2_bench_default_n${N}.err
echo "bench default N=$N rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -c 300 gpurun_out/r2_bench_default_n${N}.err
timeout 600 $TR --master-port $((29800 + N)) bench.py --workload logreg --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_logreg_n${N}.json 2> gpurun_out/r2_bench_logreg_n${N}.err
echo "bench logreg N=$N rc=$?"; tail -c 300 gpurun_out/r2_bench_logreg_n${N}.err
if [ -n "$EXTRA" ]; then
  timeout 300 $TR --master-port $((29900 + N)) bench.py --workload rosenbrock --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_rosenbrock_n${N}.json 2> gpurun_out/r2_bench_rosenbrock_n${N}.err
  echo "bench rosenbrock N=$N rc=$?"; tail -c 300 gpurun_out/r2_bench_rosenbrock_n${N}.err
  timeout 300 $TR --master-port $((30000 + N)) bench.py --workload batched --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_batched_n${N}.json 2> gpurun_out/r2_bench_batched_n${N}.err
  echo "bench batched N=$N rc=$?"; tail -c 300 gpurun_out/r2_bench_batched_n${N}.err
fi
