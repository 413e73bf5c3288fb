#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
N=${1:-8}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1 || { cat gpurun_out/r2_build.log; exit 1; }
for nopeer in 0 1; do
  CGO_NO_PEER=$nopeer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N + 10 * nopeer)) tests/multirank_worker.py > gpurun_out/r2_multirank_n${N}_nopeer${nopeer}.log 2>&1
  echo "multirank N=$N CGO_NO_PEER=$nopeer rc=$?"; grep -E "PEER_MEMORY|MISMATCH|MULTIRANK|Error|error" gpurun_out/r2_multirank_n${N}_nopeer${nopeer}.log | head -8
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_default_n${N}.json 2> gpurun_out/r2_bench_default_n${N}.err
echo "bench default N=$N rc=$?"; tail -c 300 gpurun_out/r2_bench_default_n${N}.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800 + N)) bench.py --workload logreg --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_logreg_n${N}.json 2> gpurun_out/r2_bench_logreg_n${N}.err
echo "bench logreg N=$N rc=$?"; tail -c 300 gpurun_out/r2_bench_logreg_n${N}.err
