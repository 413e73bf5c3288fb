#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
N=${1:-2}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1 || { cat gpurun_out/r2_build.log; exit 1; }
for nopeer in 0 1; do
  CGO_NO_PEER=$nopeer timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N + 10 * nopeer)) tests/multirank_worker.py > gpurun_out/r2_multirank_n${N}_nopeer${nopeer}.log 2>&1
  echo "multirank N=$N CGO_NO_PEER=$nopeer rc=$?"; grep -E "PEER_MEMORY|MISMATCH|MULTIRANK|Error|error" gpurun_out/r2_multirank_n${N}_nopeer${nopeer}.log | head -8
done
