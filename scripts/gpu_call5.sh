#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c5
( time timeout 600 python -m pytest tests/test_gpu_graph.py -q -x ) > $O/${tag}_graph.log 2>&1
echo "graph rc=$?" >> $O/${tag}_graph.log
timeout 900 python scripts/latency_small_batch.py > $O/${tag}_latency.log 2>&1
echo done > $O/${tag}_done
