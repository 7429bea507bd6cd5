#!/bin/bash
# round 2, call 29: EPI_RESID_NORM with pipelined TMEM reads (decoder o-projection; fp32-stream encoder): tests, bench x3, fp32-stream bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c29
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chain.py tests/test_gpu_item_cache.py -q -x ) > $O/${tag}_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${tag}_pytest.log
if [ $rc -ne 0 ]; then echo failed > $O/${tag}_done; exit 0; fi
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2 3; do
  timeout 300 $B > $O/${tag}_default_$rep.json 2> $O/${tag}_default_$rep.err
done
timeout 300 $B --flags 16384 > $O/${tag}_fp32res.json 2> $O/${tag}_fp32res.err
echo done > $O/${tag}_done
