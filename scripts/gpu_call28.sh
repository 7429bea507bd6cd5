#!/bin/bash
# round 2, call 28: chain kernel with pipelined TMEM reads: chain tests (bit identity with three launches), bench x3 vs --flags 1024
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c28
( timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py tests/test_gpu_live_rows.py -q -x ) > $O/${tag}_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${tag}_pytest.log
if [ $rc -ne 0 ]; then echo failed > $O/${tag}_done; exit 0; fi
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2 3; do
  timeout 300 $B > $O/${tag}_chain_$rep.json 2> $O/${tag}_chain_$rep.err
  timeout 300 $B --flags 1024 > $O/${tag}_nochain_$rep.json 2> $O/${tag}_nochain_$rep.err
done
echo done > $O/${tag}_done
