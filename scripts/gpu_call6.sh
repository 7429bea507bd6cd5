#!/bin/bash
# round 2, call 6: persistent cross-attention (tests, A/B), fp32 bench number
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c6
( time timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "cross_attention" ) > $O/${tag}_xattn.log 2>&1
rc=$?; echo "xattn rc=$rc" >> $O/${tag}_xattn.log
if [ $rc -ne 0 ]; then echo failed > $O/${tag}_done; exit 0; fi
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${tag}_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2; do
  timeout 300 $B > $O/${tag}_ab_persist_$rep.json 2> $O/${tag}_ab_persist_$rep.err
  timeout 300 $B --flags 8192 > $O/${tag}_ab_peritem_$rep.json 2> $O/${tag}_ab_peritem_$rep.err
done
timeout 900 python bench.py --config scale5 --steps 4 --warmup 3 --cpu-users 0 --no-e2e > $O/${tag}_scale5_persist.json 2> $O/${tag}_scale5_persist.err
timeout 900 python bench.py --config scale5 --steps 4 --warmup 3 --cpu-users 0 --no-e2e --flags 8192 > $O/${tag}_scale5_peritem.json 2> $O/${tag}_scale5_peritem.err
timeout 900 python bench.py --config yelp --steps 6 --warmup 3 --cpu-users 0 --no-e2e --no-item-cache > $O/${tag}_yelp_persist.json 2> $O/${tag}_yelp_persist.err
timeout 900 python bench.py --dtype fp32 --batch 472 --steps 4 --warmup 3 --cpu-users 0 --no-item-cache > $O/${tag}_fp32.json 2> $O/${tag}_fp32.err
echo done > $O/${tag}_done
