#!/bin/bash
# round 2, call 12: cross-attention with 8 consumer warps (GRAM_XATTN_WARPS=4 = the 4-warp variant): tests, A/B on beauty and scale5
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c12
( timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/${tag}_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${tag}_pytest.log
if [ $rc -ne 0 ]; then echo failed > $O/${tag}_done; exit 0; fi
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2; do
  timeout 300 $B > $O/${tag}_w8_$rep.json 2> $O/${tag}_w8_$rep.err
  GRAM_XATTN_WARPS=4 timeout 300 $B > $O/${tag}_w4_$rep.json 2> $O/${tag}_w4_$rep.err
done
S="python bench.py --config scale5 --steps 4 --warmup 3 --cpu-users 0 --no-e2e"
timeout 900 $S > $O/${tag}_scale5_w8.json 2> $O/${tag}_scale5_w8.err
GRAM_XATTN_WARPS=4 timeout 900 $S > $O/${tag}_scale5_w4.json 2> $O/${tag}_scale5_w4.err
timeout 600 python bench.py --config yelp --steps 6 --warmup 3 --cpu-users 0 --no-e2e --no-item-cache > $O/${tag}_yelp_w8.json 2> $O/${tag}_yelp_w8.err
echo done > $O/${tag}_done
