#!/bin/bash
# round 2, call 19: decoder self-attention with row-major warp mapping (GRAM_SELF_ATTN_ROWMAJ=1), users-per-step 1,888 vs 2,832, three reps each
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c19
( GRAM_SELF_ATTN_ROWMAJ=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_live_rows.py tests/test_gpu_item_cache.py -q -x ) > $O/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${tag}_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2 3; do
  timeout 300 $B > $O/${tag}_base_$rep.json 2> $O/${tag}_base_$rep.err
  GRAM_SELF_ATTN_ROWMAJ=1 timeout 300 $B > $O/${tag}_rowmaj_$rep.json 2> $O/${tag}_rowmaj_$rep.err
  timeout 400 $B --batch 2832 --steps 7 > $O/${tag}_b2832_$rep.json 2> $O/${tag}_b2832_$rep.err
done
echo done > $O/${tag}_done
