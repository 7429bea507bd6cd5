#!/bin/bash
# round 2, call 17: encoder attention with an L2 prefetch iterator (GRAM_ATTN_PREFETCH=0 = off): tests, A/B, scale5
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c17
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bf16_path.py tests/test_gpu_item_cache.py -q -x ) > $O/${tag}_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${tag}_pytest.log
if [ $rc -ne 0 ]; then echo failed > $O/${tag}_done; exit 0; fi
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2 3; do
  timeout 300 $B > $O/${tag}_pf1_$rep.json 2> $O/${tag}_pf1_$rep.err
  GRAM_ATTN_PREFETCH=0 timeout 300 $B > $O/${tag}_pf0_$rep.json 2> $O/${tag}_pf0_$rep.err
done
S="python bench.py --config scale5 --steps 4 --warmup 3 --cpu-users 0 --no-e2e"
timeout 900 $S > $O/${tag}_scale5_pf1.json 2> $O/${tag}_scale5_pf1.err
GRAM_ATTN_PREFETCH=0 timeout 900 $S > $O/${tag}_scale5_pf0.json 2> $O/${tag}_scale5_pf0.err
echo done > $O/${tag}_done
