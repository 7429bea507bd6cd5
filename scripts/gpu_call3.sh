#!/bin/bash
# GPU session 3: self-attention variants A/B (0 = 4 positions/24 warps, 2 = 4 positions/32 warps (spills), 3 = narrow lanes)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
( timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_live_rows.py -x -q -k "generate or logits or live" ) > $O/c3_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c3_pytest.log
for rep in 1 2; do
  for v in 0 2 3; do
    GRAM_SELF_ATTN_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e > $O/c3_bench_v${v}_$rep.json 2> $O/c3_bench_v${v}_$rep.err
  done
done
GRAM_SELF_ATTN_VARIANT=3 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_live_rows.py tests/test_gpu_item_cache.py -x -q > $O/c3_pytest_v3.log 2>&1
echo "pytest v3 rc=$?" >> $O/c3_pytest_v3.log
echo done > $O/c3_done
