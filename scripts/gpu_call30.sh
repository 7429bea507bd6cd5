#!/bin/bash
# round 2, call 30: transposed cross-attention with a mask-free copy of the tile for fully visible tiles (GRAM_XATTN_NOMASK=0 = off)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c30
( timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "cross_attention" ) > $O/${tag}_pytest_op.log 2>&1
echo "pytest rc=$?" >> $O/${tag}_pytest_op.log
timeout 200 python scripts/exp_xattn_hot.py > $O/${tag}_alone_fast.log 2>&1
GRAM_XATTN_NOMASK=0 timeout 200 python scripts/exp_xattn_hot.py > $O/${tag}_alone_mask.log 2>&1
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2 3; do
  timeout 300 $B > $O/${tag}_fast_$rep.json 2> $O/${tag}_fast_$rep.err
  GRAM_XATTN_NOMASK=0 timeout 300 $B > $O/${tag}_mask_$rep.json 2> $O/${tag}_mask_$rep.err
done
echo done > $O/${tag}_done
