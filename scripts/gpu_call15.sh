#!/bin/bash
# round 2, call 15: A/B of CTA pairs for the vocabulary head and of the pair threshold (decoder GEMMs on pairs), new pair-shape op test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c15
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "gemm" ) > $O/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${tag}_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2; do
  timeout 300 $B > $O/${tag}_default_$rep.json 2> $O/${tag}_default_$rep.err
  GRAM_LSE_PAIRS=1 timeout 300 $B > $O/${tag}_lsepairs_$rep.json 2> $O/${tag}_lsepairs_$rep.err
  GRAM_PAIR_TILES=8 timeout 300 $B > $O/${tag}_pt8_$rep.json 2> $O/${tag}_pt8_$rep.err
  GRAM_PAIR_TILES=4 timeout 300 $B > $O/${tag}_pt4_$rep.json 2> $O/${tag}_pt4_$rep.err
done
echo done > $O/${tag}_done
