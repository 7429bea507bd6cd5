#!/bin/bash
# round 2, call 11: pipelined LSE epilogue (tests + A/B of CTA pairs for the head), decoder chain A/B, users-per-step sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c11
( timeout 900 python -m pytest tests/test_gpu_bf16_path.py tests/test_gpu_parity.py -q -x ) > $O/${tag}_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${tag}_pytest.log
if [ $rc -ne 0 ]; then echo failed > $O/${tag}_done; exit 0; fi
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2; do
  timeout 300 $B > $O/${tag}_default_$rep.json 2> $O/${tag}_default_$rep.err
  GRAM_LSE_PAIRS=1 timeout 300 $B > $O/${tag}_lsepairs_$rep.json 2> $O/${tag}_lsepairs_$rep.err
  timeout 300 $B --flags 1024 > $O/${tag}_nodecchain_$rep.json 2> $O/${tag}_nodecchain_$rep.err
done
for b in 2832 3776; do
  timeout 400 $B --batch $b --steps 6 > $O/${tag}_batch_$b.json 2> $O/${tag}_batch_$b.err
done
timeout 300 $B --batch 1416 > $O/${tag}_batch_1416.json 2> $O/${tag}_batch_1416.err
echo done > $O/${tag}_done
