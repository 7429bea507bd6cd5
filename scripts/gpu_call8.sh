#!/bin/bash
# round 2, call 8: bf16 residual stream (GRAM_FLAG_BF16_RESID = 16384): op test, parity vs oracle, bench A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c8
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "gemm" ) > $O/${tag}_pytest_gemm.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${tag}_pytest_gemm.log
if [ $rc -ne 0 ]; then echo failed > $O/${tag}_done; exit 0; fi
( timeout 900 python scripts/parity_bf16_resid.py 64 ) > $O/${tag}_parity.json 2> $O/${tag}_parity.err
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2; do
  timeout 300 $B > $O/${tag}_ab_fp32res_$rep.json 2> $O/${tag}_ab_fp32res_$rep.err
  timeout 300 $B --flags 16384 > $O/${tag}_ab_bf16res_$rep.json 2> $O/${tag}_ab_bf16res_$rep.err
done
( timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${tag}_pytest.log
echo done > $O/${tag}_done
