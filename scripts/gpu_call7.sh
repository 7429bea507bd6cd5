#!/bin/bash
# round 2, call 7: two-group GEMM epilogue (GRAM_GEMM_EG=1 = one group): op-level check + timing, tests, bench A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c7
( timeout 600 python scripts/check_gemm_tc.py --time ) > $O/${tag}_gemm_eg2.log 2>&1
echo "rc=$?" >> $O/${tag}_gemm_eg2.log
( GRAM_GEMM_EG=1 timeout 600 python scripts/check_gemm_tc.py --time ) > $O/${tag}_gemm_eg1.log 2>&1
echo "rc=$?" >> $O/${tag}_gemm_eg1.log
if ! grep -q "ALL OK" $O/${tag}_gemm_eg2.log; then echo failed > $O/${tag}_done; exit 0; fi
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chain.py tests/test_gpu_bf16_path.py -q -x ) > $O/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${tag}_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2; do
  timeout 300 $B > $O/${tag}_ab_eg2_$rep.json 2> $O/${tag}_ab_eg2_$rep.err
  GRAM_GEMM_EG=1 timeout 300 $B > $O/${tag}_ab_eg1_$rep.json 2> $O/${tag}_ab_eg1_$rep.err
done
echo done > $O/${tag}_done
