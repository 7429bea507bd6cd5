#!/bin/bash
# round 2, call 16: two epilogue groups for the vocabulary head (GRAM_LSE_EG=2): tests under both settings, A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c16
( GRAM_LSE_EG=2 timeout 900 python -m pytest tests/test_gpu_bf16_path.py tests/test_gpu_parity.py -q -x ) > $O/${tag}_pytest_eg2.log 2>&1
echo "pytest rc=$?" >> $O/${tag}_pytest_eg2.log
( timeout 600 python -m pytest tests/test_gpu_bf16_path.py tests/test_gpu_parity.py -q -x -k "lse or head or gemm" ) > $O/${tag}_pytest_eg1.log 2>&1
echo "pytest rc=$?" >> $O/${tag}_pytest_eg1.log
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2 3; do
  timeout 300 $B > $O/${tag}_eg1_$rep.json 2> $O/${tag}_eg1_$rep.err
  GRAM_LSE_EG=2 timeout 300 $B > $O/${tag}_eg2_$rep.json 2> $O/${tag}_eg2_$rep.err
done
echo done > $O/${tag}_done
