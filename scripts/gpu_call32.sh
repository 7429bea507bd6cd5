#!/bin/bash
# round 2, call 32: log-softmax epilogue with one FFMA + MUFU + FADD per logit and no column select outside the last tile: tests, bench x3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c32
( timeout 900 python -m pytest tests/test_gpu_bf16_path.py tests/test_gpu_parity.py -q -x ) > $O/${tag}_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${tag}_pytest.log
if [ $rc -ne 0 ]; then echo failed > $O/${tag}_done; exit 0; fi
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2 3; do
  timeout 300 $B > $O/${tag}_new_$rep.json 2> $O/${tag}_new_$rep.err
done
echo done > $O/${tag}_done
