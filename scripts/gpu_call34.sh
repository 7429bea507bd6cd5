#!/bin/bash
# round 2, call 34: independent max / sum chains in the log-softmax epilogue and in the encoder-attention softmax: tests, bench x3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c34
( timeout 900 python -m pytest tests/test_gpu_bf16_path.py tests/test_gpu_parity.py tests/test_gpu_item_cache.py -q -x ) > $O/${tag}_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${tag}_pytest.log
if [ $rc -ne 0 ]; then echo failed > $O/${tag}_done; exit 0; fi
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2 3; do
  timeout 300 $B > $O/${tag}_new_$rep.json 2> $O/${tag}_new_$rep.err
done
timeout 900 python bench.py --config scale5 --steps 4 --warmup 3 --cpu-users 0 --no-e2e > $O/${tag}_scale5.json 2> $O/${tag}_scale5.err
echo done > $O/${tag}_done
