#!/bin/bash
# quick check of the in-tree library on a fresh box: smoke + the parity tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/verify_smoke.log 2>&1
echo "smoke rc=$?" >> $O/verify_smoke.log
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bf16_path.py -q ) > $O/verify_pytest.log 2>&1
echo "pytest rc=$?" >> $O/verify_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-users 0 --no-item-cache > $O/verify_bench.json 2> $O/verify_bench.err
echo done > $O/verify_done
