#!/bin/bash
# GPU session 2: self-attention variants A/B, full test suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/c2_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c2_pytest.log
for rep in 1 2; do
  GRAM_SELF_ATTN_UN=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 > $O/c2_bench_un4_$rep.json 2> $O/c2_bench_un4_$rep.err
  GRAM_SELF_ATTN_UN=8 timeout 300 python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 > $O/c2_bench_un8_$rep.json 2> $O/c2_bench_un8_$rep.err
done
( time timeout 600 python bench.py ) > $O/c2_bench_default.json 2> $O/c2_bench_default.err
echo done > $O/c2_done
