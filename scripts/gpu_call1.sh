#!/bin/bash
# GPU session: live-row compaction -- parity, A/B bench, self-attention ncu capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/c1_smi.txt 2>&1
( time timeout 300 python -m pytest tests/test_gpu_live_rows.py -x -q ) > $O/c1_live.log 2>&1
echo "live rc=$?" >> $O/c1_live.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/c1_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c1_pytest.log
for rep in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --all-rows > $O/c1_bench_allrows_$rep.json 2> $O/c1_bench_allrows_$rep.err
  timeout 300 python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 > $O/c1_bench_live_$rep.json 2> $O/c1_bench_live_$rep.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dec_self_attention64 -s 40 -c 2 -o $O/c1_selfattn -f \
  python bench.py --steps 1 --warmup 1 --batch 944 --no-item-cache --cpu-users 0 --no-e2e > $O/c1_ncu_selfattn.log 2>&1
echo done > $O/c1_done
