#!/bin/bash
# GPU session 6: persistent warp-specialised tcgen05 encoder attention
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
( timeout 180 python -m pytest tests/test_gpu_parity.py -x -q -k "tcgen05_encoder_attention" ) > $O/c6_pytest.log 2>&1
rc=$?
echo "pytest rc=$rc" >> $O/c6_pytest.log
if [ $rc -eq 0 ]; then
  for rep in 1 2; do
    timeout 200 python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e --tc-enc-attn > $O/c6_bench_tc_$rep.json 2> $O/c6_bench_tc_$rep.err
    timeout 200 python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e > $O/c6_bench_mma_$rep.json 2> $O/c6_bench_mma_$rep.err
  done
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:enc_attention_tc -s 2 -c 1 -o $O/c6_encattn_tc -f \
    python bench.py --steps 1 --warmup 1 --batch 944 --no-item-cache --cpu-users 0 --no-e2e --tc-enc-attn > $O/c6_ncu.log 2>&1
fi
echo done > $O/c6_done
