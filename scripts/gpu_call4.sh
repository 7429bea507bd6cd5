#!/bin/bash
# GPU session 4: tcgen05 encoder attention after the softmax instruction diet
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
( timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tcgen05_encoder_attention or generate_bf16" ) > $O/c4_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c4_pytest.log
for rep in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e --tc-enc-attn > $O/c4_bench_tc_$rep.json 2> $O/c4_bench_tc_$rep.err
  timeout 300 python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e > $O/c4_bench_mma_$rep.json 2> $O/c4_bench_mma_$rep.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:enc_attention_tc -s 2 -c 1 -o $O/c4_encattn_tc -f \
  python bench.py --steps 1 --warmup 1 --batch 944 --no-item-cache --cpu-users 0 --no-e2e --tc-enc-attn > $O/c4_ncu.log 2>&1
echo done > $O/c4_done
