#!/bin/bash
# round 2, call 35: ncu --set full captures of the vocabulary head and the encoder attention of the shipped binary (tag r2d)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c35
P="python bench.py --steps 1 --warmup 1 --no-item-cache --cpu-users 0 --no-e2e"
timeout 300 $P > $O/${tag}_plain.json 2> $O/${tag}_plain.err || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel<4" -s 3 -c 1 -o $O/prof_lmhead_r2d -f $P > $O/${tag}_ncu_head.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:enc_attention_tc -s 2 -c 1 -o $O/prof_encattn_r2d -f $P > $O/${tag}_ncu_encattn.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 0 -c 5 -o $O/prof_gemm_enc_r2d -f $P > $O/${tag}_ncu_gemm.log 2>&1
echo done > $O/${tag}_done
