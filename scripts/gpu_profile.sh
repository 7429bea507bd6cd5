#!/bin/bash
# Profiles of the shipped binary on one B200 (run AFTER bench.py has exited 0 without ncu):
#   launch list of one whole step + ncu --set full captures of the top kernels.  scripts/summarize_profiles.py <tag>
#   turns gpurun_out/launches_<tag>.csv and gpurun_out/prof_*_<tag>.ncu-rep into profiles/<tag>_*.
tag=${1:-r1}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --no-item-cache --cpu-users 0 --no-e2e"
timeout 300 $B > $O/prof_plain_$tag.json 2> $O/prof_plain_$tag.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_$tag.csv $B > $O/prof_launches_$tag.log 2>&1
cap() {  # name, kernel regex, skip, count
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o $O/prof_$1_$tag -f $B > $O/prof_$1_$tag.log 2>&1
}
cap xattn cross_attention_mma 20 2
cap gemm_enc gemm_tc_kernel 0 5
cap encattn enc_attention_tc 2 1
cap selfattn dec_self_attention64 40 2
echo done > $O/prof_done_$tag
