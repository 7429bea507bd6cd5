#!/bin/bash
# round 2, call 10: encoder attention with dead-warp skip / passage-wide S and PV: suite, bench, captures of both attention kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c10
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/${tag}_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${tag}_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2; do
  timeout 300 $B > $O/${tag}_default_$rep.json 2> $O/${tag}_default_$rep.err
done
timeout 900 python bench.py --config scale5 --steps 4 --warmup 3 --cpu-users 0 --no-e2e > $O/${tag}_cfg_scale5.json 2> $O/${tag}_cfg_scale5.err
P="python bench.py --steps 1 --warmup 1 --no-item-cache --cpu-users 0 --no-e2e"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cross_attention -s 20 -c 2 -o $O/prof_xattn_r2b -f $P > $O/${tag}_ncu_xattn.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:enc_attention_tc -s 2 -c 1 -o $O/prof_encattn_r2b -f $P > $O/${tag}_ncu_encattn.log 2>&1
echo done > $O/${tag}_done
