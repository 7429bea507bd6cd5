#!/bin/bash
# round 2, call 3: suite with the folded/chained decoder, A/B of the chain variants, the other BASELINE configs, ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c3
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/${tag}_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${tag}_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2; do
  timeout 300 $B > $O/${tag}_ab_default_$rep.json 2> $O/${tag}_ab_default_$rep.err
  timeout 300 $B --flags 256 > $O/${tag}_ab_encchain_$rep.json 2> $O/${tag}_ab_encchain_$rep.err
  timeout 300 $B --flags 1024 > $O/${tag}_ab_nodecchain_$rep.json 2> $O/${tag}_ab_nodecchain_$rep.err
  timeout 300 $B --unfused-norm > $O/${tag}_ab_unfused_$rep.json 2> $O/${tag}_ab_unfused_$rep.err
done
if [ $rc -eq 0 ]; then
for cfg in toys sports yelp; do
  timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --cpu-users 4 > $O/${tag}_cfg_$cfg.json 2> $O/${tag}_cfg_$cfg.err
done
timeout 900 python bench.py --config scale5 --steps 4 --warmup 3 --cpu-users 0 > $O/${tag}_cfg_scale5.json 2> $O/${tag}_cfg_scale5.err
fi
P="python bench.py --steps 1 --warmup 1 --no-item-cache --cpu-users 0 --no-e2e"
timeout 300 $P > $O/${tag}_prof_plain.json 2> $O/${tag}_prof_plain.err && {
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_r2.csv $P > $O/${tag}_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 0 -c 5 -o $O/prof_gemm_enc_r2 -f $P > $O/${tag}_ncu_gemm.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cross_attention_mma -s 20 -c 2 -o $O/prof_xattn_r2 -f $P > $O/${tag}_ncu_xattn.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:enc_chain_kernel -s 12 -c 2 -o $O/prof_decchain_r2 -f $P > $O/${tag}_ncu_decchain.log 2>&1
}
echo done > $O/${tag}_done
