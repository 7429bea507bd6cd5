#!/bin/bash
# round 2, call 9: bf16 residual stream as the default: full suite, parity report, bench A/B vs GRAM_FLAG_FP32_RESID, other configs,
# launch list + ncu captures of the encoder-layer GEMMs (tag r2b)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c9
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/${tag}_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${tag}_pytest.log
( timeout 900 python scripts/parity_bf16_resid.py 64 ) > $O/${tag}_parity.json 2> $O/${tag}_parity.err
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2; do
  timeout 300 $B > $O/${tag}_ab_default_$rep.json 2> $O/${tag}_ab_default_$rep.err
  timeout 300 $B --flags 16384 > $O/${tag}_ab_fp32res_$rep.json 2> $O/${tag}_ab_fp32res_$rep.err
done
timeout 600 python bench.py > $O/${tag}_bench.json 2> $O/${tag}_bench.err
if [ $rc -eq 0 ]; then
for cfg in toys sports yelp; do
  timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --cpu-users 0 > $O/${tag}_cfg_$cfg.json 2> $O/${tag}_cfg_$cfg.err
done
timeout 900 python bench.py --config scale5 --steps 4 --warmup 3 --cpu-users 0 > $O/${tag}_cfg_scale5.json 2> $O/${tag}_cfg_scale5.err
fi
P="python bench.py --steps 1 --warmup 1 --no-item-cache --cpu-users 0 --no-e2e"
timeout 300 $P > $O/prof_plain_r2b.json 2> $O/prof_plain_r2b.err && {
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_r2b.csv $P > $O/${tag}_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 0 -c 5 -o $O/prof_gemm_enc_r2b -f $P > $O/${tag}_ncu_gemm.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cross_attention_mma -s 20 -c 2 -o $O/prof_xattn_r2b -f $P > $O/${tag}_ncu_xattn.log 2>&1
}
echo done > $O/${tag}_done
