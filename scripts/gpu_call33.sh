#!/bin/bash
# round 2, call 33: verification of the shipped tree: smoke, full suite, driver-shaped bench, launch list + cross-attention capture (tag r2d)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c33
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/${tag}_smoke.log 2>&1
echo "smoke rc=$?" >> $O/${tag}_smoke.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${tag}_pytest.log
timeout 600 python bench.py > $O/${tag}_bench.json 2> $O/${tag}_bench.err
P="python bench.py --steps 1 --warmup 1 --no-item-cache --cpu-users 0 --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_r2d.csv $P > $O/${tag}_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cross_attention -s 20 -c 2 -o $O/prof_xattn_r2d -f $P > $O/${tag}_ncu_xattn.log 2>&1
echo done > $O/${tag}_done
