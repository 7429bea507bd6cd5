#!/bin/bash
# A/B on one B200 box: bash scripts/gpu_ab.sh <tag> "<pytest -k expr>" "<flag A>" "<flag B>" [ncu kernel regex]
tag=$1; kexpr=$2; fa=$3; fb=$4; ncuk=$5
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "$kexpr" ) > $O/${tag}_pytest.log 2>&1
rc=$?
echo "pytest rc=$rc" >> $O/${tag}_pytest.log
if [ $rc -eq 0 ]; then
  for rep in 1 2; do
    timeout 200 python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e $fa > $O/${tag}_bench_a_$rep.json 2> $O/${tag}_bench_a_$rep.err
    timeout 200 python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e $fb > $O/${tag}_bench_b_$rep.json 2> $O/${tag}_bench_b_$rep.err
  done
  if [ -n "$ncuk" ]; then
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$ncuk -s 4 -c 3 -o $O/${tag}_ncu -f \
      python bench.py --steps 1 --warmup 1 --batch 944 --no-item-cache --cpu-users 0 --no-e2e $fa > $O/${tag}_ncu.log 2>&1
  fi
fi
echo done > $O/${tag}_done
