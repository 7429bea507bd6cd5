#!/bin/bash
# round 2, call 2: chain kernel tests first (bounded), then the suite, A/B bench chain / no chain / no hints, ncu of the chain kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c2
( time timeout 400 python -m pytest tests/test_gpu_chain.py -q -x -s ) > $O/${tag}_chain.log 2>&1
rc=$?; echo "chain rc=$rc" >> $O/${tag}_chain.log
if [ $rc -ne 0 ]; then echo failed > $O/${tag}_done; exit 0; fi
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${tag}_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2; do
  timeout 300 $B > $O/${tag}_bench_chain_$rep.json 2> $O/${tag}_bench_chain_$rep.err
  timeout 300 $B --flags 256 > $O/${tag}_bench_nochain_$rep.json 2> $O/${tag}_bench_nochain_$rep.err
  timeout 300 $B --flags 512 > $O/${tag}_bench_nohints_$rep.json 2> $O/${tag}_bench_nohints_$rep.err
done
P="python bench.py --steps 1 --warmup 1 --batch 944 --no-item-cache --cpu-users 0 --no-e2e"
timeout 300 $P > $O/${tag}_prof_plain.json 2> $O/${tag}_prof_plain.err && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_chain_kernel -s 2 -c 2 -o $O/${tag}_ncu_chain -f $P > $O/${tag}_ncu_chain.log 2>&1
echo done > $O/${tag}_done
