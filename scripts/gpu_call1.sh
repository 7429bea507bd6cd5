#!/bin/bash
# round 2, call 1: whole GPU suite (new bf16-path parity tests included), smoke, default bench, 1 vs 2 streams
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c1
( time timeout 1500 python -m pytest tests -m gpu -q -x -s ) > $O/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${tag}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/${tag}_smoke.log 2>&1
echo "smoke rc=$?" >> $O/${tag}_smoke.log
( time timeout 600 python bench.py ) > $O/${tag}_bench.json 2> $O/${tag}_bench.err
for n in 1 2; do timeout 300 python scripts/exp_streams.py $n 944 1200000 > $O/${tag}_streams_$n.log 2>&1; done
for n in 1 2; do timeout 300 python scripts/exp_streams.py $n 1888 2350000 > $O/${tag}_streams1888_$n.log 2>&1; done
echo done > $O/${tag}_done
