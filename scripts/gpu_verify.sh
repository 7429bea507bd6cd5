#!/bin/bash
# Full verification on a B200: GPU test suite, smoke, default bench.  Logs under gpurun_out/<tag>_*.
tag=${1:-verify}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${tag}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/${tag}_smoke.log 2>&1
echo "smoke rc=$?" >> $O/${tag}_smoke.log
( time timeout 600 python bench.py ) > $O/${tag}_bench.json 2> $O/${tag}_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e --unfused-norm > $O/${tag}_bench_unfused.json 2> $O/${tag}_bench_unfused.err
echo done > $O/${tag}_done
