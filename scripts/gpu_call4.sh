#!/bin/bash
# round 2, call 4: suite (long-passage tcgen05 attention, pinned-output generate, fast runner), smoke, default bench, scale5 A/B, full splits
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c4
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/${tag}_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${tag}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/${tag}_smoke.log 2>&1
echo "smoke rc=$?" >> $O/${tag}_smoke.log
( time timeout 900 python bench.py ) > $O/${tag}_bench.json 2> $O/${tag}_bench.err
timeout 900 python bench.py --config scale5 --steps 4 --warmup 3 --cpu-users 0 > $O/${tag}_scale5_tc.json 2> $O/${tag}_scale5_tc.err
timeout 900 python bench.py --config scale5 --steps 4 --warmup 3 --cpu-users 0 --no-e2e --flags 2048 > $O/${tag}_scale5_mma.json 2> $O/${tag}_scale5_mma.err
for ds in Beauty Toys Sports Yelp; do
  timeout 600 python scripts/eval_full.py --dataset $ds --batch 944 > $O/${tag}_eval_$ds.json 2> $O/${tag}_eval_$ds.err
done
timeout 600 python scripts/eval_full.py --dataset Beauty --batch 1888 > $O/${tag}_eval_Beauty_1888.json 2> $O/${tag}_eval_Beauty_1888.err
timeout 600 python scripts/eval_full.py --dataset Beauty --batch 944 --item-cache > $O/${tag}_eval_Beauty_cached.json 2> $O/${tag}_eval_Beauty_cached.err
echo done > $O/${tag}_done
