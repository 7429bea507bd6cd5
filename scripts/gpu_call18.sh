#!/bin/bash
# round 2, call 18: final verification of the shipped tree: smoke, full suite, driver-shaped bench, reference arm, full splits through the runner
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c18
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/${tag}_smoke.log 2>&1
echo "smoke rc=$?" >> $O/${tag}_smoke.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${tag}_pytest.log
timeout 600 python bench.py > $O/${tag}_bench.json 2> $O/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${tag}_bench_ref.json 2> $O/${tag}_bench_ref.err
for ds in Beauty Toys Sports Yelp; do
  timeout 600 python scripts/eval_full.py --dataset $ds --batch 944 > $O/${tag}_eval_$ds.json 2> $O/${tag}_eval_$ds.err
done
timeout 600 python scripts/eval_full.py --dataset Beauty --batch 944 --item-cache > $O/${tag}_eval_Beauty_cached.json 2> $O/${tag}_eval_Beauty_cached.err
echo done > $O/${tag}_done
