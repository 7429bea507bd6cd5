#!/bin/bash
# round 2, call 25: transposed cross-attention tiles (beams on N; GRAM_XATTN_T=0 = beams on M): op tests, alone, suite subset, bench A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c25
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "cross_attention" ) > $O/${tag}_pytest_op.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${tag}_pytest_op.log
timeout 300 python scripts/exp_xattn_hot.py > $O/${tag}_alone_t.log 2>&1
GRAM_XATTN_T=0 timeout 300 python scripts/exp_xattn_hot.py > $O/${tag}_alone_m.log 2>&1
if [ $rc -ne 0 ]; then echo failed > $O/${tag}_done; exit 0; fi
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_live_rows.py tests/test_gpu_edge.py tests/test_gpu_item_cache.py tests/test_gpu_graph.py tests/test_gpu_bf16_path.py -q -x ) > $O/${tag}_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/${tag}_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-item-cache --cpu-users 0 --no-e2e"
for rep in 1 2 3; do
  timeout 300 $B > $O/${tag}_t_$rep.json 2> $O/${tag}_t_$rep.err
  GRAM_XATTN_T=0 timeout 300 $B > $O/${tag}_m_$rep.json 2> $O/${tag}_m_$rep.err
done
echo done > $O/${tag}_done
