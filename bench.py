#!/usr/bin/env python
"""bench.py -- users/sec of the GRAM inference/scoring hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W
    python bench.py --config {beauty,toys,sports,yelp,scale5} ...      (default beauty = BASELINE.json configs[1])
    torchrun --nproc-per-node 2 ... bench.py --gpus 2 --check          (N-rank rankings == 1-rank rankings, bit for bit)

Workload (BASELINE.json configs[1]): Beauty full test set, T5-small, beam 20, 20 returned sequences,
max_length 10, bf16 -- passage-batched encoder -> fused FiD memory -> cross-attention decode ->
trie-constrained beam search over the 12,101-item trie.  A "step" is one pass of the hot path over one
batch of `--batch` test users (leave-one-out split of the shipped Beauty user sequences; surrogate
tokenizer + synthetic metadata tokens + random-init tied weights, because no tokenizer / text /
checkpoint exists offline -- see gram_b200/data.py).  Users are sharded across ranks with no
data-path collective ("scaling": "weak": every rank processes `--batch` users per step); the path's one
collective -- an all-gather of the ranked lists -- runs after the timed loop and is timed as `gather_ms`.

Keys of the JSON line (see the task contract): value = whole-job users/s with inputs already resident
in HBM; e2e = the same through `GRAM.generate(...)` with pinned HOST tensors (H2D of ids/mask and D2H of
the ranked ids/scores inside the timed region); roofline = the dominant kernel class (tensor-core GEMM)
against MEASURED_PEAKS.json, on the work actually EXECUTED; cpu_baseline = the oracle port of the reference
timed on this box's cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

PROFILE_TAG = "r2d"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="beauty", choices=["beauty", "toys", "sports", "yelp", "scale5"],
                    help="beauty = BASELINE.json configs[1] (the headline); toys / sports = configs[2]; yelp = configs[3] "
                         "(30,000 synthetic users x 10 history items on the real Yelp trie); scale5 = configs[4] (T5-base, "
                         "32 x 256 tokens, beam 50, 1M-item synthetic trie)")
    ap.add_argument("--batch", type=int, default=0,
                    help="users per step per GPU; 0 = the config's default (beauty: 1888 users x 20 beams = 37760 decoder rows "
                         "= 295 row tiles of 128: two 148-SM waves for every decoder GEMM; ~65 GB of workspace)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-users", type=int, default=-1, help="users timed for cpu_baseline (0 = skip, -1 = config default)")
    ap.add_argument("--simt", action="store_true", help="force the CUDA-core GEMM (A/B timing)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--max-tokens", type=int, default=0,
                    help="workspace rows (valid encoder tokens per step); 0 = the largest step of this run + 2 %%")
    ap.add_argument("--no-item-cache", action="store_true", help="skip the extra (non-headline) cached-item measurement")
    ap.add_argument("--gemm-1cta", action="store_true", help="keep every tcgen05 GEMM on single-CTA tiles (A/B timing)")
    ap.add_argument("--mma-enc-attn", action="store_true", help="encoder attention through the mma.sync kernel instead of the tcgen05 one (A/B timing)")
    ap.add_argument("--unfused-norm", action="store_true", help="encoder RMSNorms as separate kernels instead of folded into the tcgen05 GEMMs (A/B timing)")
    ap.add_argument("--unfused-head", action="store_true", help="materialise the logits instead of the fused log-softmax head (A/B timing)")
    ap.add_argument("--flags", type=int, default=0, help="extra GRAM_FLAG_* bits (A/B timing of engine variants)")
    ap.add_argument("--all-rows", action="store_true",
                    help="decode dead beams / finished users too, as the reference does (A/B timing of live-row compaction)")
    ap.add_argument("--check", action="store_true",
                    help="no timing: 64 users sharded over the ranks through the eval loop (NCCL all-gather of the rankings) must "
                         "equal the same users evaluated by rank 0 alone, bit for bit, in fp32 and bf16")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops_sustained"], tensor_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        # one long-running nvidia-smi that prints a row every 100 ms (re-launching it per sample costs ~100 ms each)
        try:
            self._proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                           "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                          stderr=subprocess.DEVNULL, text=True)
            for line in self._proc.stdout:
                if self._stop.is_set():
                    break
                if line.strip():
                    self.rows.append([c.strip() for c in line.strip().split(",")])
        except Exception:
            pass

    def __enter__(self):
        self._proc = None
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        time.sleep(0.25)                       # let the first samples arrive before the timed region starts
        self.rows.clear()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._proc is not None:
            try:
                self._proc.terminate()         # the exact process this object started
            except Exception:
                pass
        self._t.join(timeout=3)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# --------------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs)
# --------------------------------------------------------------------------------------------------
class DatasetWorkload:
    """configs[1]-[3]: the leave-one-out test split of a shipped dataset (Yelp: synthetic users on the real trie)."""

    DEFAULT_BATCH = dict(beauty=1888, toys=1888, sports=1888, yelp=1416)

    def __init__(self, name):
        from gram_b200 import GramConfig, Trie, prefix_allowed_tokens_fn, synth
        from gram_b200.data import GramTestData
        self.name = name
        self.dataset = dict(beauty="Beauty", toys="Toys", sports="Sports", yelp="Yelp")[name]
        self.data = GramTestData(self.dataset, synthetic_users=30000 if name == "yelp" else 0)
        self.K = 20
        self.cfg = GramConfig.t5_small(max_seq_len=self.data.L, max_item_num=self.data.max_his)
        self.sd = synth.make_state_dict(self.cfg, seed=0)
        self.cands = self.data.encoded_candidates()
        self.max_length = max(len(c) for c in self.cands)
        self.trie = Trie(self.cands)
        self.fn = prefix_allowed_tokens_fn(self.trie)
        self.default_batch = self.DEFAULT_BATCH[name]
        self.default_cpu_users = 24
        self.has_item_cache = True
        self.N, self.L = self.data.max_his + 1, self.data.L
        self.n_users = self.data.n_users

    def users(self, step, batch, rank, world):
        """Users of one step on one rank: contiguous shards, wrapping around the test split."""
        start = ((step * world + rank) * batch) % self.n_users
        return [(start + i) % self.n_users for i in range(batch)]

    def collate(self, users):
        b = self.data.collate(users)
        return b["item_text_ids"], b["item_text_masks"]

    def describe(self, batch):
        d = self.data
        who = f"{d.n_users} synthetic users x 10 history items" if self.name == "yelp" else f"{d.n_users} users"
        return (f"{self.dataset} full test set ({who}, {d.n_items}-item trie), T5-small, beam {self.K}, return {self.K}, "
                f"max_length {self.max_length}, max_his {d.max_his} x {d.L} tokens")

    inputs = "surrogate tokenizer + synthetic metadata tokens + random-init tied weights (seed 0)"


class Scale5Workload:
    """configs[4]: T5-base random-init, 32 passages x 256 tokens all valid, beam 50, 1,000,000-item synthetic trie."""

    def __init__(self, items=1000000):
        from gram_b200 import GramConfig, Trie, prefix_allowed_tokens_fn, synth
        self.name = "scale5"
        self.K, self.N, self.L = 50, 32, 256
        self.cfg = GramConfig.t5_base(max_seq_len=self.L, max_item_num=self.N - 1)
        self.sd = synth.make_state_dict(self.cfg, seed=0)
        self.cands = synth.make_item_sequences(items, [64, 25, 25, 5, 5, 1], self.cfg.vocab_size, seed=7)
        self.max_length = max(len(s) for s in self.cands)
        self.trie = Trie(self.cands)
        self.fn = prefix_allowed_tokens_fn(self.trie)
        self.default_batch = 128
        self.default_cpu_users = 1
        self.has_item_cache = False
        self.n_users = 4096

    def users(self, step, batch, rank, world):
        start = ((step * world + rank) * batch) % self.n_users
        return [(start + i) % self.n_users for i in range(batch)]

    def collate(self, users):
        ids = np.empty((len(users), self.N, self.L), dtype=np.int64)
        for i, u in enumerate(users):
            ids[i] = np.random.default_rng(2023 + u).integers(2, self.cfg.vocab_size - 28, size=(self.N, self.L))
        ids[:, :, -1] = 1
        return ids, np.ones(ids.shape, dtype=bool)

    def describe(self, batch):
        return (f"configs[4] scale stressor: T5-base random-init, {self.N} passages x {self.L} tokens all valid, beam {self.K}, "
                f"{len(self.cands)}-item synthetic trie (branching 64/25/25/5/5/1), max_length {self.max_length}")

    inputs = "random token ids + random-init tied weights (seed 0)"


def make_workload(name):
    return Scale5Workload() if name == "scale5" else DatasetWorkload(name)


def flops_and_bytes(cfg, tokens, users, K, T, esz):
    """Algorithmic work of one step (SURVEY.md section 8(d) formulas)."""
    d, HD, F, V = cfg.d_model, cfg.inner_dim, cfg.d_ff, cfg.vocab_size
    Le, Ld = cfg.num_layers, cfg.num_decoder_layers
    rows = users * K
    gemm_enc = tokens * Le * (8 * d * HD + 4 * d * F)
    gemm_kv = tokens * Ld * 4 * d * HD
    gemm_dec = T * rows * Ld * (12 * d * HD + 4 * d * F)
    gemm_head = T * rows * 2 * d * V
    xattn_bytes = T * Ld * 2 * tokens * HD * esz
    return dict(gemm_enc=gemm_enc, gemm_kv=gemm_kv, gemm_dec=gemm_dec, gemm_head=gemm_head,
                gemm_total=gemm_enc + gemm_kv + gemm_dec + gemm_head, xattn_bytes=xattn_bytes)


# --------------------------------------------------------------------------------------------------
def cpu_oracle_users_per_sec(wl, users, warm=1):
    """The reference's CPU path (oracle port: reference modules' math + restated HF-4.26 beam search,
    eval_batch_size 1 as the reference runs it) timed on this box's host cores."""
    from oracle.gram_oracle import OracleGRAM, OracleTrie
    torch.set_num_threads(os.cpu_count() or 1)
    ora = OracleGRAM(wl.cfg, wl.sd)
    trie = OracleTrie(wl.cands)
    batches = []
    for u in users:
        ids, mask = wl.collate([u])
        batches.append((torch.from_numpy(ids), torch.from_numpy(mask)))
    for ids, mask in batches[:warm]:
        ora.generate(ids, mask, wl.max_length, trie, wl.K, wl.K, 1.0)
    t0 = time.perf_counter()
    for ids, mask in batches[warm:]:
        ora.generate(ids, mask, wl.max_length, trie, wl.K, wl.K, 1.0)
    dt = time.perf_counter() - t0
    n = len(batches) - warm
    return n / dt, n, dt


def metric_name(wl):
    return f"users/sec, trie-constrained beam-{wl.K} Recall@10 eval"


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = make_workload(args.config)
    batch = args.batch or wl.default_batch
    per_step = 1                                   # users per step: a bounded sample of the batch
    users = [u for s in range(args.warmup + args.steps) for u in wl.users(s, batch, 0, 1)[:per_step]]
    from oracle.gram_oracle import OracleGRAM, OracleTrie
    torch.set_num_threads(os.cpu_count() or 1)
    ora = OracleGRAM(wl.cfg, wl.sd)
    otrie = OracleTrie(wl.cands)
    times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        for u in users[s * per_step:(s + 1) * per_step]:
            ids, mask = wl.collate([u])
            ora.generate(torch.from_numpy(ids), torch.from_numpy(mask), wl.max_length, otrie, wl.K, wl.K, 1.0)
        times.append(time.perf_counter() - t0)
    timed = times[args.warmup:]
    total = sum(timed)
    value = per_step * args.steps / total
    sample = (f"{per_step} users per step (eval_batch_size 1, as the reference runs) x {args.steps} steps of the "
              f"{args.config} workload, fp32, torch CPU")
    line = dict(metric=metric_name(wl), value=value, unit="users/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=1000 * total / args.steps, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=workload_config(args, wl, batch),
                cpu_baseline=dict(value=value, unit="users/s", cores=os.cpu_count(), kind="port", sample=sample),
                e2e=dict(value=value, unit="users/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


def ncu_metrics(name, tag=PROFILE_TAG):
    """Per-launch means of the committed `ncu --set full` capture profiles/<tag>_<name>_metrics.csv (same command at the
    default batch): DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum), duration; None when no capture is present."""
    import csv
    for t in (tag, "r2c", "r2b", "r2", "r1"):
        path = os.path.join(ROOT, "profiles", f"{t}_{name}_metrics.csv")
        if os.path.exists(path):
            break
    else:
        return None
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    try:
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        vals = [float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in rows[2:]]
        return dict(traffic=float(np.mean(vals)), launches=len(vals), file=os.path.relpath(path, ROOT))
    except Exception:
        return None


def workload_config(args, wl, batch):
    return dict(workload=wl.describe(batch), users_per_step_per_gpu=batch,
                workspace="sized by the valid tokens of the largest step (max_tokens), not by users x N full passages",
                parallelism=f"user-sharded dp{args.gpus}", inputs=wl.inputs,
                decode_rows="every beam row (reference behaviour)" if args.all_rows else
                "live beams only: dead (-inf) beams and finished users are compacted away on the device before each decode step; rankings and scores bit-identical (tests/test_gpu_live_rows.py)",
                residual_stream=("fp32 (parity mode)" if args.dtype == "fp32" else
                                 "fp32 in the encoder (GRAM_FLAG_FP32_RESID)" if (int(args.flags) & 16384) else
                                 "bf16 in the encoder (updated in place by the residual GEMMs, as under model.bfloat16()), fp32 in the decoder"),
                cache="per-step working set (K/V memory + activations, > 8 GB) exceeds the 126 MB L2; every step uses different users")


# --------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def capture_stdout():
    """stdout must carry exactly ONE JSON line, but C libraries write there too (NCCL prints its version line on
    stdout whatever NCCL_DEBUG_FILE says): point fd 1 at stderr for the whole run and keep the real stdout aside."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def engine_flags(args):
    from gram_b200 import _cabi
    return ((_cabi.GRAM_FLAG_SIMT_GEMM if args.simt else 0) | (_cabi.GRAM_FLAG_MMA_ENC_ATTN if args.mma_enc_attn else 0) |
            (_cabi.GRAM_FLAG_GEMM_1CTA if args.gemm_1cta else 0) | (_cabi.GRAM_FLAG_ALL_ROWS if args.all_rows else 0) |
            (_cabi.GRAM_FLAG_UNFUSED_NORM if args.unfused_norm else 0) | (_cabi.GRAM_FLAG_UNFUSED_HEAD if args.unfused_head else 0) |
            int(args.flags))


def run_check(args, rank, world, dev, dist):
    """Multi-GPU correctness of the product path: the eval loop over 64 users, user-sharded over `world` ranks with the NCCL
    all-gather of the ranked lists, against the same loop run by rank 0 alone (reference: the distributed runner's metrics
    must equal the single-GPU runner's, distributed_runner_gram.py:832-838)."""
    from gram_b200 import GRAM
    from gram_b200.runner import GramEvalLoader, GramRunner
    wl = make_workload(args.config)
    if not isinstance(wl, DatasetWorkload):
        raise SystemExit("--check runs on the dataset configs")
    data = wl.data
    users = [(i * 97) % data.n_users for i in range(64)]

    class Args:
        metrics = "hit@5,hit@10,ndcg@5,ndcg@10"
        beam_size = wl.K
        length_penalty = 1.0
        item_id_type = "split"

    report = dict(check="multi-gpu rankings == single-gpu rankings", n_gpus=world, users=len(users), config=args.config)
    ok = True
    for dtype in ("fp32", "bf16"):
        model = GRAM(wl.cfg, dtype=dtype, device=dev, flags=engine_flags(args))
        model.load_state_dict(wl.sd)
        sharded = GramRunner(model, data.tokenizer, dev, Args(), rank, world).test_dataset_task(
            GramEvalLoader(data, 16, rank, world, users=users), "check")
        if rank == 0:
            single = GramRunner(model, data.tokenizer, dev, Args(), 0, 1).test_dataset_task(
                GramEvalLoader(data, 64, 0, 1, users=users), "check-single")
            same = (np.array_equal(sharded["sequences"], single["sequences"]) and
                    np.array_equal(sharded["sequences_scores"], single["sequences_scores"]) and
                    np.array_equal(sharded["users"], single["users"]) and sharded["metrics"] == single["metrics"] and
                    np.array_equal(sharded["hit_rank_histogram"], single["hit_rank_histogram"]))
            report[dtype] = dict(identical=bool(same), gather_seconds=sharded["gather_seconds"], metrics=sharded["metrics"])
            ok = ok and same
        del model
    if dist is not None:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.broadcast(flag, 0)
        ok = bool(flag.item())
    if rank == 0:
        report["ok"] = ok
        emit(report)
    if dist is not None:
        dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


def main():
    args = parse()
    capture_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (gram_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        # NCCL writes its version / debug lines to stdout by default; stdout must carry exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    if args.check:
        run_check(args, rank, world, dev, dist)
        return

    from gram_b200 import GRAM
    wl = make_workload(args.config)
    cfg, trie, fn, max_length = wl.cfg, wl.trie, wl.fn, wl.max_length
    model = GRAM(cfg, dtype=args.dtype, device=dev, flags=engine_flags(args))
    model.load_state_dict(wl.sd)
    B, K, W, S = (args.batch or wl.default_batch), wl.K, args.warmup, args.steps
    model.user_limit = B

    # ---- inputs: host (pinned) and device copies for every step -----------------------------------
    host_in, dev_in, tokens = [], [], []
    for s in range(W + S):
        ids_np, mask_np = wl.collate(wl.users(s, B, rank, world))
        ids = torch.from_numpy(ids_np).pin_memory()
        mask = torch.from_numpy(mask_np).pin_memory()
        host_in.append((ids, mask))
        dev_in.append((ids.to(dev), mask.to(dev)))
        tokens.append(int(mask_np.sum()))
    out_seq = torch.zeros((B * K, max_length), dtype=torch.int64, device=dev)
    out_scores = torch.zeros((B * K,), dtype=torch.float32, device=dev)
    out_width = torch.zeros((1,), dtype=torch.int32, device=dev)
    # every timed step's ranked lists, packed as the gather moves them: int32 ids + the fp32 score bit-cast
    res_pack = torch.zeros((S, B * K, max_length + 1), dtype=torch.int32, device=dev)
    # workspace sized by the VALID tokens of the largest step (+2 %), not by users x N full passages: that is what
    # lets 1,888 users share a step (the library checks the count on the device and reports a batch that exceeds it)
    max_tokens = args.max_tokens if args.max_tokens > 0 else int(max(tokens) * 1.02) + 1024
    max_tokens = min(max_tokens, B * wl.N * wl.L)
    model.configure(max_users=B, max_beams=K, max_length=max_length, max_passages=wl.N, max_seq_len=wl.L, max_tokens=max_tokens)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def kernel_step(i, keep=-1):
        ids, mask = dev_in[i]
        model.generate_into(ids, mask, max_length, trie, K, K, 1.0, out_seq, out_scores, out_width)
        if keep >= 0:
            res_pack[keep, :, :max_length].copy_(out_seq)
            res_pack[keep, :, max_length].copy_(out_scores.view(torch.int32))

    # ---- value: inputs resident in HBM ------------------------------------------------------------
    for i in range(W):
        kernel_step(i, 0)
    barrier()
    model.check_errors()
    gemm_classes = ["gemm_enc", "gemm_kv", "gemm_dec", "lm_head"]
    model.profile_begin(gemm_classes)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        ev0.record()
        for i in range(W, W + S):
            kernel_step(i, i - W)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    prof = model.profile_end()
    model.check_errors()                 # sticky device-side flags of the whole loop (the *_into calls never synchronise)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * S / (ms / 1000.0)

    # ---- the path's one collective: all-gather of the ranked lists of every timed step (after the loop) ----
    gather = dict(ms=0.0, bytes_per_rank=int(res_pack.numel() * 4), collective="none (1 GPU)")
    if dist is not None:
        recv = torch.empty((world,) + tuple(res_pack.shape), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(recv.view(-1), res_pack.view(-1))          # warm-up (communicator set-up, buffers)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(recv.view(-1), res_pack.view(-1))
        g1.record()
        barrier()
        gms = torch.tensor([g0.elapsed_time(g1)], device=dev)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        if not torch.equal(recv[rank], res_pack):
            raise SystemExit("bench.py: the gathered rankings of this rank differ from what it sent")
        gather = dict(ms=float(gms.item()), bytes_per_rank=int(res_pack.numel() * 4),
                      collective="ncclAllGather (dist.all_gather_into_tensor) of int32 [steps, users*K, max_length+1] per rank",
                      value_incl_gather=world * B * S / ((ms + float(gms.item())) / 1000.0))

    # ---- per-class device time of the same steps (second pass, all classes bracketed) ----------------
    model.profile_begin(None)
    executed = []                       # work the decode loop actually ran (live-row compaction, compact step 0)
    for i in range(W, W + S):
        kernel_step(i)
        st = model.stats()              # synchronises: second pass only
        executed.append((st["decoded_rows"], st["kv_tokens_read"]))
    prof_all = model.profile_end()

    # ---- e2e: public API with pinned host tensors ----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        def e2e_step(i):
            ids, mask = host_in[i]
            return model.generate(input_ids=ids, attention_mask=mask, max_length=max_length, prefix_allowed_tokens_fn=fn,
                                  num_beams=K, num_return_sequences=K, output_scores=True, return_dict_in_generate=True,
                                  length_penalty=1.0)
        for i in range(W):
            e2e_step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = None
        for i in range(W, W + S):
            res = e2e_step(i)
        e1.record()
        barrier()
        ems = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ems], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        ids0, mask0 = host_in[W]
        e2e = dict(value=world * B * S / (ems / 1000.0), unit="users/s",
                   h2d_bytes_per_step=int(ids0.numel() * 8 + mask0.numel()),
                   d2h_bytes_per_step=int(B * K * max_length * 8 + B * K * 4 + 4),
                   ms_per_step=ems / S)
        assert res["sequences"].shape[0] == B * K

    # ---- extra, NOT the headline: the same users through the per-item encoder-state cache ((f)-1) -------------
    item_cache = None
    if not args.no_item_cache and wl.has_item_cache:
        data = wl.data
        tab, tmask = data.item_table()
        t0 = time.time()
        model.cache_items(torch.from_numpy(tab), torch.from_numpy(tmask))
        cin = []
        for s in range(W + S):
            b = data.collate_cached(wl.users(s, B, rank, world))
            cin.append(tuple(torch.from_numpy(b[k]).pin_memory() for k in ("prompt_ids", "prompt_masks", "item_index")))
        cdev = [tuple(x.to(dev) for x in c) for c in cin]

        def cached_step(i):
            p, m, it = cdev[i]
            model.generate_cached_into(p, m, it, max_length, trie, K, K, 1.0, out_seq, out_scores, out_width)
        cached_step(0)
        torch.cuda.synchronize(dev)
        build_s = time.time() - t0
        for i in range(W):
            cached_step(i)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for i in range(W, W + S):
            cached_step(i)
        c1.record()
        barrier()
        cms = c0.elapsed_time(c1)
        model.check_errors()
        model.profile_begin(None)
        for i in range(W, W + S):
            cached_step(i)
        cprof = model.profile_end()
        # end to end: pinned host (prompt, item index) tensors in, rankings out
        def cached_e2e(i):
            p, m, it = cin[i]
            return model.generate_cached(p, m, it, max_length, prefix_allowed_tokens_fn=fn, num_beams=K,
                                         num_return_sequences=K, return_dict_in_generate=True, length_penalty=1.0)
        for i in range(W):
            cached_e2e(i)
        barrier()
        c0.record()
        for i in range(W, W + S):
            cached_e2e(i)
        c1.record()
        barrier()
        cems = c0.elapsed_time(c1)
        if dist is not None:
            t = torch.tensor([cms, cems], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cms, cems = float(t[0].item()), float(t[1].item())
        p0, m0, it0 = cin[W]
        item_cache = dict(value=world * B * S / (cms / 1000.0), unit="users/s", ms_per_step=cms / S,
                          e2e=dict(value=world * B * S / (cems / 1000.0), unit="users/s",
                                   h2d_bytes_per_step=int(p0.numel() * 8 + m0.numel() + it0.numel() * 4),
                                   d2h_bytes_per_step=int(B * K * max_length * 8 + B * K * 4 + 4)),
                          table_build_seconds=build_s, table_bytes=int(tab.shape[0] * tab.shape[1] * cfg.d_model * 4),
                          kernel_classes={c: round(v["ms"] / S, 3) for c, v in cprof.items()},
                          note="NOT the headline metric: every item passage is encoded once (gram_cache_items, outside the "
                               "timed region) and each step encodes only the user prompts; rankings and scores are "
                               "bit-identical to the passage-batched path (tests/test_gpu_item_cache.py)")

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel class -------------------------------------------------------
    # achieved = FLOPs the GEMM launches EXECUTED / their CUDA-event time.  The encoder and K/V projection execute exactly the
    # algorithmic work of SURVEY.md 8(d); the decode loop executes less than the reference computes (one row per user at step
    # 0, live beams only later), so its GEMMs are credited with the rows they ran (gram_get_stats), not with T*B*K rows.
    peaks = load_peaks()
    esz = 2 if args.dtype == "bf16" else 4
    T = max_length - 1
    tok_timed = tokens[W:W + S]
    work = [flops_and_bytes(cfg, t, B, K, T, esz) for t in tok_timed]
    dec_rows = sum(e[0] for e in executed)
    kv_tok = sum(e[1] for e in executed)
    d_, HD_, F_, V_, Ld_ = cfg.d_model, cfg.inner_dim, cfg.d_ff, cfg.vocab_size, cfg.num_decoder_layers
    exec_fl = dict(gemm_enc=sum(w["gemm_enc"] for w in work), gemm_kv=sum(w["gemm_kv"] for w in work),
                   gemm_dec=dec_rows * Ld_ * (12 * d_ * HD_ + 4 * d_ * F_), lm_head=dec_rows * 2 * d_ * V_)
    alg_fl = dict(gemm_enc=exec_fl["gemm_enc"], gemm_kv=exec_fl["gemm_kv"], gemm_dec=sum(w["gemm_dec"] for w in work),
                  lm_head=sum(w["gemm_head"] for w in work))
    gemm_ms = sum(prof[c]["ms"] for c in gemm_classes)
    gemm_launches = sum(prof[c]["launches"] for c in gemm_classes)
    achieved_tf = sum(exec_fl.values()) / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
    alg_tf = sum(alg_fl.values()) / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
    cap = ncu_metrics("gemm_enc")
    roofline = dict(bound="tensor", kernel="gemm (all nn.Linear of the path: encoder, K/V projection, decoder, lm_head)",
                    achieved=achieved_tf, peak=peaks["tensor"], unit="TFLOP/s", frac=achieved_tf / peaks["tensor"],
                    work="executed FLOPs (decode-phase GEMMs counted on the rows actually decoded)",
                    algorithmic_tflops=alg_tf, algorithmic_frac=alg_tf / peaks["tensor"],
                    traffic=cap["traffic"] if cap else None,
                    traffic_note=(f"mean DRAM bytes per launch over the {cap['launches']} encoder-GEMM launches of {cap['file']} "
                                  "(one layer: q|k|v, o, wi, wo); per-launch traffic is shape dependent") if cap else None,
                    peak_source=f"{peaks['src']} sustained bf16", launches=gemm_launches,
                    avg_launch_ms=gemm_ms / max(gemm_launches, 1), share_of_step=gemm_ms / ms)
    by_class = {}
    for cname in gemm_classes:
        t_ms = prof[cname]["ms"]
        if t_ms > 0:
            tf = exec_fl[cname] / (t_ms / 1000.0) / 1e12
            by_class[cname] = dict(tflops=tf, frac=tf / peaks["tensor"], ms_per_step=t_ms / S,
                                   algorithmic_tflops=alg_fl[cname] / (t_ms / 1000.0) / 1e12)
    roofline["by_class"] = by_class
    if not (args.unfused_norm or args.simt or args.dtype == "fp32"):
        roofline["note"] = ("the encoder GEMM time includes the 12 per-layer RMSNorms folded into the GEMM epilogues (12 ms per "
                            "step as separate kernels at the default batch; --unfused-norm shows the GEMMs alone)")
    exec_xa_bytes = kv_tok * Ld_ * 2 * HD_ * esz
    executed_work = dict(
        decoder_rows_per_step=dec_rows / S, algorithmic_decoder_rows_per_step=float(T * B * K),
        kv_tokens_read_per_step=kv_tok / S, algorithmic_kv_tokens_per_step=float(T * np.mean(tok_timed)),
        note="the decode loop runs fewer rows / reads fewer K/V tokens than the reference computes; rooflines use these figures")
    xa_ms = prof_all["cross_attn"]["ms"]
    xa_alg_bytes = sum(w["xattn_bytes"] for w in work)
    xa_gbs = exec_xa_bytes / (xa_ms / 1000.0) / 1e9 if xa_ms > 0 else 0.0
    total_all = sum(v["ms"] for v in prof_all.values())
    kernels = {c: dict(ms_per_step=v["ms"] / S, launches_per_step=v["launches"] // S,
                       share=v["ms"] / total_all if total_all else 0.0) for c, v in prof_all.items()}
    xcap = ncu_metrics("xattn")
    roofline_cross = dict(bound="hbm", kernel="cross_attention_persist_kernel, transposed tiles (cross_attention_mma_kernel with --flags 8192)", achieved=xa_gbs, peak=peaks["hbm"], unit="GB/s",
                          frac=xa_gbs / peaks["hbm"], work="executed bytes (K and V of users that still have live beams)",
                          algorithmic_gbs=xa_alg_bytes / (xa_ms / 1000.0) / 1e9 if xa_ms > 0 else 0.0,
                          traffic=xcap["traffic"] if xcap else None,
                          executed_bytes_per_launch=exec_xa_bytes / max(prof_all["cross_attn"]["launches"], 1),
                          peak_source=f"{peaks['src']} copy bandwidth")

    # ---- CPU baseline (bounded sample) -----------------------------------------------------------------
    cpu = None
    n_cpu = wl.default_cpu_users if args.cpu_users < 0 else args.cpu_users
    if n_cpu > 0 and world == 1:
        ups, n, dt = cpu_oracle_users_per_sec(wl, wl.users(W, B, 0, 1)[:n_cpu + 1])
        cpu = dict(value=ups, unit="users/s", cores=os.cpu_count(), kind="port",
                   sample=f"{n} users of the first timed batch, eval_batch_size 1, fp32 torch CPU, {dt:.1f} s",
                   note="a reported baseline, not a like-for-like comparison: the CPU port is fp32 at 1 user per call (as the "
                        "reference runs), the GPU line is bf16 at users_per_step_per_gpu users per call")

    line = dict(metric=metric_name(wl), value=value, unit="users/s", n_gpus=world,
                steps=S, warmup=W, ms_per_step=ms / S, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype=args.dtype, data="synthetic", config=workload_config(args, wl, B),
                clocks=clocks.summary(), e2e=e2e, gpu_launches=int(model.stats()["launches"]) * S,
                roofline=roofline, roofline_cross_attention=roofline_cross, executed_work=executed_work, kernel_classes=kernels,
                gather=gather, cpu_baseline=cpu, item_cache=item_cache, tokens_per_step=float(np.mean(tok_timed)),
                gemm_impl="simt" if (args.simt or args.dtype == "fp32") else "tcgen05",
                notes="roofline = all tcgen05 GEMM launches of the timed steps (CUDA events recorded by the library on the "
                      "launching stream); kernel_classes come from a second pass over the same steps with every class bracketed; "
                      "gather = the ranked-list all-gather after the timed loop (not part of value)")
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
