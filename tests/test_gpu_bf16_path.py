"""GPU parity of the path that is BENCHMARKED: bf16, fused log-softmax head (EPI_LSE epilogue + lse_combine + candidate logits
recomputed as hidden . head[token]), tcgen05 encoder attention at full passage length -- each compared with the oracle directly.

Bars (BASELINE.json north_star): logits-level quantities within 2e-2 relative (bf16) -- "relative" = max |a - b| over the
tensor divided by max |reference logits| --, ranked ids bit-exact under fp32, top-10 overlap reported for bf16.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from helpers import CASES, GOLDEN_DIR, oracle_for, rel_err

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2
K20 = 20


def _golden(name):
    return np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))


def test_fused_lse_head_op():
    """gram_op_lse_head = the vocabulary GEMM with the log-softmax statistics in its epilogue (ex2.approx) + lse_combine,
    against torch.logsumexp of the fp32 product of the same bf16 operands; T5-small head shape, ragged row counts."""
    from gram_b200 import _cabi
    lib = _cabi.load_library()
    V, D = 32128, 512
    g = torch.Generator(device="cpu").manual_seed(5)
    head = (torch.randn(V, D, generator=g)).cuda().bfloat16()
    for M in (1, 77, 300, 5000):
        hid = (torch.randn(M, D, generator=g) * (D ** -0.5) * 3.0).cuda().bfloat16()
        lse = torch.zeros(M, device="cuda")
        part = torch.zeros(M * ((V + 127) // 128 + 2) * 2, device="cuda")
        rc = lib.gram_op_lse_head(0, C.c_void_p(hid.data_ptr()), C.c_void_p(head.data_ptr()), C.c_void_p(lse.data_ptr()),
                                  C.c_void_p(part.data_ptr()), M, V, D, None)
        assert rc == 0, lib.gram_last_error(None)
        torch.cuda.synchronize()
        logits = hid.float() @ head.float().t()
        want = torch.logsumexp(logits.double(), -1)
        err = float((lse.double() - want).abs().max() / logits.abs().max())
        print(f"[fused lse head] M={M} rel_err={err:.3e} (|logits|max {float(logits.abs().max()):.2f})")
        assert err < 1e-4                                   # same bf16 operands: only ex2.approx / summation order differ


@pytest.mark.parametrize("name", list(CASES))
def test_bf16_fused_head_step_taps(name):
    """Per decode step of the shipped bf16 path (GRAM_FLAG_KEEP_LOGITS now only records): the log-sum-exp of every live beam
    row and the beam scores (sum of candidate log-probs, candidates recomputed as hidden . head[token] - lse) against the
    fp32 oracle's row with the same token prefix."""
    from gram_b200 import GRAM, Trie, _cabi, prefix_allowed_tokens_fn
    from oracle.gram_oracle import OracleTrie
    case = CASES[name]
    sd, ids, mask, seqs, max_length = case.build()
    B, K = case.n_users, case.num_beams
    m = GRAM(case.cfg, dtype="bf16", device="cuda:0", flags=_cabi.GRAM_FLAG_KEEP_LOGITS)
    m.load_state_dict(sd)
    fn = prefix_allowed_tokens_fn(Trie(seqs))
    m.generate(input_ids=ids.cuda(), attention_mask=mask.cuda(), max_length=max_length, prefix_allowed_tokens_fn=fn,
               num_beams=K, num_return_sequences=K, return_dict_in_generate=True, length_penalty=case.length_penalty)
    lse, bsc, tok = m.step_taps(B, K)
    rec = []
    oracle_for(case, sd).generate(ids, mask, max_length, OracleTrie(seqs), K, K, case.length_penalty, record=rec)
    absmax = float(_golden(name)["logits_absmax"])
    worst_lse = worst_sc = 0.0
    matched = total = 0
    for t, r in enumerate(rec):
        cur_len = t + 1
        o_seqs, o_lse, o_sc = r["seqs"].numpy(), r["lse"].numpy(), r["beam_scores"].numpy()
        for u in range(B):
            table = {}
            for j in range(K):
                row = u * K + j
                if o_sc[row] > -1e8:
                    table.setdefault(tuple(o_seqs[row].tolist()), row)
            for j in range(K):
                row = u * K + j
                if not (bsc[t, row] > -1e8) or (t > 0 and tok[t, row, cur_len - 1] == case.cfg.pad_token_id):
                    continue                                  # dead beam / finished user: not decoded
                total += 1
                o = table.get(tuple(tok[t, row, :cur_len].tolist()))
                if o is None:
                    continue                                  # bf16 kept a prefix the fp32 beam dropped
                matched += 1
                worst_lse = max(worst_lse, abs(float(lse[t, row]) - float(o_lse[o])) / absmax)
                worst_sc = max(worst_sc, abs(float(bsc[t, row]) - float(o_sc[o])) / (absmax * max(t, 1)))
    print(f"[bf16 fused head taps] case={name}: {matched}/{total} live rows share a prefix with the fp32 oracle; "
          f"lse rel_err {worst_lse:.3e}, beam-score rel_err per step {worst_sc:.3e}")
    assert matched >= 0.6 * total and matched >= B            # step 0 always matches
    assert worst_lse < BF16_TOL and worst_sc < BF16_TOL


def test_bf16_fused_head_equals_unfused_head():
    """A/B of the two heads on the same handle configuration: rankings of the fused path against GRAM_FLAG_UNFUSED_HEAD
    (fp32 logits materialised, lse_rows) -- scores within the bf16 bar, top-1 identical."""
    from gram_b200 import GRAM, Trie, _cabi, prefix_allowed_tokens_fn
    case = CASES["small"]
    sd, ids, mask, seqs, max_length = case.build()
    K = case.num_beams
    fn = prefix_allowed_tokens_fn(Trie(seqs))
    outs = []
    for flags in (0, _cabi.GRAM_FLAG_UNFUSED_HEAD):
        m = GRAM(case.cfg, dtype="bf16", device="cuda:0", flags=flags)
        m.load_state_dict(sd)
        o = m.generate(input_ids=ids.cuda(), attention_mask=mask.cuda(), max_length=max_length, prefix_allowed_tokens_fn=fn,
                       num_beams=K, num_return_sequences=K, return_dict_in_generate=True)
        outs.append((o["sequences"].cpu().numpy(), o["sequences_scores"].cpu().numpy()))
    (s0, c0), (s1, c1) = outs
    w = min(s0.shape[1], s1.shape[1])
    assert np.array_equal(s0[::K, :w], s1[::K, :w])           # top-1 of every user
    assert np.abs(c0.reshape(-1, K)[:, 0] - c1.reshape(-1, K)[:, 0]).max() < BF16_TOL * float(_golden("small")["logits_absmax"])


@pytest.mark.parametrize("stream", ["fp32_stream", "bf16_stream"])
def test_bf16_encoder_full_length_vs_oracle(stream):
    """bf16 `encode()` at the headline shape -- 21 passages x 128 tokens per user, which is what puts the persistent tcgen05
    attention kernel (attention_tc.cu) and the folded-RMSNorm GEMMs on the path -- against OracleGRAM.encode (fp32,
    bit-identical to the reference modules): < 2e-2 of the memory's magnitude.  Both residual-stream precisions: fp32
    (GRAM_FLAG_FP32_RESID) and bf16 (the default: in-place EPI_RESID_BF16 updates, RMSNorm gains folded into the q|k|v / wi weights)."""
    from gram_b200 import GRAM, GramConfig, _cabi, synth
    from oracle.gram_oracle import OracleGRAM
    cfg = GramConfig.t5_small(max_seq_len=128, max_item_num=20)
    sd = synth.make_state_dict(cfg, seed=0)
    ids, mask = synth.make_user_batch(cfg, 3, (21, 21), 128, seed=17, full=True)
    mask = mask.copy()
    ids = ids.copy()
    mask[1, 20, :] = False                                    # the collator's all-masked extra passage
    ids[1, 20, :] = 0
    mask[2, 5, 100:] = False                                  # a shorter passage (ids 0 where masked, EOS last)
    ids[2, 5, 100:] = 0
    ids[2, 5, 99] = 1
    ids, mask = torch.from_numpy(ids), torch.from_numpy(mask)
    m = GRAM(cfg, dtype="bf16", device="cuda:0", flags=_cabi.GRAM_FLAG_FP32_RESID if stream == "fp32_stream" else 0)
    m.load_state_dict(sd)
    mem = m.encode(ids.cuda(), mask.cuda()).cpu()
    torch.set_num_threads(os.cpu_count() or 1)
    want = OracleGRAM(cfg, sd).encode(ids, mask)
    fm = mask.view(3, -1)
    err = rel_err(mem[fm], want[fm])
    print(f"[bf16 encoder, 21 x 128, {stream}] memory rel_err vs oracle = {err:.3e}")
    assert torch.isfinite(mem).all() and err < BF16_TOL


def _near_tie_ok(got, want, wsc, tol):
    """rankings that differ only by permutations inside runs of oracle scores closer than `tol`"""
    if got.shape != want.shape:
        return False
    i, n = 0, len(want)
    while i < n:
        j = i
        while j + 1 < n and abs(float(wsc[j + 1]) - float(wsc[j])) < tol:
            j += 1
        a = sorted(map(tuple, got[i:j + 1].tolist()))
        b = sorted(map(tuple, want[i:j + 1].tolist()))
        if a != b:
            return False
        i = j + 1
    return True


def _parity_vs_oracle(dataset, n_users, stride, synthetic_users=0, bf16=True, both_streams=False):
    from gram_b200 import GRAM, GramConfig, Trie, _cabi, prefix_allowed_tokens_fn, synth
    from gram_b200.data import GramTestData
    from oracle.gram_oracle import OracleGRAM, OracleTrie
    data = GramTestData(dataset, synthetic_users=synthetic_users)
    cfg = GramConfig.t5_small(max_seq_len=data.L, max_item_num=data.max_his)
    sd = synth.make_state_dict(cfg, seed=0)
    cands = data.encoded_candidates()
    ml = max(len(c) for c in cands)
    fn = prefix_allowed_tokens_fn(Trie(cands))
    users = [(i * stride) % data.n_users for i in range(n_users)]
    batch = data.collate(users)
    ids, mask = torch.from_numpy(batch["item_text_ids"]), torch.from_numpy(batch["item_text_masks"])
    out = {}
    variants = [("fp32", "fp32", 0)] + ([("bf16", "bf16", 0)] if bf16 else []) + \
               ([("bf16_fp32stream", "bf16", _cabi.GRAM_FLAG_FP32_RESID)] if both_streams else [])
    for name, dtype, flags in variants:
        m = GRAM(cfg, dtype=dtype, device="cuda:0", max_users=n_users, flags=flags)
        m.load_state_dict(sd)
        o = m.generate(input_ids=ids.cuda(), attention_mask=mask.cuda(), max_length=ml, prefix_allowed_tokens_fn=fn,
                       num_beams=K20, num_return_sequences=K20, return_dict_in_generate=True)
        out[name] = (o["sequences"].cpu().numpy(), o["sequences_scores"].cpu().numpy())
        del m
    torch.set_num_threads(os.cpu_count() or 1)
    ora, otrie = OracleGRAM(cfg, sd), OracleTrie(cands)
    rep = dict(exact=0, near_tie=[], mismatch=[], overlaps=[], top1=0, err32=0.0, err16=0.0, min_gap=np.inf, ml=ml,
               overlaps_f=[], top1_f=0, err16_f=0.0)
    for i, u in enumerate(users):
        b1 = data.collate([u])                                # the reference evaluates one user per call
        ref = ora.generate(torch.from_numpy(b1["item_text_ids"]), torch.from_numpy(b1["item_text_masks"]), ml, otrie, K20, K20, 1.0)
        want, wsc = ref["sequences"].numpy(), ref["sequences_scores"].numpy()
        w = want.shape[1]
        got32 = out["fp32"][0][i * K20:(i + 1) * K20]
        assert not got32[:, w:].any()                         # nothing beyond the reference's width
        got32 = got32[:, :w]
        rep["min_gap"] = min(rep["min_gap"], float(np.abs(np.diff(wsc)).min()))
        rep["err32"] = max(rep["err32"], float(np.abs(out["fp32"][1][i * K20:(i + 1) * K20] - wsc).max()))
        if np.array_equal(got32, want):
            rep["exact"] += 1
        elif _near_tie_ok(got32, want, wsc, 4 * np.spacing(np.float32(np.abs(wsc).max()))):
            rep["near_tie"].append(int(u))
        else:
            rep["mismatch"].append(int(u))
        if bf16:
            got16 = out["bf16"][0][i * K20:(i + 1) * K20, :w]
            rep["overlaps"].append(len({tuple(r) for r in want[:10].tolist()} & {tuple(r) for r in got16[:10].tolist()}) / 10)
            rep["top1"] += bool(np.array_equal(got16[0], want[0]))
            rep["err16"] = max(rep["err16"], float(np.abs(out["bf16"][1][i * K20:(i + 1) * K20] - wsc).max()))
        if both_streams:
            got = out["bf16_fp32stream"][0][i * K20:(i + 1) * K20, :w]
            rep["overlaps_f"].append(len({tuple(r) for r in want[:10].tolist()} & {tuple(r) for r in got[:10].tolist()}) / 10)
            rep["top1_f"] += bool(np.array_equal(got[0], want[0]))
            rep["err16_f"] = max(rep["err16_f"], float(np.abs(out["bf16_fp32stream"][1][i * K20:(i + 1) * K20] - wsc).max()))
    return rep


@pytest.mark.timeout(900)
def test_beauty_64_users_vs_oracle():
    """Headline configuration (Beauty, T5-small, 12,101-item trie, beam 20), 64 real test users spread over the split, the
    CUDA path in ONE batched call against the oracle one user per call: fp32 ranked ids identical (users whose ranking
    differs only inside a run of oracle scores closer than 4 ulp are reported separately, SURVEY.md 8(c) tie zone); bf16
    (the default: bf16 residual stream in the encoder) top-1 identical for every user, top-10 overlap >= 0.8 for every user
    and >= 0.95 on average; the same bars for bf16 with the fp32 stream (GRAM_FLAG_FP32_RESID)."""
    rep = _parity_vs_oracle("Beauty", 64, 97, both_streams=True)
    print(f"[Beauty 64 users] fp32 identical {rep['exact']}/64, near-tie users {rep['near_tie']}, mismatches {rep['mismatch']}, "
          f"max fp32 score err {rep['err32']:.2e}, min adjacent oracle gap {rep['min_gap']:.2e}; bf16 top-1 {rep['top1']}/64, "
          f"top-10 overlap mean {np.mean(rep['overlaps']):.3f} min {np.min(rep['overlaps']):.1f}, max score err {rep['err16']:.3f}; "
          f"bf16 with the fp32 stream: top-1 {rep['top1_f']}/64, top-10 overlap mean {np.mean(rep['overlaps_f']):.3f} "
          f"min {np.min(rep['overlaps_f']):.1f}, max score err {rep['err16_f']:.3f}")
    assert not rep["mismatch"]
    assert rep["exact"] + len(rep["near_tie"]) == 64 and len(rep["near_tie"]) <= 3
    assert rep["err32"] < 2e-4
    assert rep["top1"] == 64 and min(rep["overlaps"]) >= 0.8 and np.mean(rep["overlaps"]) >= 0.95
    # (the same bars for both streams: which single user loses a second item at the rank-10 boundary moves with any change of
    #  summation order -- an experiment that split the softmax row sums into four chains turned the fp32 stream's min 0.9 into
    #  0.8 and the bf16 stream's 0.8 into 0.9)
    assert rep["top1_f"] == 64 and min(rep["overlaps_f"]) >= 0.8 and np.mean(rep["overlaps_f"]) >= 0.95


@pytest.mark.timeout(900)
@pytest.mark.parametrize("dataset", ["Toys", "Sports", "Yelp"])
def test_other_tries_fp32_bit_exact_vs_oracle(dataset):
    """BASELINE configs[2]/[3]: the other shipped tries (id lengths 7/8, 9/10, 11/12 tokens; root fan-outs 30/28/21 > K is NOT
    given, so step 1 starts with dead beams) -- 8 users each, fp32 ranked ids against the oracle, bf16 overlap reported."""
    rep = _parity_vs_oracle(dataset, 8, 131, synthetic_users=64 if dataset == "Yelp" else 0)
    print(f"[{dataset} 8 users, max_length {rep['ml']}] fp32 identical {rep['exact']}/8, near-tie {rep['near_tie']}, mismatches "
          f"{rep['mismatch']}, max fp32 score err {rep['err32']:.2e}, min gap {rep['min_gap']:.2e}; bf16 top-1 {rep['top1']}/8, "
          f"top-10 overlap {rep['overlaps']}")
    assert not rep["mismatch"] and rep["exact"] + len(rep["near_tie"]) == 8
    assert rep["err32"] < 2e-4
    assert rep["top1"] >= 7 and min(rep["overlaps"]) >= 0.8
