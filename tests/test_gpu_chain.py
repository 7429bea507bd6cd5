"""GPU tests of the row-block chain kernel (gemm_chain.cu): o-projection -> RMSNorm -> wi -> ReLU -> wo -> RMSNorm of one
encoder layer in ONE launch, ff / normalised rows handed from GEMM to GEMM through an L2-resident per-CTA scratch.
Reference arithmetic: T5LayerSelfAttention output projection + T5LayerFF (src/model/gram_t5_modeling.py:297-310,337-352,
622,634-667).  The chain does the same arithmetic per element as the three separate tcgen05 GEMM launches, so the two
must agree BIT FOR BIT; both are checked against an fp64 restatement."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _inputs(M, D, HD, F, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    ao = torch.randn(M, HD, generator=g).cuda().bfloat16()
    w_o = (torch.randn(D, HD, generator=g) * HD ** -0.5).cuda().bfloat16()
    w_i = (torch.randn(F, D, generator=g) * D ** -0.5).cuda().bfloat16()
    w_o2 = (torch.randn(D, F, generator=g) * F ** -0.5).cuda().bfloat16()
    x0 = (torch.randn(M, D, generator=g) * 2.0).cuda()
    ln1 = (1.0 + 0.25 * torch.randn(D, generator=g)).cuda()
    ln2 = (1.0 + 0.25 * torch.randn(D, generator=g)).cuda()
    return ao, w_o, w_i, w_o2, x0, ln1, ln2


def _chain(lib, ao, w_o, w_i, w_o2, x0, ln1, ln2, eps, hints=1):
    M, HD = ao.shape
    D, F = w_o.shape[0], w_i.shape[0]
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    x = x0.clone()
    xn = torch.zeros(M, D, device="cuda", dtype=torch.bfloat16)
    ss = torch.zeros(M, D // 128, device="cuda")
    scratch = torch.zeros(sms * 128 * F, device="cuda", dtype=torch.bfloat16)
    err = torch.zeros(1, device="cuda", dtype=torch.int32)
    rc = lib.gram_op_enc_chain(0, _p(ao), _p(w_o), _p(x), _p(xn), _p(ss), _p(w_i), _p(w_o2), _p(scratch), scratch.numel() * 2,
                               _p(ln1), _p(ln2), C.c_float(eps), M, D, HD, F, hints, _p(err), None)
    assert rc == 0, lib.gram_last_error(None)
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    return x, xn, ss


def _three_launches(lib, ao, w_o, w_i, w_o2, x0, ln1, ln2, eps):
    """the same layer through gram_op_gemm_norm: EPI_RESID_NORM, row-scaled EPI_RELU, EPI_RESID_NORM (or EPI_RESID)"""
    M, HD = ao.shape
    D, F = w_o.shape[0], w_i.shape[0]
    x = x0.clone()
    xn = torch.zeros(M, D, device="cuda", dtype=torch.bfloat16)
    ss = torch.zeros(M, D // 128, device="cuda")
    ff = torch.zeros(M, F, device="cuda", dtype=torch.bfloat16)
    rc = lib.gram_op_gemm_norm(0, 2, 5, _p(ao), _p(w_o), _p(x), _p(xn), _p(ss), _p(ln1), None, C.c_float(0.0), M, D, HD, None)
    assert rc == 0, lib.gram_last_error(None)
    rc = lib.gram_op_gemm_norm(0, 2, 1, _p(xn), _p(w_i), _p(ff), None, None, None, _p(ss), C.c_float(eps), M, F, D, None)
    assert rc == 0, lib.gram_last_error(None)
    if ln2 is not None:
        rc = lib.gram_op_gemm_norm(0, 2, 5, _p(ff), _p(w_o2), _p(x), _p(xn), _p(ss), _p(ln2), None, C.c_float(0.0), M, D, F, None)
    else:
        rc = lib.gram_op_gemm(0, 1, 2, 2, _p(ff), _p(w_o2), _p(x), M, D, F, None)
    assert rc == 0, lib.gram_last_error(None)
    torch.cuda.synchronize()
    return x, xn, ss


@pytest.mark.parametrize("shape", [(100, 512, 512, 2048), (389, 512, 512, 2048), (40000, 512, 512, 2048), (1000, 768, 768, 3072)])
def test_chain_op_vs_fp64_and_three_launches(shape):
    from gram_b200 import _cabi
    lib = _cabi.load_library()
    M, D, HD, F = shape
    eps = 1e-6
    ao, w_o, w_i, w_o2, x0, ln1, ln2 = _inputs(M, D, HD, F, seed=M + D)
    x, xn, ss = _chain(lib, ao, w_o, w_i, w_o2, x0, ln1, ln2, eps)
    # fp64 restatement with the path's bf16 roundings (xn, ff)
    x1 = x0.double() + ao.double() @ w_o.double().t()
    xn1 = (x1 * ln1.double()).bfloat16()
    r = torch.rsqrt((x1 ** 2).mean(-1, keepdim=True) + eps)
    ff = ((xn1.double() @ w_i.double().t()) * r).clamp_min(0).bfloat16()
    x2 = x1 + ff.double() @ w_o2.double().t()
    # (the fp64 restatement rounds xn / ff to bf16 from slightly different values: a flipped rounding moves an ff element by
    #  one bf16 ulp, ~1e-4 of |x| per flip -- the exact check is the bit equality with the three launches below)
    assert rel_err(x, x2) < 2e-3
    assert rel_err(xn, x2 * ln2.double()) < 1e-2
    assert rel_err(ss, (x2 ** 2).view(M, D // 128, 128).sum(-1)) < 5e-3
    # the three separate launches compute the same thing element for element
    y, yn, ys = _three_launches(lib, ao, w_o, w_i, w_o2, x0, ln1, ln2, eps)
    assert torch.equal(x, y) and torch.equal(xn, yn) and torch.equal(ss, ys)
    # without cache hints: same bits
    x3, xn3, ss3 = _chain(lib, ao, w_o, w_i, w_o2, x0, ln1, ln2, eps, hints=0)
    assert torch.equal(x, x3) and torch.equal(xn, xn3) and torch.equal(ss, ss3)


def test_chain_op_last_layer_variant():
    """ln_next = NULL (the last encoder layer): the second-half tiles only update x; xn / ss keep the o-projection's values"""
    from gram_b200 import _cabi
    lib = _cabi.load_library()
    M, D, HD, F = 700, 512, 512, 2048
    ao, w_o, w_i, w_o2, x0, ln1, _ = _inputs(M, D, HD, F, seed=9)
    x, xn, ss = _chain(lib, ao, w_o, w_i, w_o2, x0, ln1, None, 1e-6)
    y, yn, ys = _three_launches(lib, ao, w_o, w_i, w_o2, x0, ln1, None, 1e-6)
    assert torch.equal(x, y) and torch.equal(xn, yn) and torch.equal(ss, ys)


def test_chained_encoder_is_bit_identical_to_three_launches():
    """The engine with the encoder chain (GRAM_FLAG_ENC_CHAIN) against the three launches per layer on a many-passage batch:
    the fused memory, and the rankings that follow from it, are identical bits; the decoder-side chain likewise
    (GRAM_FLAG_NO_DEC_CHAIN is the three-launch decoder)."""
    from gram_b200 import GRAM, GramConfig, Trie, _cabi, prefix_allowed_tokens_fn, synth
    cfg = GramConfig.t5_small(max_seq_len=128, max_item_num=8)
    sd = synth.make_state_dict(cfg, seed=2)
    ids, mask = synth.make_user_batch(cfg, 96, (1, 8), 128, seed=31, min_len=2)
    ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    seqs = synth.make_item_sequences(3000, [40, 15, 5, 2], cfg.vocab_size, seed=13, variable_tail=True)
    fn = prefix_allowed_tokens_fn(Trie(seqs))
    ml = max(len(s) for s in seqs)
    outs = []
    F32 = _cabi.GRAM_FLAG_FP32_RESID     # the encoder chain works on the fp32 residual stream: compare like with like
    for flags in (_cabi.GRAM_FLAG_ENC_CHAIN, F32, F32 | _cabi.GRAM_FLAG_NO_DEC_CHAIN, 0, _cabi.GRAM_FLAG_NO_DEC_CHAIN):
        m = GRAM(cfg, dtype="bf16", device="cuda:0", flags=flags, max_users=96)
        m.load_state_dict(sd)
        mem = m.encode(ids, mask)
        o = m.generate(input_ids=ids, attention_mask=mask, max_length=ml, prefix_allowed_tokens_fn=fn, num_beams=20,
                       num_return_sequences=20, return_dict_in_generate=True)
        outs.append((mem.cpu(), o["sequences"].cpu(), o["sequences_scores"].cpu()))
        del m
    assert torch.isfinite(outs[0][0]).all() and torch.isfinite(outs[3][0]).all()
    for o in outs[1:3]:
        assert torch.equal(outs[0][0], o[0])
        assert torch.equal(outs[0][1], o[1]) and torch.equal(outs[0][2], o[2])
    # the default encoder (bf16 residual stream) with and without the decoder chain
    assert torch.equal(outs[3][0], outs[4][0]) and torch.equal(outs[3][1], outs[4][1]) and torch.equal(outs[3][2], outs[4][2])
