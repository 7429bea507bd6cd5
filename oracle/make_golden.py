"""TEST INFRASTRUCTURE ONLY -- regenerate `tests/golden/*.npz` from the REAL reference modules.

Run in the build container (needs `/root/reference`):   python -m oracle.make_golden

For every case in `tests/helpers.py:CASES` the reference's own `GRAM` model
(`/root/reference/src/model`, imported through `oracle/ref_shim.py`, weights from
`gram_b200.synth.make_state_dict`) computes

  memory      EncoderWrapper.forward output             (src/model/gram.py:200-256)
  logits      GRAM.forward teacher-forced logits         (src/model/gram.py:51-69)
  sequences / sequences_scores / per-step lse + candidates
              the restated transformers-4.26 beam loop (`oracle.gram_oracle.hf426_beam_search`)
              driving the reference `forward` with its tuple KV cache, the reference
              `_reorder_cache` (src/model/gram_t5.py:320-348) and the reference `Trie`
              (src/utils/generation_trie.py)

Large tensors are stored as deterministic sub-samples (the index arrays are stored with them).
The fixtures cannot be produced on the GPU box (no reference there); they travel in git.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import CASES, GOLDEN_DIR  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.gram_oracle import hf426_beam_search  # noqa: E402


def teacher_tokens(seqs, n_users, q):
    """decoder_input_ids for the logits fixture: real item-id prefixes (padded with 0)."""
    out = np.zeros((n_users, q), dtype=np.int64)
    for u in range(n_users):
        s = seqs[(u * 7 + 3) % len(seqs)]
        s = s[:-1][:q]                       # drop EOS
        out[u, :len(s)] = s
    return out


def vocab_sample(V, n=256):
    idx = np.unique((np.arange(n, dtype=np.int64) * 2654435761 % V))
    return idx


def make_case(case):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = ref_shim.load_reference()
    sd, ids, mask, seqs, max_length = case.build()
    cfg = case.cfg
    hf = ref_shim.make_reference_config(
        vocab_size=cfg.vocab_size, d_model=cfg.d_model, d_kv=cfg.d_kv, d_ff=cfg.d_ff, num_layers=cfg.num_layers,
        num_decoder_layers=cfg.num_decoder_layers, num_heads=cfg.num_heads, max_seq_len=cfg.max_seq_len,
        max_item_num=cfg.max_item_num)
    model = ref_shim.build_reference_model(hf, sd)
    BO = sys.modules["gram_ref_model.gram_t5_outputs"].BaseModelOutputWithPastAndCrossAttentions
    B, N, L = ids.shape
    model.encoder.n_passages = N
    memory = model.encoder(input_ids=ids.view(B, -1), attention_mask=mask.view(B, -1), return_dict=True)[0]
    q = max_length - 1
    dec = torch.from_numpy(teacher_tokens(seqs, B, q))
    logits = model(input_ids=ids, attention_mask=mask, decoder_input_ids=dec, return_dict=True).logits

    def decode_fn(dec_in, mem, mem_mask, past):
        o = model(decoder_input_ids=dec_in, past_key_values=past, encoder_outputs=BO(last_hidden_state=mem),
                  attention_mask=mem_mask, use_cache=True, return_dict=True)
        return o.logits, o.past_key_values

    trie = ref.generation_trie.Trie(seqs)
    rec = []
    out = hf426_beam_search(decode_fn, model._reorder_cache, trie, memory, mask.view(B, -1), max_length,
                            case.num_beams, case.num_beams, case.length_penalty, cfg.vocab_size,
                            cfg.eos_token_id, cfg.pad_token_id, cfg.decoder_start_token_id, record=rec)
    vs = vocab_sample(cfg.vocab_size)
    valid = mask.view(B, -1).numpy()
    rows = np.argwhere(valid)                       # (b, s) of valid memory positions
    rows = rows[:: max(1, len(rows) // 96)]
    sc = out["sequences_scores"].view(B, -1)
    gold = dict(
        memory_rows=rows.astype(np.int32),
        memory=memory.numpy()[rows[:, 0], rows[:, 1]].astype(np.float32),
        dec_ids=dec.numpy(),
        vocab_idx=vs.astype(np.int32),
        logits=logits.numpy()[:, :, vs].astype(np.float32),
        logits_lse=torch.logsumexp(logits, -1).numpy().astype(np.float32),
        logits_absmax=np.float32(logits.abs().max().item()),
        sequences=out["sequences"].numpy(),
        sequences_scores=out["sequences_scores"].numpy(),
        n_steps=np.int32(out["n_steps"]),
        step_lse=np.stack([r["lse"].numpy() for r in rec]).astype(np.float32),
        step_cand_scores=np.stack([r["cand_scores"].numpy() for r in rec]).astype(np.float32),
        step_cand_tokens=np.stack([r["cand_tokens"].numpy() for r in rec]).astype(np.int32),
        step_cand_beams=np.stack([r["cand_beams"].numpy() for r in rec]).astype(np.int32),
        min_rank_gap=np.float32((sc[:, :-1] - sc[:, 1:]).min().item()),
        max_length=np.int32(max_length),
    )
    return gold


def main():
    if not ref_shim.reference_available():
        raise SystemExit("the reference tree is not mounted; golden fixtures can only be made in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, case in CASES.items():
        gold = make_case(case)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **gold)
        print(f"{name}: sequences {gold['sequences'].shape} steps {int(gold['n_steps'])} "
              f"min_rank_gap {float(gold['min_rank_gap']):.3e} -> {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
