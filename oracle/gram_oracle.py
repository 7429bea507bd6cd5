"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the GRAM inference/scoring hot path.

This file is a CPU restatement (plain PyTorch fp32 ops, no custom kernels) of the reference's
algorithm.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it; nothing under `gram_b200/` does, and the product path raises
when its CUDA library is missing rather than falling back to this.

What it restates, and where the original lives (all paths relative to the reference root):

  model math          src/model/gram.py:74-107,200-256 (FiD wrapper: per-passage encode, passage
                      position embedding, view as one fused memory)
                      src/model/gram_t5.py:118-287 (enc -> dec -> d_model**-0.5 -> lm_head),
                      :289-315 (last-token slicing), :320-348 (_reorder_cache)
                      src/model/gram_t5_modeling.py:253-276 (T5LayerNorm), :297-310 (wi/ReLU/wo),
                      :337-352, :397-477 (relative position buckets + bias), :479-631 (attention),
                      :634-705, :723-837 (block), :1037-1296 (stack)
  trie                src/utils/generation_trie.py:5-95
  metrics             src/utils/evaluate.py:5-58
  beam search         third-party `transformers==4.26.0` (reference `requirements.txt:1`), NOT under
                      the reference tree and not installable offline: `generation/utils.py`
                      (`GenerationMixin.generate`, `_expand_inputs_for_generation`, `beam_search`),
                      `generation/beam_search.py` (`BeamSearchScorer`, `BeamHypotheses`),
                      `generation/logits_process.py` (`PrefixConstrainedLogitsProcessor`).  Restated
                      from the published 4.26.0 algorithm and anchored on the reference call sites
                      `src/model/gram.py:93-99`, `src/runner/single_runner_gram.py:641-651`.

Pinning status.  The model math is pinned: `tests/test_oracle.py` checks this file bit-for-bit
(torch.equal) against the real reference modules when `/root/reference` is present, and
`tests/golden/*.npz` (made by `oracle/make_golden.py`, which drives the real reference modules) are
checked everywhere.  The trie and metrics are pinned against the reference's own functions and its
one documented known answer (`src/runner/single_runner_gram.py:591-593`).  The beam-search loop is
PARITY UNPINNED against transformers 4.26.0 itself: the reference ships no golden outputs for it
and that wheel is not available here; its golden files are produced by this
restatement driving the real reference `forward`, `_reorder_cache` and `Trie`.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


# ==============================================================================================
# trie  (reference src/utils/generation_trie.py:5-95)
# ==============================================================================================

class OracleTrie:
    """Nested dict-of-dicts prefix trie; `get(prefix)` returns the allowed next tokens."""

    def __init__(self, sequences: List[List[int]] = ()):
        self.trie_dict: Dict[int, dict] = {}
        self.len = 0
        for s in sequences:
            self.add(s)

    def add(self, sequence: List[int]) -> None:
        node = self.trie_dict
        for t in sequence:
            node = node.setdefault(int(t), {})
        self.len += 1

    def get(self, prefix: List[int]) -> List[int]:
        node = self.trie_dict
        for t in prefix:
            if t in node:
                node = node[t]
            else:
                return []
        return list(node.keys())

    def __len__(self):
        return self.len


# ==============================================================================================
# metrics  (reference src/utils/evaluate.py:5-58)
# ==============================================================================================

def rel_results(predictions, targets, scores, k):
    results = []
    for b in range(len(targets)):
        seqs = predictions[b * k:(b + 1) * k]
        scs = scores[b * k:(b + 1) * k]
        pairs = sorted(zip(seqs, scs), key=lambda x: x[1], reverse=True)   # stable, descending
        results.append([1 if p[0] == targets[b] else 0 for p in pairs])
    return results


def hit_at_k(relevance, k):
    return float(sum(1 for row in relevance if sum(row[:k]) > 0))


def ndcg_at_k(relevance, k):
    total = 0.0
    for row in relevance:
        one = 0.0
        for i, r in enumerate(row[:k]):
            one += r / math.log(i + 2, 2)
        total += one
    return total


def get_metrics_results(rel, metrics):
    res = []
    for m in metrics:
        k = int(m.split("@")[1])
        if m.lower().startswith("hit"):
            res.append(hit_at_k(rel, k))
        elif m.lower().startswith("ndcg"):
            res.append(ndcg_at_k(rel, k))
    return np.array(res)


# ==============================================================================================
# model math
# ==============================================================================================

def relative_position_bucket(relative_position, bidirectional, num_buckets, max_distance):
    """reference gram_t5_modeling.py:397-450 -- kept in the same torch fp32 ops because the
    log/truncate is rounding-sensitive at bucket edges."""
    relative_buckets = 0
    if bidirectional:
        num_buckets //= 2
        relative_buckets += (relative_position > 0).to(torch.long) * num_buckets
        relative_position = torch.abs(relative_position)
    else:
        relative_position = -torch.min(relative_position, torch.zeros_like(relative_position))
    max_exact = num_buckets // 2
    is_small = relative_position < max_exact
    if_large = max_exact + (
        torch.log(relative_position.float() / max_exact) / math.log(max_distance / max_exact)
        * (num_buckets - max_exact)
    ).to(torch.long)
    if_large = torch.min(if_large, torch.full_like(if_large, num_buckets - 1))
    relative_buckets += torch.where(is_small, relative_position, if_large)
    return relative_buckets


class OracleGRAM:
    """Functional restatement of the reference GRAM model in eval mode (dropout = identity)."""

    def __init__(self, cfg, state_dict, dtype=torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        sd = {}
        for k, v in state_dict.items():
            if isinstance(v, np.ndarray):
                v = torch.from_numpy(v)
            sd[k.replace(".module.", ".")] = v.to(dtype)
        self.sd = sd
        self.H, self.dk, self.d = cfg.num_heads, cfg.d_kv, cfg.d_model
        ep = "encoder.encoder." if "encoder.encoder.final_layer_norm.weight" in sd else "encoder."
        self.ep = ep
        self.shared = sd["shared.weight"]
        self.lm_head = sd.get("lm_head.weight", self.shared)
        self.pos_emb = sd.get("position_embedding.weight", sd.get("encoder.position_embedding.weight"))

    # ---- primitives -------------------------------------------------------------------------
    def _ln(self, x, w):
        var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
        x = x * torch.rsqrt(var + self.cfg.layer_norm_epsilon)
        return w * x

    def _shape(self, x, bsz):
        return x.view(bsz, -1, self.H, self.dk).transpose(1, 2)

    def _unshape(self, x, bsz):
        return x.transpose(1, 2).contiguous().view(bsz, -1, self.H * self.dk)

    def _compute_bias(self, prefix, qlen, klen, bidirectional):
        ctx = torch.arange(qlen, dtype=torch.long)[:, None]
        mem = torch.arange(klen, dtype=torch.long)[None, :]
        bucket = relative_position_bucket(mem - ctx, bidirectional,
                                          self.cfg.relative_attention_num_buckets,
                                          self.cfg.relative_attention_max_distance)
        values = F.embedding(bucket, self.sd[prefix + ".relative_attention_bias.weight"])
        return values.permute([2, 0, 1]).unsqueeze(0)

    def _attend(self, q, k, v, position_bias, o_w, bsz):
        scores = torch.matmul(q, k.transpose(3, 2))
        scores += position_bias
        attn = F.softmax(scores.float(), dim=-1).type_as(scores)
        out = self._unshape(torch.matmul(attn, v), bsz)
        return F.linear(out, o_w)

    def _ff(self, prefix, x):
        n = self._ln(x, self.sd[prefix + ".layer_norm.weight"])
        h = F.linear(n, self.sd[prefix + ".DenseReluDense.wi.weight"])
        h = F.relu(h)
        h = F.linear(h, self.sd[prefix + ".DenseReluDense.wo.weight"])
        return x + h

    # ---- encoder (gram.py:200-256 + T5Stack encoder) ------------------------------------------
    def encode(self, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        """ids/mask `[B,N,L]` -> fused memory `[B, N*L, d]`."""
        B, N, L = input_ids.shape
        ids = input_ids.reshape(B * N, L)
        mask = attention_mask.reshape(B * N, L)
        x = F.embedding(ids, self.shared)
        ext = mask[:, None, None, :].to(self.dtype)
        ext = (1.0 - ext) * torch.finfo(self.dtype).min
        position_bias = None
        for i in range(self.cfg.num_layers):
            p = f"{self.ep}block.{i}.layer"
            a = p + ".0.SelfAttention"
            n = self._ln(x, self.sd[p + ".0.layer_norm.weight"])
            q = self._shape(F.linear(n, self.sd[a + ".q.weight"]), B * N)
            k = self._shape(F.linear(n, self.sd[a + ".k.weight"]), B * N)
            v = self._shape(F.linear(n, self.sd[a + ".v.weight"]), B * N)
            if position_bias is None:
                position_bias = self._compute_bias(a, L, L, True) + ext
            x = x + self._attend(q, k, v, position_bias, self.sd[a + ".o.weight"], B * N)
            x = self._ff(p + ".1", x)
        x = self._ln(x, self.sd[self.ep + "final_layer_norm.weight"])
        if self.pos_emb is not None and self.cfg.use_position_embedding:
            pos_ids = torch.arange(N).expand(B, N)
            pe = F.embedding(pos_ids, self.pos_emb).view(B * N, 1, -1)
            x = x + pe
        return x.view(B, N * L, -1)

    # ---- decoder (T5Stack decoder with tuple KV cache) ---------------------------------------
    def decode(self, decoder_input_ids, memory, memory_mask, past=None, use_cache=True):
        """decoder_input_ids `[R, q]` (q=1 when `past` given), memory `[R,S,d]`, mask `[R,S]`.
        Returns (logits `[R,q,V]`, present) with present[layer] = (self_k, self_v, cross_k, cross_v)."""
        R, qlen = decoder_input_ids.shape
        x = F.embedding(decoder_input_ids, self.shared)
        past_len = past[0][0].shape[2] if past is not None else 0
        real_len = past_len + qlen
        # causal self-attention mask (all-visible for the single-token cached step)
        seq = torch.arange(qlen)
        causal = (seq[None, None, :].repeat(R, qlen, 1) <= seq[None, :, None]).to(self.dtype)
        if past_len:
            causal = torch.cat([torch.ones(R, qlen, past_len, dtype=self.dtype), causal], dim=-1)
        ext_self = (1.0 - causal[:, None, :, :]) * torch.finfo(self.dtype).min
        ext_cross = (1.0 - memory_mask[:, None, None, :].to(self.dtype)) * torch.finfo(self.dtype).min
        position_bias = None
        cross_bias = None
        present = []
        for i in range(self.cfg.num_decoder_layers):
            p = f"decoder.block.{i}.layer"
            a = p + ".0.SelfAttention"
            pk = past[i] if past is not None else None
            n = self._ln(x, self.sd[p + ".0.layer_norm.weight"])
            q = self._shape(F.linear(n, self.sd[a + ".q.weight"]), R)
            k = self._shape(F.linear(n, self.sd[a + ".k.weight"]), R)
            v = self._shape(F.linear(n, self.sd[a + ".v.weight"]), R)
            if pk is not None:
                k = torch.cat([pk[0], k], dim=2)
                v = torch.cat([pk[1], v], dim=2)
            if position_bias is None:
                position_bias = self._compute_bias(a, real_len, real_len, False)
                if pk is not None:
                    position_bias = position_bias[:, :, -qlen:, :]
                position_bias = position_bias + ext_self
            x = x + self._attend(q, k, v, position_bias, self.sd[a + ".o.weight"], R)
            # cross attention
            c = p + ".1.EncDecAttention"
            n = self._ln(x, self.sd[p + ".1.layer_norm.weight"])
            cq = self._shape(F.linear(n, self.sd[c + ".q.weight"]), R)
            if pk is not None:
                ck, cv = pk[2], pk[3]
            else:
                ck = self._shape(F.linear(memory, self.sd[c + ".k.weight"]), R)
                cv = self._shape(F.linear(memory, self.sd[c + ".v.weight"]), R)
            if cross_bias is None:
                cross_bias = torch.zeros((1, self.H, real_len, memory.shape[1]), dtype=self.dtype)
                if pk is not None:
                    cross_bias = cross_bias[:, :, -qlen:, :]
                cross_bias = cross_bias + ext_cross
            x = x + self._attend(cq, ck, cv, cross_bias, self.sd[c + ".o.weight"], R)
            x = self._ff(p + ".2", x)
            if use_cache:
                present.append((k, v, ck, cv))
        x = self._ln(x, self.sd["decoder.final_layer_norm.weight"])
        if self.cfg.tie_word_embeddings:
            x = x * (self.cfg.d_model ** -0.5)
        logits = F.linear(x, self.lm_head)
        return logits, (tuple(present) if use_cache else None)

    def forward(self, input_ids, attention_mask, decoder_input_ids):
        """Teacher-forced logits `[B, q, V]` (no cache) -- mirrors `GRAM.forward` for parity checks."""
        B = input_ids.shape[0]
        memory = self.encode(input_ids, attention_mask)
        logits, _ = self.decode(decoder_input_ids, memory, attention_mask.reshape(B, -1), None, use_cache=False)
        return logits

    @staticmethod
    def reorder_cache(past, beam_idx):
        """reference gram_t5.py:320-348: index_select on all four tensors of every layer."""
        return tuple(tuple(t.index_select(0, beam_idx) for t in layer) for layer in past)

    # ---- generate = encoder + HF-4.26 beam search ----------------------------------------------
    def generate(self, input_ids, attention_mask, max_length, trie, num_beams,
                 num_return_sequences=None, length_penalty=1.0, record=None, memory=None):
        """Restated `GRAM.generate` (gram.py:74-107) -> HF 4.26 `generate(num_beams=K, ...)`.

        `trie` is any object with `.get(list[int]) -> list[int]` (the reference `Trie` or
        `OracleTrie`).  Returns dict(sequences int64 [B*R, W], sequences_scores fp32 [B*R]).
        `record`, if a list, receives one dict per step with the tensors the parity tests compare.
        """
        B = input_ids.shape[0]
        if memory is None:
            memory = self.encode(input_ids, attention_mask)
        return hf426_beam_search(
            decode_fn=self.decode, reorder_fn=self.reorder_cache, trie=trie, memory=memory,
            memory_mask=attention_mask.reshape(B, -1), max_length=max_length, num_beams=num_beams,
            num_return_sequences=num_return_sequences, length_penalty=length_penalty,
            vocab_size=self.cfg.vocab_size, eos=self.cfg.eos_token_id, pad=self.cfg.pad_token_id,
            start=self.cfg.decoder_start_token_id, record=record)


def hf426_beam_search(decode_fn, reorder_fn, trie, memory, memory_mask, max_length, num_beams,
                      num_return_sequences, length_penalty, vocab_size, eos, pad, start, record=None):
    """transformers 4.26.0 `generate(num_beams=K, num_return_sequences=R, prefix_allowed_tokens_fn=...)`
    for an encoder-decoder model whose encoder already ran: `_expand_inputs_for_generation`,
    `beam_search`, `PrefixConstrainedLogitsProcessor`, `BeamSearchScorer.process/finalize`.

    decode_fn(decoder_input_ids, memory, memory_mask, past) -> (logits [rows, q, V], past)
    reorder_fn(past, beam_idx) -> past            (reference `_reorder_cache`, gram_t5.py:320-348)
    """
    K = num_beams
    Rn = num_return_sequences if num_return_sequences is not None else 1
    if Rn > K:
        raise ValueError("`num_return_sequences` has to be smaller or equal to `num_beams`.")
    B = memory.shape[0]
    V = vocab_size
    mem_mask = memory_mask
    # _expand_inputs_for_generation: repeat_interleave(K) of memory and mask
    expand = torch.arange(B).view(-1, 1).repeat(1, K).view(-1)
    memory = memory.index_select(0, expand)
    mem_mask = mem_mask.index_select(0, expand)

    seqs = torch.full((B * K, 1), start, dtype=torch.long)
    beam_scores = torch.zeros((B, K), dtype=torch.float)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    hyps = [_BeamHypotheses(K, length_penalty) for _ in range(B)]
    done = [False] * B
    past = None
    cur_len = 1
    while True:
        dec_in = seqs[:, -1:] if past is not None else seqs
        logits, past = decode_fn(dec_in, memory, mem_mask, past)
        next_token_logits = logits[:, -1, :]
        lsm = F.log_softmax(next_token_logits, dim=-1)
        processed = prefix_constrained_scores(trie, seqs, lsm)
        scores = processed + beam_scores[:, None].expand_as(processed)
        scores = scores.view(B, K * V)
        next_scores, nt = torch.topk(scores, 2 * K, dim=1, largest=True, sorted=True)
        next_indices = torch.div(nt, V, rounding_mode="floor")
        next_tokens = nt % V
        if record is not None:
            record.append(dict(cur_len=cur_len, seqs=seqs.clone(), logits=next_token_logits.clone(),
                               lse=torch.logsumexp(next_token_logits.float(), dim=-1),
                               beam_scores=beam_scores.clone(), cand_scores=next_scores.clone(),
                               cand_tokens=next_tokens.clone(), cand_beams=next_indices.clone()))
        # ---- BeamSearchScorer.process ----
        nb_scores = torch.zeros((B, K), dtype=next_scores.dtype)
        nb_tokens = torch.zeros((B, K), dtype=torch.long)
        nb_idx = torch.zeros((B, K), dtype=torch.long)
        for b in range(B):
            if done[b]:
                nb_scores[b, :] = 0
                nb_tokens[b, :] = pad
                nb_idx[b, :] = 0
                continue
            slot = 0
            for rank in range(2 * K):
                tok = int(next_tokens[b, rank])
                sc = next_scores[b, rank]
                row = b * K + int(next_indices[b, rank])
                if tok == eos:
                    if rank >= K:
                        continue
                    hyps[b].add(seqs[row].clone(), sc.item())
                else:
                    nb_scores[b, slot] = sc
                    nb_tokens[b, slot] = tok
                    nb_idx[b, slot] = row
                    slot += 1
                if slot == K:
                    break
            if slot < K:
                raise ValueError(f"At most {K} tokens can be equal to eos_token_id")
            done[b] = done[b] or hyps[b].is_done(next_scores[b].max().item(), cur_len)
        beam_scores = nb_scores.view(-1)
        beam_idx = nb_idx.view(-1)
        seqs = torch.cat([seqs[beam_idx, :], nb_tokens.view(-1).unsqueeze(-1)], dim=-1)
        past = reorder_fn(past, beam_idx)
        cur_len += 1
        if all(done) or seqs.shape[-1] >= max_length:
            break
    # ---- BeamSearchScorer.finalize ----
    for b in range(B):
        if done[b]:
            continue
        for j in range(K):
            row = b * K + j
            hyps[b].add(seqs[row], beam_scores[row].item())
    best, best_scores, lengths = [], [], []
    for b in range(B):
        srt = sorted(hyps[b].beams, key=lambda x: x[0])
        for _ in range(Rn):
            sc, hyp = srt.pop()
            best.append(hyp)
            best_scores.append(sc)
            lengths.append(len(hyp))
    width = min(max(lengths) + 1, max_length)
    out = torch.full((B * Rn, width), pad, dtype=torch.long)
    for i, hyp in enumerate(best):
        out[i, :lengths[i]] = hyp
        if lengths[i] < width:
            out[i, lengths[i]] = eos
    return dict(sequences=out, sequences_scores=torch.tensor(best_scores, dtype=torch.float32),
                n_steps=cur_len - 1)


def prefix_constrained_scores(trie, input_ids, scores):
    """transformers 4.26.0 `PrefixConstrainedLogitsProcessor.__call__` with the reference's
    `prefix_allowed_tokens_fn` (src/utils/generation_trie.py:89-95: `trie.get(sentence.tolist())`): a mask that is -inf
    everywhere and 0 at the allowed tokens of every row, ADDED to the (already log-softmaxed) scores -- no
    renormalisation.  4.26 does not raise on an empty allowed list (5.x does): such a row becomes all -inf.
    Pinned against the installed transformers' processor in tests/test_oracle.py."""
    mask = torch.full_like(scores, -math.inf)
    for row in range(input_ids.shape[0]):
        allowed = trie.get(input_ids[row].tolist())
        mask[row, allowed] = 0
    return scores + mask


class _BeamHypotheses:
    """transformers 4.26.0 `BeamHypotheses` with `early_stopping=False`."""

    def __init__(self, num_beams, length_penalty):
        self.num_beams = num_beams
        self.length_penalty = length_penalty
        self.beams = []
        self.worst_score = 1e9

    def __len__(self):
        return len(self.beams)

    def add(self, hyp, sum_logprobs):
        score = sum_logprobs / (hyp.shape[-1] ** self.length_penalty)
        if len(self) < self.num_beams or score > self.worst_score:
            self.beams.append((score, hyp))
            if len(self) > self.num_beams:
                srt = sorted([(s, idx) for idx, (s, _) in enumerate(self.beams)])
                del self.beams[srt[0][1]]
                self.worst_score = srt[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs, cur_len):
        if len(self) < self.num_beams:
            return False
        cur_score = best_sum_logprobs / cur_len ** self.length_penalty
        return self.worst_score >= cur_score
