"""Deterministic synthetic weights, item tries and user batches.

There is no network in the build or bench environment: no `t5-small` checkpoint, no SentencePiece
model, no `item_plain_text.txt` (SURVEY.md "facts" table).  Everything the tests and the bench feed
to the hot path is therefore generated here from integer hashes, so the *same* tensors are obtained
on every machine (no dependence on a library RNG's vectorised code path).

Weight scales follow the reference initialiser `T5PreTrainedModel._init_weights`
(reference `src/model/gram_t5_modeling.py:865-929`) with `lm_head` tied to `shared`, and the
passage-position table uses std 0.02 (`src/model/gram.py:32-33`).
"""
from __future__ import annotations

import numpy as np

from .config import GramConfig

_MASK64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK64
    return z ^ (z >> np.uint64(31))


def hash_u64(n: int, seed: int) -> np.ndarray:
    """n 64-bit hashes, a pure function of (index, seed)."""
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64)
        s = _splitmix64(np.full(1, seed, dtype=np.uint64))[0]
        return _splitmix64(idx ^ s)


def pseudo_normal(shape, std: float, seed: int) -> np.ndarray:
    """Unit-variance bell-shaped values (Irwin-Hall of four 16-bit uniforms) times `std`, fp32.

    Only exactly-representable integer/float64 arithmetic is used, so the result is bit-identical
    across platforms."""
    n = int(np.prod(shape))
    h = hash_u64(n, seed)
    m = np.uint64(0xFFFF)
    s = ((h & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(32)) & m) + ((h >> np.uint64(48)) & m))
    x = (s.astype(np.float64) - 2.0 * 65535.0) / 65536.0          # sum of 4 U(0,1) minus 2
    x = x * np.sqrt(3.0)                                            # var(sum of 4 U) = 1/3
    return (x * std).astype(np.float32).reshape(shape)


def _name_seed(seed: int, name: str) -> int:
    h = 1469598103934665603
    for ch in name.encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return (h ^ (seed * 0x9E3779B97F4A7C15)) & 0xFFFFFFFFFFFFFFFF


def make_state_dict(cfg: GramConfig, seed: int = 0, as_torch: bool = True, ln_jitter: float = 0.1):
    """Random-init GRAM weights under the reference's state-dict key names.

    Layer-norm weights are 1 + jitter so that the scale vector is actually exercised."""
    d, dk, H, dff = cfg.d_model, cfg.d_kv, cfg.num_heads, cfg.d_ff
    inner = H * dk
    sd = {}

    def put(name, shape, std, base=0.0):
        w = pseudo_normal(shape, std, _name_seed(seed, name))
        if base:
            w = (w + np.float32(base)).astype(np.float32)
        sd[name] = w

    put("shared.weight", (cfg.vocab_size, d), 1.0)

    def attn(prefix, rel_bias):
        put(prefix + ".q.weight", (inner, d), (d * dk) ** -0.5)
        put(prefix + ".k.weight", (inner, d), d ** -0.5)
        put(prefix + ".v.weight", (inner, d), d ** -0.5)
        put(prefix + ".o.weight", (d, inner), inner ** -0.5)
        if rel_bias:
            # scaled up from the init value d**-0.5 so the bias visibly changes attention
            put(prefix + ".relative_attention_bias.weight", (cfg.relative_attention_num_buckets, H), 0.5)

    for i in range(cfg.num_layers):
        p = f"encoder.encoder.block.{i}.module.layer"
        attn(p + ".0.SelfAttention", i == 0)
        put(p + ".0.layer_norm.weight", (d,), ln_jitter, 1.0)
        put(p + ".1.DenseReluDense.wi.weight", (dff, d), d ** -0.5)
        put(p + ".1.DenseReluDense.wo.weight", (d, dff), dff ** -0.5)
        put(p + ".1.layer_norm.weight", (d,), ln_jitter, 1.0)
    put("encoder.encoder.final_layer_norm.weight", (d,), ln_jitter, 1.0)
    for i in range(cfg.num_decoder_layers):
        p = f"decoder.block.{i}.layer"
        attn(p + ".0.SelfAttention", i == 0)
        put(p + ".0.layer_norm.weight", (d,), ln_jitter, 1.0)
        attn(p + ".1.EncDecAttention", False)
        put(p + ".1.layer_norm.weight", (d,), ln_jitter, 1.0)
        put(p + ".2.DenseReluDense.wi.weight", (dff, d), d ** -0.5)
        put(p + ".2.DenseReluDense.wo.weight", (d, dff), dff ** -0.5)
        put(p + ".2.layer_norm.weight", (d,), ln_jitter, 1.0)
    put("decoder.final_layer_norm.weight", (d,), ln_jitter, 1.0)
    if cfg.use_position_embedding:
        put("position_embedding.weight", (cfg.max_item_num + 1, d), 0.02)
        sd["encoder.position_embedding.weight"] = sd["position_embedding.weight"]
    sd["encoder.encoder.embed_tokens.weight"] = sd["shared.weight"]
    sd["decoder.embed_tokens.weight"] = sd["shared.weight"]
    sd["lm_head.weight"] = sd["shared.weight"]          # tied head
    if as_torch:
        import torch
        cache = {}
        out = {}
        for k, v in sd.items():
            if id(v) not in cache:
                cache[id(v)] = torch.from_numpy(v)
            out[k] = cache[id(v)]
        return out
    return sd


# ----------------------------------------------------------------------------------------------
# item tries
# ----------------------------------------------------------------------------------------------

def make_item_sequences(n_items: int, branching, vocab_size: int, seed: int = 7,
                        variable_tail: bool = False):
    """Synthetic lexical item IDs as token sequences `[0, t1..td, 1]`.

    `branching[l]` is the fan-out at depth l (BASELINE.json config 5: `[64, 25, 25, 5, 5, 1]`).
    Children of a node draw distinct tokens from `[2, vocab_size - 28)` (T5 keeps the top ids for
    sentinels).  Items are the leaves in depth-first order, truncated to `n_items`.
    With `variable_tail`, every third leaf drops its last piece (so IDs have two lengths, as the
    shipped ID files do: SURVEY.md section 8(a) row 13).
    """
    hi = vocab_size - 28
    assert hi > 2 + max(branching)
    depth = len(branching)
    seqs = []

    def child_tokens(node_key: int, fan: int):
        # distinct tokens: hash, then de-duplicate deterministically
        got, out, salt = set(), [], 0
        while len(out) < fan:
            hs = hash_u64(fan * 2, (seed * 1000003 + node_key * 7919 + salt) & 0xFFFFFFFFFFFF)
            for h in hs:
                t = 2 + int(h % np.uint64(hi - 2))
                if t not in got:
                    got.add(t)
                    out.append(t)
                    if len(out) == fan:
                        break
            salt += 1
        return out

    # iterative depth-first expansion (leaves in DFS order)
    stack = [([], 1, 0)]
    while stack and len(seqs) < n_items:
        prefix, key, lvl = stack.pop()
        if lvl == depth:
            seqs.append([0] + prefix + [1])
            continue
        toks = child_tokens(key, branching[lvl])
        for i in range(len(toks) - 1, -1, -1):
            t = toks[i]
            stack.append((prefix + [t], (key * 1315423911 + t + i) & 0xFFFFFFFFFFFF, lvl + 1))
    if variable_tail:
        # every third item drops its last piece when the shorter id is still unique
        seen = {tuple(s) for s in seqs}
        for i in range(2, len(seqs), 3):
            s = seqs[i]
            if len(s) <= 4:
                continue
            short = tuple(s[:-2] + [1])
            if short not in seen:
                seen.discard(tuple(s))
                seen.add(short)
                seqs[i] = list(short)
    return seqs


# ----------------------------------------------------------------------------------------------
# user batches
# ----------------------------------------------------------------------------------------------

def make_user_batch(cfg: GramConfig, n_users: int, n_passages, seq_len: int, seed: int = 2023,
                    min_len: int = None, full: bool = False, pad_extra_passage: bool = True):
    """Synthetic `item_text_ids [B,N,L] int64` / `item_text_masks [B,N,L] bool` obeying the collator
    contract (reference `src/processor/Collator.py:342-450`): valid tokens are a prefix of each
    passage, the last valid token is EOS(1), ids are 0 where masked, and a user with fewer passages
    than the batch maximum is padded with all-masked passages (the collator's off-by-one gives at
    least one such passage to users below the history cap: SURVEY.md section 8(a) row 1).

    `n_passages` is an int (every user) or a (lo, hi) range of *real* passages per user.
    """
    if isinstance(n_passages, int):
        lo = hi = n_passages
    else:
        lo, hi = n_passages
    h = hash_u64(n_users, seed ^ 0xABCDEF)
    real = (lo + (h % np.uint64(hi - lo + 1))).astype(np.int64)
    N = int(real.max()) + (1 if (pad_extra_passage and lo != hi) else 0)
    N = min(N, cfg.max_item_num + 1) if N > cfg.max_item_num + 1 else N
    ids = np.zeros((n_users, N, seq_len), dtype=np.int64)
    mask = np.zeros((n_users, N, seq_len), dtype=bool)
    if min_len is None:
        min_len = max(2, seq_len // 2)
    tok = hash_u64(n_users * N * seq_len, seed ^ 0x5151).reshape(n_users, N, seq_len)
    lens = hash_u64(n_users * N, seed ^ 0x7777).reshape(n_users, N)
    for u in range(n_users):
        for p in range(int(real[u])):
            ln = seq_len if full else int(min_len + (lens[u, p] % np.uint64(seq_len - min_len + 1)))
            row = 2 + (tok[u, p, :ln] % np.uint64(cfg.vocab_size - 30)).astype(np.int64)
            row[ln - 1] = 1
            ids[u, p, :ln] = row
            mask[u, p, :ln] = True
    return ids, mask
