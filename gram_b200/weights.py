"""State-dict import: reference key names -> canonical tensor names of the C ABI.

The reference model registers its tensors under:

  shared.weight, lm_head.weight, position_embedding.weight (== encoder.position_embedding.weight)
  encoder.encoder.block.<i>.module.layer.0.SelfAttention.{q,k,v,o}.weight           (wrapped FiD encoder,
  encoder.encoder.block.<i>.module.layer.0.layer_norm.weight                         src/model/gram.py:140-149,
  encoder.encoder.block.<i>.module.layer.1.DenseReluDense.{wi,wo}.weight             291-300)
  encoder.encoder.block.0.module.layer.0.SelfAttention.relative_attention_bias.weight
  encoder.encoder.final_layer_norm.weight
  decoder.block.<i>.layer.0.SelfAttention.*, layer.1.EncDecAttention.*, layer.2.DenseReluDense.*
  decoder.final_layer_norm.weight

Plain-T5 names (`encoder.block.<i>.layer...`, what `load_t5` receives from a HF checkpoint,
src/model/gram.py:162-165) are accepted as well.
"""
from __future__ import annotations

import math
import re
from typing import Dict

import numpy as np
import torch

from .config import GramConfig

_ENC = re.compile(r"^encoder\.(?:encoder\.)?block\.(\d+)\.(?:module\.)?layer\.(\d)\.(.+)$")
_DEC = re.compile(r"^decoder\.block\.(\d+)\.layer\.(\d)\.(.+)$")

_ATTN = {"q.weight": "q", "k.weight": "k", "v.weight": "v", "o.weight": "o",
         "relative_attention_bias.weight": "rel_bias"}
_XATTN = {"q.weight": "cq", "k.weight": "ck", "v.weight": "cv", "o.weight": "co"}
_FF = {"wi.weight": "wi", "wo.weight": "wo"}


def canonical_name(key: str):
    """Reference state-dict key -> canonical name (None for aliases / keys the path ignores)."""
    if key == "shared.weight":
        return "shared"
    if key == "lm_head.weight":
        return "lm_head"
    if key in ("position_embedding.weight", "encoder.position_embedding.weight"):
        return "pos_emb"
    if key in ("encoder.encoder.final_layer_norm.weight", "encoder.final_layer_norm.weight"):
        return "enc.final_ln"
    if key == "decoder.final_layer_norm.weight":
        return "dec.final_ln"
    m = _ENC.match(key)
    if m:
        i, sub, rest = int(m.group(1)), int(m.group(2)), m.group(3)
        if sub == 0 and rest.startswith("SelfAttention."):
            w = _ATTN.get(rest[len("SelfAttention."):])
            return f"enc.{i}.{w}" if w else None
        if rest == "layer_norm.weight":
            return f"enc.{i}.ln{sub}"
        if sub == 1 and rest.startswith("DenseReluDense."):
            w = _FF.get(rest[len("DenseReluDense."):])
            return f"enc.{i}.{w}" if w else None
        return None
    m = _DEC.match(key)
    if m:
        i, sub, rest = int(m.group(1)), int(m.group(2)), m.group(3)
        if sub == 0 and rest.startswith("SelfAttention."):
            w = _ATTN.get(rest[len("SelfAttention."):])
            return f"dec.{i}.{w}" if w else None
        if sub == 1 and rest.startswith("EncDecAttention."):
            w = _XATTN.get(rest[len("EncDecAttention."):])
            return f"dec.{i}.{w}" if w else None
        if rest == "layer_norm.weight":
            return f"dec.{i}.ln{sub}"
        if sub == 2 and rest.startswith("DenseReluDense."):
            w = _FF.get(rest[len("DenseReluDense."):])
            return f"dec.{i}.{w}" if w else None
        return None
    return None


def canonicalize(state_dict) -> Dict[str, np.ndarray]:
    """Map a reference-named state dict to {canonical name: contiguous fp32 ndarray}."""
    out: Dict[str, np.ndarray] = {}
    for key, val in state_dict.items():
        name = canonical_name(key)
        if name is None:
            continue
        if isinstance(val, torch.Tensor):
            val = val.detach().to(torch.float32).cpu().numpy()
        arr = np.ascontiguousarray(val, dtype=np.float32)
        if name in out and name != "pos_emb" and out[name].shape != arr.shape:
            raise ValueError(f"conflicting shapes for {name}")
        out[name] = arr
    if "lm_head" not in out and "shared" in out:
        out["lm_head"] = out["shared"]            # tied head
    return out


def relative_position_buckets(cfg: GramConfig, max_seq_len: int, max_length: int):
    """Bucket index tables for the two attention flavours, computed with the same torch fp32 ops as
    reference `src/model/gram_t5_modeling.py:397-450` (`_relative_position_bucket`): the
    `log(...)/log(...)*(...)` followed by `.to(long)` truncation is rounding-sensitive at bucket
    edges, so it is NOT re-derived in integer or double arithmetic.

    Returns (enc int32[2*max_seq_len-1] indexed by (mem - ctx) + max_seq_len - 1,
             dec int32[max_length] indexed by ctx - mem >= 0)."""
    nb, md = cfg.relative_attention_num_buckets, cfg.relative_attention_max_distance

    def bucket(rel: torch.Tensor, bidirectional: bool) -> torch.Tensor:
        n = nb
        res = torch.zeros_like(rel)
        if bidirectional:
            n //= 2
            res = res + (rel > 0).to(torch.long) * n
            rel = torch.abs(rel)
        else:
            rel = -torch.min(rel, torch.zeros_like(rel))
        max_exact = n // 2
        small = rel < max_exact
        large = max_exact + (torch.log(rel.float() / max_exact) / math.log(md / max_exact)
                             * (n - max_exact)).to(torch.long)
        large = torch.min(large, torch.full_like(large, n - 1))
        return res + torch.where(small, rel, large)

    enc_rel = torch.arange(-(max_seq_len - 1), max_seq_len, dtype=torch.long)
    dec_rel = -torch.arange(0, max_length, dtype=torch.long)
    enc = bucket(enc_rel, True).to(torch.int32).numpy()
    dec = bucket(dec_rel, False).to(torch.int32).numpy()
    return np.ascontiguousarray(enc), np.ascontiguousarray(dec)
