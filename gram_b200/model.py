"""`GRAM` -- the reference's model surface over the B200 engine.

Drop-in for the inference-side API of reference `src/model/gram.py` (`GRAM`, `EncoderWrapper`
attributes) and `src/model/__init__.py:9-24` (`create_model`):

  GRAM(config)                      config: the reference's T5Config-like object or a GramConfig
  .load_t5(state_dict) / .load_state_dict(state_dict)     reference key names
  .generate(input_ids[B,N,L], attention_mask[B,N,L], max_length, prefix_allowed_tokens_fn=...,
            num_beams=K, num_return_sequences=R, output_scores=True, return_dict_in_generate=True,
            length_penalty=lp)      -> {"sequences": int64 [B*R, W], "sequences_scores": fp32 [B*R]}
  .forward(input_ids, attention_mask, decoder_input_ids=...|labels=...)  -> .logits (teacher forced)
  .encoder.n_passages, .position_embedding, .config, .module (so `.module.generate` works as in
  src/runner/distributed_runner_gram.py:775)

All arithmetic happens in `libgram_b200.so` (hand-written sm_100a kernels) through the C ABI in
`include/gram_b200.h`.  There is no eager-PyTorch or CPU fallback: without the library or without a
B200 the calls raise.  PyTorch is used for tensor storage, the current CUDA stream and, in the
runner, `torch.distributed`.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from . import _cabi
from .config import GramConfig
from .weights import canonicalize, relative_position_buckets


class GenerateOutput(dict):
    """dict with attribute access, standing in for HF `BeamSearchEncoderDecoderOutput`."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class _Holder:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class GRAM(torch.nn.Module):
    """An `nn.Module` like the reference's `GRAM` (src/model/gram.py:22), so the runners' wrapping code works:
    `DistributedDataParallel(model, device_ids=[rank])` (src/runner/distributed_runner_gram.py:47-52) needs a module with
    at least one trainable parameter -- `_ddp_anchor` is that parameter and nothing else; the weights live in the engine."""

    main_input_name = "input_ids"

    def __init__(self, config, dtype: str = None, device=None, max_users: int = 1, max_beams: int = None,
                 max_length: int = 16, max_passages: int = None, max_seq_len: int = None, flags: int = 0):
        super().__init__()
        self._ddp_anchor = torch.nn.Parameter(torch.zeros(1))
        self.config = config
        self.gcfg = config if isinstance(config, GramConfig) else GramConfig.from_hf(config)
        g = self.gcfg
        self.max_seq_len = g.max_seq_len
        self.max_item_num = g.max_item_num
        self.use_position_embedding = g.use_position_embedding
        dtype = dtype or getattr(config, "gram_b200_dtype", "bf16")
        if dtype not in ("fp32", "bf16"):
            raise ValueError("dtype must be 'fp32' or 'bf16'")
        self.dtype = dtype
        self.flags = int(flags)
        self._device = torch.device(device) if device is not None else None
        self._cap = dict(max_users=max_users, max_beams=max_beams or 1, max_length=max_length,
                         max_passages=max_passages or (g.max_item_num + 1),
                         max_seq_len=max_seq_len or g.max_seq_len, max_tokens=0)
        self._weights = None            # canonical name -> fp32 ndarray (host copy, reloaded on regrow)
        self._handle = None
        self._trie_key = None
        self._lib = None
        self.encoder = _Holder(n_passages=None, position_embedding=None, main_input_name="input_ids")
        self.position_embedding = None
        self.training = False           # inference only: the module is born in eval mode
        self.user_limit = 256           # auto-grow max_users up to this; larger batches are chunked

    # ---- nn.Module-ish conveniences the runner touches -------------------------------------------
    @property
    def module(self):
        """`model_rec.module.generate(...)` (distributed_runner_gram.py:775) also works on the unwrapped model."""
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("gram_b200 implements the inference/scoring path only")
        return super().train(False)

    def to(self, device=None, *args, **kwargs):
        if device is not None and not isinstance(device, torch.dtype):
            self._device = torch.device(device)
            super().to(self._device)
        return self

    def cuda(self, device=None):
        return self.to(f"cuda:{device}" if isinstance(device, int) else (device or "cuda"))

    @property
    def device(self):
        if self._device is None:
            self._device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        if self._device.type == "cuda" and self._device.index is None:
            self._device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        return self._device

    # ---- weights ----------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = False):
        canon = canonicalize(state_dict)
        if self._weights is None:
            self._weights = {}
        self._weights.update(canon)
        if "pos_emb" in self._weights:
            w = torch.from_numpy(self._weights["pos_emb"])
            self.position_embedding = SimpleNamespace(weight=w)
            self.encoder.position_embedding = self.position_embedding
        self._destroy_handle()
        return SimpleNamespace(missing_keys=[], unexpected_keys=[])

    def load_t5(self, state_dict):
        """reference `GRAM.load_t5` (src/model/gram.py:162-165): load a plain-T5 state dict; the
        passage-position table keeps its current value (N(0, 0.02) init, src/model/gram.py:32-33,
        unless already loaded)."""
        self.load_state_dict(state_dict, strict=False)
        if self.gcfg.use_position_embedding and "pos_emb" not in self._weights:
            from .synth import pseudo_normal
            pe = pseudo_normal((self.gcfg.max_item_num + 1, self.gcfg.d_model), 0.02, 20250101)
            self.load_state_dict({"position_embedding.weight": pe})

    def state_dict(self, *args, **kwargs):
        return dict(self._weights or {})

    # ---- engine lifetime ----------------------------------------------------------------------------
    def _destroy_handle(self):
        if self._handle is not None and self._lib is not None:
            self._lib.gram_destroy(self._handle)
        self._handle = None
        self._trie_key = None

    def __del__(self):
        try:
            self._destroy_handle()
        except Exception:
            pass

    def configure(self, **caps):
        """Set capacities (max_users, max_beams, max_length, max_passages, max_seq_len, max_tokens) up front so the
        engine is created once with the right workspace.  max_tokens (0 = every passage full) sizes the encoder /
        K-V workspace by VALID tokens per call, so a larger user batch fits when histories are short."""
        changed = False
        for k, v in caps.items():
            if k not in self._cap:
                raise KeyError(k)
            if v is not None and v != self._cap[k]:
                self._cap[k] = int(v)
                changed = True
        if changed:
            self._destroy_handle()
        return self

    def _ensure(self, B=1, N=1, L=1, K=1, max_length=2):
        grow = dict(max_users=B, max_passages=N, max_seq_len=L, max_beams=K, max_length=max_length)
        need = any(grow[k] > self._cap[k] for k in grow)
        if need:
            for k in grow:
                self._cap[k] = max(self._cap[k], grow[k])
            self._destroy_handle()
        if self._handle is not None:
            return
        if self._weights is None:
            raise RuntimeError("GRAM: no weights loaded (call load_state_dict / load_t5 first)")
        if not torch.cuda.is_available():
            raise _cabi.GramLibraryError("GRAM: no CUDA device -- gram_b200 has no CPU fallback")
        lib = _cabi.load_library()
        self._lib = lib
        g, cap = self.gcfg, self._cap
        cc = _cabi.GramConfigC(
            vocab_size=g.vocab_size, d_model=g.d_model, d_kv=g.d_kv, d_ff=g.d_ff, num_layers=g.num_layers,
            num_decoder_layers=g.num_decoder_layers, num_heads=g.num_heads,
            rel_buckets=g.relative_attention_num_buckets, rel_max_distance=g.relative_attention_max_distance,
            ln_eps=g.layer_norm_epsilon, pad_id=g.pad_token_id, eos_id=g.eos_token_id,
            start_id=g.decoder_start_token_id, tie_word_embeddings=int(g.tie_word_embeddings),
            n_positions=(g.max_item_num + 1) if g.use_position_embedding else 0,
            dtype=_cabi.GRAM_DTYPE_F32 if self.dtype == "fp32" else _cabi.GRAM_DTYPE_BF16,
            device=self.device.index or 0,
            max_users=cap["max_users"], max_passages=cap["max_passages"], max_seq_len=cap["max_seq_len"],
            max_beams=cap["max_beams"], max_length=cap["max_length"], max_tokens=cap["max_tokens"], flags=self.flags)
        hp = C.c_void_p()
        _cabi.check(lib.gram_create(C.byref(cc), C.byref(hp)), None, "gram_create")
        self._handle = hp
        for name, arr in self._weights.items():
            arr = np.ascontiguousarray(arr, dtype=np.float32)
            shape = (C.c_int64 * arr.ndim)(*arr.shape)
            _cabi.check(lib.gram_load_weight(hp, name.encode(), C.c_void_p(arr.ctypes.data), shape, arr.ndim),
                        hp, f"gram_load_weight({name})")
        enc_b, dec_b = relative_position_buckets(g, cap["max_seq_len"], cap["max_length"])
        _cabi.check(lib.gram_set_rel_buckets(hp, C.c_void_p(enc_b.ctypes.data), len(enc_b),
                                             C.c_void_p(dec_b.ctypes.data), len(dec_b)), hp, "gram_set_rel_buckets")
        _cabi.check(lib.gram_finalize_weights(hp), hp, "gram_finalize_weights")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _set_trie(self, trie):
        key = (id(trie), getattr(trie, "_version", 0), len(trie))
        if self._trie_key == key:
            return
        csr = trie.to_csr(self.gcfg.decoder_start_token_id)
        self._csr_keepalive = csr
        _cabi.check(self._lib.gram_set_trie(
            self._handle, C.c_void_p(csr["child_offsets"].ctypes.data), C.c_void_p(csr["child_tokens"].ctypes.data),
            C.c_void_p(csr["child_nodes"].ctypes.data), csr["n_nodes"], csr["n_edges"], csr["root_node"]),
            self._handle, "gram_set_trie")
        self._trie_key = key

    @staticmethod
    def _prep_inputs(input_ids, attention_mask):
        if input_ids.dim() != 3 or attention_mask.shape != input_ids.shape:
            raise ValueError("input_ids / attention_mask must both be [B, N, L]")
        ids = input_ids.to(torch.int64).contiguous()
        mask = attention_mask
        if mask.dtype != torch.bool:
            mask = mask != 0
        mask = mask.contiguous().view(torch.uint8)
        return ids, mask

    # ---- hot path -------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, input_ids, attention_mask, max_length, prefix_allowed_tokens_fn=None, num_beams: int = 1,
                 num_return_sequences: int = 1, output_scores: bool = False, return_dict_in_generate: bool = False,
                 length_penalty: float = 1.0, **kwargs):
        trie = getattr(prefix_allowed_tokens_fn, "candidate_trie", None)
        if trie is None:
            raise NotImplementedError(
                "GRAM.generate needs `prefix_allowed_tokens_fn` built by gram_b200.generation_trie."
                "prefix_allowed_tokens_fn(trie): the trie is walked on the device; arbitrary Python "
                "callables would need a per-step host round trip, which this path does not have")
        if num_return_sequences > num_beams:
            raise ValueError("`num_return_sequences` has to be smaller or equal to `num_beams`.")
        ids, mask = self._prep_inputs(input_ids, attention_mask)
        B, N, L = ids.shape
        self.encoder.n_passages = N
        K, R = int(num_beams), int(num_return_sequences)
        max_length = int(max_length)
        self._ensure(min(B, max(self._cap["max_users"], self.user_limit)), N, L, K, max_length)
        self._set_trie(trie)
        len_pow = (C.c_double * (max_length + 1))(*[float(c) ** float(length_penalty) if c > 0 else 1.0
                                                    for c in range(max_length + 1)])
        cap_u = self._cap["max_users"]
        seq_parts, score_parts, widths = [], [], []
        for b0 in range(0, B, cap_u):
            b1 = min(B, b0 + cap_u)
            nb = b1 - b0
            out_seq, out_scores, out_width = self._out_buffers(nb * R, max_length, ids.device, len(seq_parts))
            _cabi.check(self._lib.gram_generate(
                self._handle, _ptr(ids[b0:b1]), _ptr(mask[b0:b1]), nb, N, L, K, R, max_length, len_pow,
                _ptr(out_seq), _ptr(out_width), _ptr(out_scores), self._stream()), self._handle, "gram_generate")
            seq_parts.append(out_seq)
            score_parts.append(out_scores)
            widths.append(out_width)
        return self._finish_generate(seq_parts, score_parts, widths, ids.device, return_dict_in_generate, "gram_generate")

    def _out_buffers(self, rows, max_length, device, slot=0):
        """Result buffers of one C-ABI call.  Host callers get PINNED buffers (kept per shape and reused: the library's
        device-to-host copies then run at full PCIe rate and nothing is allocated per call); the finished result is copied
        out of them in `_finish_generate`."""
        if device.type == "cuda":
            return (torch.empty((rows, max_length), dtype=torch.int64, device=device),
                    torch.empty((rows,), dtype=torch.float32, device=device), torch.empty((1,), dtype=torch.int32, device=device))
        cache = self.__dict__.setdefault("_pinned_out", {})
        key = (rows, max_length, slot)                  # one buffer per chunk of a call: chunks are concatenated afterwards
        buf = cache.get(key)
        if buf is None:
            pin = torch.cuda.is_available()
            buf = (torch.empty((rows, max_length), dtype=torch.int64, pin_memory=pin),
                   torch.empty((rows,), dtype=torch.float32, pin_memory=pin), torch.empty((1,), dtype=torch.int32, pin_memory=pin))
            if len(cache) > 16:
                cache.clear()
            cache[key] = buf
        return buf

    def _finish_generate(self, seq_parts, score_parts, widths, device, return_dict, what):
        if device.type == "cuda":
            # device outputs: the calls above did not synchronise; one check (it synchronises the stream) surfaces the
            # sticky device-side input errors, after which the widths are plain reads
            _cabi.check(self._lib.gram_check_errors(self._handle, self._stream()), self._handle, what)
        # host outputs: gram_generate synchronised and checked the error flag itself before returning
        width = max(int(w[0]) for w in widths)
        if len(seq_parts) == 1:
            sequences = seq_parts[0][:, :width].clone() if device.type != "cuda" else seq_parts[0][:, :width].contiguous()
            scores = score_parts[0].clone() if device.type != "cuda" else score_parts[0]
        else:
            sequences = torch.cat([p[:, :width] for p in seq_parts], 0)
            scores = torch.cat(score_parts, 0)
        if not return_dict:
            return sequences
        # `scores` / `beam_indices`: the reference asks for them (output_scores=True) but never reads them; producing the
        # per-step [B*K, V] score matrices is exactly what the fused head avoids, so they stay None
        return GenerateOutput(sequences=sequences, sequences_scores=scores, scores=None, beam_indices=None)

    @torch.no_grad()
    def generate_into(self, input_ids, attention_mask, max_length, trie, num_beams, num_return_sequences,
                      length_penalty, out_seq, out_scores, out_width):
        """`generate` without the Python-side post-processing: one C-ABI call writing into caller
        tensors (`out_seq` int64 [B*R, max_length], `out_scores` fp32 [B*R], `out_width` int32 [1]).
        With device tensors nothing synchronises -- the bench's kernel-only loop uses this."""
        ids, mask = self._prep_inputs(input_ids, attention_mask)
        B, N, L = ids.shape
        K, R = int(num_beams), int(num_return_sequences)
        self._ensure(B, N, L, K, int(max_length))
        self._set_trie(trie)
        key = (int(max_length), float(length_penalty))
        if getattr(self, "_len_pow_key", None) != key:
            self._len_pow = (C.c_double * (max_length + 1))(*[float(c) ** float(length_penalty) if c > 0 else 1.0
                                                              for c in range(max_length + 1)])
            self._len_pow_key = key
        _cabi.check(self._lib.gram_generate(
            self._handle, _ptr(ids), _ptr(mask), B, N, L, K, R, int(max_length), self._len_pow,
            _ptr(out_seq), _ptr(out_width), _ptr(out_scores), self._stream()), self._handle, "gram_generate")

    # ---- per-item encoder-state cache (no counterpart in the reference) -------------------
    @torch.no_grad()
    def cache_items(self, item_ids, item_mask):
        """Encode every item passage once (`item_ids` int64 / `item_mask` bool `[n_items, L]`) and keep the encoder
        states on the device; `generate_cached` then only encodes the user prompts.  Exact: item passages do not
        depend on the user (src/utils/indexing.py:209-211,315-320) and the position row is added after the encoder
        (src/model/gram.py:238-249)."""
        ids = torch.as_tensor(item_ids).to(torch.int64).contiguous()
        mask = torch.as_tensor(item_mask)
        if ids.dim() != 2 or mask.shape != ids.shape:
            raise ValueError("item_ids / item_mask must both be [n_items, L]")
        if mask.dtype != torch.bool:
            mask = mask != 0
        self._item_table = (ids, mask.contiguous().view(torch.uint8))
        self._item_table_handle = None

    def _ensure_items(self):
        if getattr(self, "_item_table", None) is None:
            raise RuntimeError("generate_cached: no item table (call cache_items first)")
        if self._item_table_handle is not self._handle:          # first use, or the engine was re-created
            ids, mask = self._item_table
            _cabi.check(self._lib.gram_cache_items(self._handle, _ptr(ids), _ptr(mask), ids.shape[0], ids.shape[1],
                                                   self._stream()), self._handle, "gram_cache_items")
            self._item_table_handle = self._handle

    @torch.no_grad()
    def generate_cached(self, prompt_ids, prompt_mask, item_index, max_length, prefix_allowed_tokens_fn=None,
                        num_beams: int = 1, num_return_sequences: int = 1, return_dict_in_generate: bool = False,
                        length_penalty: float = 1.0, **kwargs):
        """`generate` for users given as (prompt passage `[B, L]`, history item indices `[B, NI]`, -1 = none) over
        the table built by `cache_items`.  Same outputs, bit for bit, as `generate` on the `[B, 1+NI, L]` tensors
        whose passage 1+j is the cached passage of `item_index[b][j]`."""
        trie = getattr(prefix_allowed_tokens_fn, "candidate_trie", None)
        if trie is None:
            raise NotImplementedError("generate_cached needs `prefix_allowed_tokens_fn` built by "
                                      "gram_b200.generation_trie.prefix_allowed_tokens_fn(trie)")
        if num_return_sequences > num_beams:
            raise ValueError("`num_return_sequences` has to be smaller or equal to `num_beams`.")
        ids = torch.as_tensor(prompt_ids).to(torch.int64).contiguous()
        mask = torch.as_tensor(prompt_mask)
        if mask.dtype != torch.bool:
            mask = mask != 0
        mask = mask.contiguous().view(torch.uint8)
        items = torch.as_tensor(item_index).to(torch.int32).contiguous()
        if ids.dim() != 2 or mask.shape != ids.shape or items.dim() != 2 or items.shape[0] != ids.shape[0]:
            raise ValueError("prompt_ids / prompt_mask must be [B, L] and item_index [B, NI]")
        B, L = ids.shape
        NI = items.shape[1]
        K, R = int(num_beams), int(num_return_sequences)
        max_length = int(max_length)
        self.encoder.n_passages = NI + 1
        self._ensure(min(B, max(self._cap["max_users"], self.user_limit)), NI + 1, L, K, max_length)
        self._set_trie(trie)
        self._ensure_items()
        len_pow = (C.c_double * (max_length + 1))(*[float(c) ** float(length_penalty) if c > 0 else 1.0
                                                    for c in range(max_length + 1)])
        cap_u = self._cap["max_users"]
        seq_parts, score_parts, widths = [], [], []
        for b0 in range(0, B, cap_u):
            b1 = min(B, b0 + cap_u)
            nb = b1 - b0
            out_seq, out_scores, out_width = self._out_buffers(nb * R, max_length, ids.device, len(seq_parts))
            _cabi.check(self._lib.gram_encode_cached(self._handle, _ptr(ids[b0:b1]), _ptr(mask[b0:b1]), _ptr(items[b0:b1]),
                                                     nb, NI, L, self._stream()), self._handle, "gram_encode_cached")
            _cabi.check(self._lib.gram_generate(
                self._handle, None, None, nb, NI + 1, L, K, R, max_length, len_pow,
                _ptr(out_seq), _ptr(out_width), _ptr(out_scores), self._stream()), self._handle, "gram_generate")
            seq_parts.append(out_seq)
            score_parts.append(out_scores)
            widths.append(out_width)
        return self._finish_generate(seq_parts, score_parts, widths, ids.device, return_dict_in_generate, "gram_generate_cached")

    @torch.no_grad()
    def generate_cached_into(self, prompt_ids, prompt_mask, item_index, max_length, trie, num_beams, num_return_sequences,
                             length_penalty, out_seq, out_scores, out_width):
        """`generate_cached` without the Python-side post-processing (see `generate_into`): two C-ABI calls, no
        synchronisation with device tensors."""
        ids = prompt_ids.to(torch.int64).contiguous()
        mask = (prompt_mask if prompt_mask.dtype == torch.bool else prompt_mask != 0).contiguous().view(torch.uint8)
        items = item_index.to(torch.int32).contiguous()
        B, L = ids.shape
        NI = items.shape[1]
        K, R = int(num_beams), int(num_return_sequences)
        self._ensure(B, NI + 1, L, K, int(max_length))
        self._set_trie(trie)
        self._ensure_items()
        key = (int(max_length), float(length_penalty))
        if getattr(self, "_len_pow_key", None) != key:
            self._len_pow = (C.c_double * (max_length + 1))(*[float(c) ** float(length_penalty) if c > 0 else 1.0
                                                              for c in range(max_length + 1)])
            self._len_pow_key = key
        _cabi.check(self._lib.gram_encode_cached(self._handle, _ptr(ids), _ptr(mask), _ptr(items), B, NI, L, self._stream()),
                    self._handle, "gram_encode_cached")
        _cabi.check(self._lib.gram_generate(
            self._handle, None, None, B, NI + 1, L, K, R, int(max_length), self._len_pow,
            _ptr(out_seq), _ptr(out_width), _ptr(out_scores), self._stream()), self._handle, "gram_generate")

    @torch.no_grad()
    def encode_cached(self, prompt_ids, prompt_mask, item_index):
        """Fused memory `[B, (1+NI)*L, d_model]` fp32 of the cached path -- parity tap, compare with `encode`."""
        ids = torch.as_tensor(prompt_ids).to(torch.int64).contiguous()
        mask = torch.as_tensor(prompt_mask)
        mask = (mask if mask.dtype == torch.bool else mask != 0).contiguous().view(torch.uint8)
        items = torch.as_tensor(item_index).to(torch.int32).contiguous()
        B, L = ids.shape
        NI = items.shape[1]
        self._ensure(B, NI + 1, L, 1, 2)
        self._ensure_items()
        _cabi.check(self._lib.gram_encode_cached(self._handle, _ptr(ids), _ptr(mask), _ptr(items), B, NI, L, self._stream()),
                    self._handle, "gram_encode_cached")
        dev = ids.device if ids.is_cuda else self.device
        out = torch.empty((B, (NI + 1) * L, self.gcfg.d_model), dtype=torch.float32, device=dev)
        _cabi.check(self._lib.gram_get_memory(self._handle, _ptr(out), self._stream()), self._handle, "gram_get_memory")
        return out.to(ids.device)

    @torch.no_grad()
    def encode(self, input_ids, attention_mask):
        """Fused FiD memory `[B, N*L, d_model]` fp32 (zeros at skipped positions) -- parity tap."""
        ids, mask = self._prep_inputs(input_ids, attention_mask)
        B, N, L = ids.shape
        self.encoder.n_passages = N
        self._ensure(B, N, L, 1, 2)
        _cabi.check(self._lib.gram_encode(self._handle, _ptr(ids), _ptr(mask), B, N, L, self._stream()),
                    self._handle, "gram_encode")
        dev = ids.device if ids.is_cuda else self.device
        out = torch.empty((B, N * L, self.gcfg.d_model), dtype=torch.float32, device=dev)
        _cabi.check(self._lib.gram_get_memory(self._handle, _ptr(out), self._stream()), self._handle, "gram_get_memory")
        return out.to(ids.device)

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, decoder_input_ids=None, labels=None, **kwargs):
        """Teacher-forced logits `[B, q, V]` through the cached decode step (reference
        `GRAM.forward`, src/model/gram.py:51-69 -> gram_t5.py:118-287)."""
        if decoder_input_ids is None:
            if labels is None:
                raise ValueError("forward needs decoder_input_ids or labels")
            decoder_input_ids = self._shift_right(labels)
        ids, mask = self._prep_inputs(input_ids, attention_mask)
        B, N, L = ids.shape
        q = decoder_input_ids.shape[1]
        self.encoder.n_passages = N
        self._ensure(B, N, L, 1, max(q, 2))
        _cabi.check(self._lib.gram_encode(self._handle, _ptr(ids), _ptr(mask), B, N, L, self._stream()),
                    self._handle, "gram_encode")
        dec = decoder_input_ids.to(torch.int64).contiguous()
        dev = ids.device if ids.is_cuda else self.device
        logits = torch.empty((B, q, self.gcfg.vocab_size), dtype=torch.float32, device=dev)
        _cabi.check(self._lib.gram_decoder_logits(self._handle, _ptr(dec), q, _ptr(logits), self._stream()),
                    self._handle, "gram_decoder_logits")
        logits = logits.to(ids.device)
        loss = None
        if labels is not None:
            loss = torch.nn.functional.cross_entropy(logits.view(-1, logits.size(-1)).float(),
                                                     labels.reshape(-1).to(logits.device), ignore_index=-100)
        return GenerateOutput(loss=loss, logits=logits)

    def _shift_right(self, labels):
        start, pad = self.gcfg.decoder_start_token_id, self.gcfg.pad_token_id
        out = labels.new_zeros(labels.shape)
        out[..., 1:] = labels[..., :-1].clone()
        out[..., 0] = start
        out.masked_fill_(out == -100, pad)
        return out

    def check_errors(self):
        """Raise the sticky device-side error of earlier calls, if any (token id outside the vocabulary, more valid
        tokens than `max_tokens`, candidate overflow ...).  `generate` / `generate_cached` call it themselves; loops over
        `generate_into` / `generate_cached_into` (which never synchronise) call it once at the end."""
        if self._handle is not None:
            _cabi.check(self._lib.gram_check_errors(self._handle, self._stream()), self._handle, "gram_check_errors")

    # ---- measurement helpers (bench.py / tests) -----------------------------------------------------------
    def stats(self):
        st = _cabi.GramStatsC()
        _cabi.check(self._lib.gram_get_stats(self._handle, C.byref(st)), self._handle, "gram_get_stats")
        return dict(launches=st.launches, packed_tokens=st.packed_tokens, kv_bytes=st.kv_bytes,
                    workspace_bytes=st.workspace_bytes, decoded_rows=st.decoded_rows, kv_tokens_read=st.kv_tokens_read)

    def profile_begin(self, classes=None):
        mask = 0
        for c in (classes or _cabi.K_CLASSES):
            mask |= 1 << _cabi.K_CLASSES.index(c)
        _cabi.check(self._lib.gram_profile_begin(self._handle, mask), self._handle, "gram_profile_begin")

    def profile_end(self):
        ms = (C.c_float * _cabi.GRAM_K_COUNT)()
        n = (C.c_int64 * _cabi.GRAM_K_COUNT)()
        _cabi.check(self._lib.gram_profile_end(self._handle, ms, n), self._handle, "gram_profile_end")
        return {c: dict(ms=float(ms[i]), launches=int(n[i])) for i, c in enumerate(_cabi.K_CLASSES)}

    def step_taps(self, B, K):
        """(lse [T, B*K], beam_scores [T, B*K], tokens [T, B*K, cap_len]) of the last generate; needs
        flags=GRAM_FLAG_KEEP_LOGITS."""
        cap_len = self._cap["max_length"]
        n_steps = C.c_int32(0)
        _cabi.check(self._lib.gram_get_step_taps(self._handle, None, None, None, C.byref(n_steps)), self._handle, "taps")
        T = n_steps.value
        lse = np.zeros((T, B * K), dtype=np.float32)
        bsc = np.zeros((T, B * K), dtype=np.float32)
        tok = np.zeros((T, B * K, cap_len), dtype=np.int32)
        _cabi.check(self._lib.gram_get_step_taps(self._handle, C.c_void_p(lse.ctypes.data), C.c_void_p(bsc.ctypes.data),
                                                 C.c_void_p(tok.ctypes.data), C.byref(n_steps)), self._handle, "taps")
        return lse, bsc, tok


def create_model(model_class, config, **kw):
    """reference `src/model/__init__.py:9-24`."""
    if model_class != "gram":
        raise NotImplementedError(f"model class {model_class!r}: only 'gram' is on the hot path")
    return GRAM(config, **kw)
