"""ctypes binding of `libgram_b200.so` (declared in `include/gram_b200.h`).

There is no CPU fallback: if the shared library is missing, fails to load, or no sm_100 device is
visible, the product path raises.  The library is built in-tree by `__graft_entry__.build()`
(`python __graft_entry__.py`).
"""
from __future__ import annotations

import ctypes as C
import os

# GRAM_B200_LIB selects another build of the same library (A/B timing of two builds inside one process launch each)
_LIB_PATH = os.environ.get("GRAM_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib",
                                                            "libgram_b200.so")

GRAM_DTYPE_F32, GRAM_DTYPE_BF16 = 0, 1
GRAM_FLAG_SIMT_GEMM, GRAM_FLAG_KEEP_LOGITS, GRAM_FLAG_SIMT_ATTN, GRAM_FLAG_MMA_ENC_ATTN, GRAM_FLAG_GEMM_1CTA = 1, 2, 4, 8, 16
GRAM_FLAG_ALL_ROWS = 32
GRAM_FLAG_UNFUSED_NORM = 64
GRAM_FLAG_UNFUSED_HEAD = 128
GRAM_FLAG_ENC_CHAIN = 256
GRAM_FLAG_NO_DEC_CHAIN = 1024
GRAM_FLAG_MMA_LONG_ATTN = 2048
GRAM_FLAG_CUDA_GRAPH = 4096
GRAM_FLAG_XATTN_PER_ITEM = 8192
GRAM_FLAG_NO_L2_HINTS = 512
GRAM_FLAG_FP32_RESID = 16384
K_CLASSES = ["gemm_enc", "enc_attn", "gemm_kv", "gemm_dec", "cross_attn", "lm_head", "beam", "other", "self_attn",
             "norm_enc", "norm_dec"]
GRAM_K_COUNT = len(K_CLASSES)

# every symbol include/gram_b200.h declares (tests check the .so exports all of them)
EXPORTED_SYMBOLS = [
    "gram_create", "gram_destroy", "gram_last_error", "gram_version", "gram_load_weight",
    "gram_set_rel_buckets", "gram_finalize_weights", "gram_set_trie", "gram_encode", "gram_generate",
    "gram_get_memory", "gram_decoder_logits", "gram_get_step_taps", "gram_get_stats",
    "gram_profile_begin", "gram_profile_end", "gram_op_gemm", "gram_op_gemm_norm", "gram_op_cross_attention",
    "gram_cache_items", "gram_encode_cached", "gram_check_errors", "gram_op_lse_head", "gram_op_enc_chain",
]


class GramConfigC(C.Structure):
    _fields_ = [
        ("vocab_size", C.c_int32), ("d_model", C.c_int32), ("d_kv", C.c_int32), ("d_ff", C.c_int32),
        ("num_layers", C.c_int32), ("num_decoder_layers", C.c_int32), ("num_heads", C.c_int32),
        ("rel_buckets", C.c_int32), ("rel_max_distance", C.c_int32),
        ("ln_eps", C.c_float),
        ("pad_id", C.c_int32), ("eos_id", C.c_int32), ("start_id", C.c_int32),
        ("tie_word_embeddings", C.c_int32),
        ("n_positions", C.c_int32),
        ("dtype", C.c_int32),
        ("device", C.c_int32),
        ("max_users", C.c_int32), ("max_passages", C.c_int32), ("max_seq_len", C.c_int32),
        ("max_beams", C.c_int32), ("max_length", C.c_int32),
        ("max_tokens", C.c_int64),
        ("flags", C.c_int32),
    ]


class GramStatsC(C.Structure):
    _fields_ = [("launches", C.c_int64), ("packed_tokens", C.c_int64), ("kv_bytes", C.c_int64),
                ("workspace_bytes", C.c_int64), ("decoded_rows", C.c_int64), ("kv_tokens_read", C.c_int64)]


class GramLibraryError(RuntimeError):
    pass


_lib = None


def lib_path() -> str:
    return _LIB_PATH


def load_library():
    """Load the C-ABI library; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise GramLibraryError(
            f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(gram_b200 has no CPU or eager-PyTorch fallback)")
    lib = C.CDLL(_LIB_PATH)
    vp, i32, i64p, u8p, f32p, i32p = C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    lib.gram_create.argtypes = [C.POINTER(GramConfigC), C.POINTER(vp)]
    lib.gram_create.restype = C.c_int
    lib.gram_destroy.argtypes = [vp]
    lib.gram_destroy.restype = None
    lib.gram_last_error.argtypes = [vp]
    lib.gram_last_error.restype = C.c_char_p
    lib.gram_version.argtypes = []
    lib.gram_version.restype = C.c_char_p
    lib.gram_load_weight.argtypes = [vp, C.c_char_p, f32p, C.POINTER(C.c_int64), i32]
    lib.gram_load_weight.restype = C.c_int
    lib.gram_set_rel_buckets.argtypes = [vp, i32p, i32, i32p, i32]
    lib.gram_set_rel_buckets.restype = C.c_int
    lib.gram_finalize_weights.argtypes = [vp]
    lib.gram_finalize_weights.restype = C.c_int
    lib.gram_set_trie.argtypes = [vp, i32p, i32p, i32p, i32, i32, i32]
    lib.gram_set_trie.restype = C.c_int
    lib.gram_encode.argtypes = [vp, i64p, u8p, i32, i32, i32, vp]
    lib.gram_encode.restype = C.c_int
    lib.gram_cache_items.argtypes = [vp, i64p, u8p, i32, i32, vp]
    lib.gram_cache_items.restype = C.c_int
    lib.gram_encode_cached.argtypes = [vp, i64p, u8p, i32p, i32, i32, i32, vp]
    lib.gram_encode_cached.restype = C.c_int
    lib.gram_check_errors.argtypes = [vp, vp]
    lib.gram_check_errors.restype = C.c_int
    lib.gram_generate.argtypes = [vp, i64p, u8p, i32, i32, i32, i32, i32, i32, C.c_void_p, i64p, i32p, f32p, vp]
    lib.gram_generate.restype = C.c_int
    lib.gram_get_memory.argtypes = [vp, f32p, vp]
    lib.gram_get_memory.restype = C.c_int
    lib.gram_decoder_logits.argtypes = [vp, i64p, i32, f32p, vp]
    lib.gram_decoder_logits.restype = C.c_int
    lib.gram_get_step_taps.argtypes = [vp, f32p, f32p, i32p, C.POINTER(C.c_int32)]
    lib.gram_get_step_taps.restype = C.c_int
    lib.gram_get_stats.argtypes = [vp, C.POINTER(GramStatsC)]
    lib.gram_get_stats.restype = C.c_int
    lib.gram_profile_begin.argtypes = [vp, C.c_uint32]
    lib.gram_profile_begin.restype = C.c_int
    lib.gram_profile_end.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int64)]
    lib.gram_profile_end.restype = C.c_int
    lib.gram_op_gemm.argtypes = [i32, i32, i32, i32, vp, vp, vp, i32, i32, i32, vp]
    lib.gram_op_gemm.restype = C.c_int
    lib.gram_op_gemm_norm.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, C.c_float, i32, i32, i32, vp]
    lib.gram_op_gemm_norm.restype = C.c_int
    lib.gram_op_cross_attention.argtypes = [i32, i32, i32, vp, vp, i32, i32p, u8p, vp, i32, i32, i32, i32, vp]
    lib.gram_op_cross_attention.restype = C.c_int
    lib.gram_op_enc_chain.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int64, vp, vp, C.c_float, i32, i32, i32, i32, i32, vp, vp]
    lib.gram_op_enc_chain.restype = C.c_int
    lib.gram_op_lse_head.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, vp]
    lib.gram_op_lse_head.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, handle=None, what: str = ""):
    """Map C return codes onto the exceptions the reference stack raises."""
    if rc == 0:
        return
    lib = load_library()
    msg = lib.gram_last_error(handle)
    msg = msg.decode() if msg else ""
    text = f"{what}: {msg}" if what else msg
    if rc == 1:
        raise ValueError(text)
    raise GramLibraryError(f"[gram_b200 rc={rc}] {text}")
