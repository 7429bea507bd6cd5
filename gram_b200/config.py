"""Model hyper-parameters of the GRAM hot path.

Mirrors the attributes the reference reads from its `T5Config` object
(reference `src/model/gram_t5_config.py:85-105`) plus the three GRAM additions the entry point
attaches to it (`src/main_generative_gram.py:67-70`: `max_seq_len`, `max_item_num`,
`use_position_embedding`).
"""
from __future__ import annotations

from dataclasses import dataclass, asdict


@dataclass
class GramConfig:
    vocab_size: int = 32128
    d_model: int = 512
    d_kv: int = 64
    d_ff: int = 2048
    num_layers: int = 6
    num_decoder_layers: int = 6
    num_heads: int = 8
    relative_attention_num_buckets: int = 32
    relative_attention_max_distance: int = 128
    layer_norm_epsilon: float = 1e-6
    pad_token_id: int = 0
    eos_token_id: int = 1
    decoder_start_token_id: int = 0
    tie_word_embeddings: bool = True
    # GRAM additions
    max_seq_len: int = 128          # item_prompt_max_len (L)
    max_item_num: int = 20          # max_his; position table has max_item_num + 1 rows
    use_position_embedding: bool = True

    @property
    def inner_dim(self) -> int:
        return self.num_heads * self.d_kv

    def to_dict(self):
        return asdict(self)

    @staticmethod
    def t5_small(**kw) -> "GramConfig":
        return GramConfig(**kw)

    @staticmethod
    def t5_base(**kw) -> "GramConfig":
        base = dict(d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_decoder_layers=12, num_heads=12)
        base.update(kw)
        return GramConfig(**base)

    @staticmethod
    def tiny(**kw) -> "GramConfig":
        """A small shape for fast CPU tests (same structure, every code path exercised)."""
        base = dict(vocab_size=384, d_model=64, d_kv=16, d_ff=128, num_layers=2, num_decoder_layers=2,
                    num_heads=4, max_seq_len=16, max_item_num=4)
        base.update(kw)
        return GramConfig(**base)

    @staticmethod
    def from_hf(cfg) -> "GramConfig":
        """Build from a transformers-style config object (duck-typed)."""
        g = lambda name, default=None: getattr(cfg, name, default)  # noqa: E731
        return GramConfig(
            vocab_size=g("vocab_size"), d_model=g("d_model"), d_kv=g("d_kv"), d_ff=g("d_ff"),
            num_layers=g("num_layers"), num_decoder_layers=g("num_decoder_layers") or g("num_layers"),
            num_heads=g("num_heads"),
            relative_attention_num_buckets=g("relative_attention_num_buckets", 32),
            relative_attention_max_distance=g("relative_attention_max_distance", 128),
            layer_norm_epsilon=g("layer_norm_epsilon", 1e-6),
            pad_token_id=g("pad_token_id", 0), eos_token_id=g("eos_token_id", 1),
            decoder_start_token_id=g("decoder_start_token_id", 0) or 0,
            tie_word_embeddings=bool(g("tie_word_embeddings", True)),
            max_seq_len=g("max_seq_len", 128), max_item_num=g("max_item_num", 20),
            use_position_embedding=bool(g("use_position_embedding", True)),
        )
