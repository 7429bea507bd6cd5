"""gram_b200 -- B200-native implementation of GRAM's inference/scoring hot path.

Host-side mirror of the reference interface (`GRAM.generate`, `Trie`/`prefix_allowed_tokens_fn`,
`evaluate`, the runner's `test_dataset_task`) over `libgram_b200.so` (C ABI: include/gram_b200.h).
"""
from .config import GramConfig
from .generation_trie import Trie, prefix_allowed_tokens_fn
from .model import GRAM, create_model

__all__ = ["GramConfig", "Trie", "prefix_allowed_tokens_fn", "GRAM", "create_model"]
