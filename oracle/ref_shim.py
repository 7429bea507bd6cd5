"""TEST INFRASTRUCTURE ONLY -- loads the real reference (`/root/reference/src`) in THIS container.

The reference is Python/PyTorch written against transformers==4.26.0.  This container ships
transformers 5.5.0, so a handful of import-level shims are needed before `src/model` imports
(SURVEY.md section 8(c)):

  * stub modules `IPython` and `undecorated` (imported at reference `src/model/gram.py:8`,
    `src/model/gram_t5.py:21`, `src/model/gram_t5_modeling.py:23`)
  * stub `transformers.utils.model_parallel_utils` (`gram_t5.py:30`, `gram_t5_modeling.py:45`)
  * `transformers.pytorch_utils.find_pruneable_heads_and_indices` (`gram_t5_modeling.py:33`)
  * `PreTrainedModel.get_head_mask` (used at `gram_t5_modeling.py:1152-1155`)
  * `config.tie_word_embeddings = True` (read at `gram_t5.py:249`)

No reference code is modified or copied.  `/root/reference` does not exist on the GPU box, so this
module is only ever imported by `oracle/make_golden.py` and by the CPU-only pin tests, which skip
when the directory is absent.  Nothing under `gram_b200/` may import it.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GRAM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "model"))


def _install_shims() -> None:
    import torch  # noqa: F401
    import transformers  # noqa: F401

    if "IPython" not in sys.modules:
        m = types.ModuleType("IPython")
        m.embed = lambda *a, **k: None
        sys.modules["IPython"] = m
    if "undecorated" not in sys.modules:
        m = types.ModuleType("undecorated")
        m.undecorated = lambda f: f
        sys.modules["undecorated"] = m
    name = "transformers.utils.model_parallel_utils"
    if name not in sys.modules:
        try:
            importlib.import_module(name)
        except Exception:
            m = types.ModuleType(name)
            m.assert_device_map = lambda *a, **k: None
            m.get_device_map = lambda *a, **k: {}
            sys.modules[name] = m
    import transformers.pytorch_utils as pu

    if not hasattr(pu, "find_pruneable_heads_and_indices"):
        pu.find_pruneable_heads_and_indices = lambda *a, **k: (set(), None)
    from transformers.modeling_utils import PreTrainedModel

    if not hasattr(PreTrainedModel, "get_head_mask"):
        def get_head_mask(self, head_mask, num_hidden_layers, is_attention_chunked=False):
            assert head_mask is None
            return [None] * num_hidden_layers

        PreTrainedModel.get_head_mask = get_head_mask


_ref_pkg = None


def load_reference():
    """Return a namespace with the reference's `model`, `generation_trie`, `evaluate` modules."""
    global _ref_pkg
    if _ref_pkg is not None:
        return _ref_pkg
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_shims()
    src = os.path.join(REFERENCE_ROOT, "src")
    # import `src/model` as a top-level package called `gram_ref_model` without touching sys.path
    # for the other (non-importable) reference packages.
    spec = importlib.util.spec_from_file_location(
        "gram_ref_model", os.path.join(src, "model", "__init__.py"),
        submodule_search_locations=[os.path.join(src, "model")])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["gram_ref_model"] = pkg
    spec.loader.exec_module(pkg)

    def _load_file(modname, path):
        sp = importlib.util.spec_from_file_location(modname, path)
        mod = importlib.util.module_from_spec(sp)
        sys.modules[modname] = mod
        sp.loader.exec_module(mod)
        return mod

    trie = _load_file("gram_ref_generation_trie", os.path.join(src, "utils", "generation_trie.py"))
    evaluate = _load_file("gram_ref_evaluate", os.path.join(src, "utils", "evaluate.py"))
    ns = types.SimpleNamespace(model=pkg, generation_trie=trie, evaluate=evaluate)
    _ref_pkg = ns
    return ns


def make_reference_config(**overrides):
    """A transformers.T5Config carrying the extra attributes the reference entry point sets
    (`src/main_generative_gram.py:61-70`)."""
    from transformers import T5Config

    kw = dict(vocab_size=32128, d_model=512, d_kv=64, d_ff=2048, num_layers=6, num_decoder_layers=6,
              num_heads=8, relative_attention_num_buckets=32, relative_attention_max_distance=128,
              dropout_rate=0.1, layer_norm_epsilon=1e-6, feed_forward_proj="relu",
              pad_token_id=0, eos_token_id=1, decoder_start_token_id=0)
    max_seq_len = overrides.pop("max_seq_len", 128)
    max_item_num = overrides.pop("max_item_num", 20)
    kw.update(overrides)
    cfg = T5Config(**kw)
    cfg.max_seq_len = max_seq_len
    cfg.max_item_num = max_item_num
    cfg.use_position_embedding = True
    cfg.sample_num = 1
    cfg.tie_word_embeddings = True
    cfg.use_cache = True
    return cfg


def build_reference_model(cfg, state_dict):
    """Instantiate the reference GRAM model and load `state_dict` (reference key names)."""
    import torch

    ref = load_reference()
    model = ref.model.create_model("gram", cfg)
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    # lm_head / embed_tokens aliases may be reported missing when tied; everything else must load
    bad = [k for k in missing if not (k.endswith("embed_tokens.weight") or k == "lm_head.weight"
                                      or k == "encoder.position_embedding.weight")]
    assert not bad, bad
    model.eval()
    torch.set_grad_enabled(False)
    return model
