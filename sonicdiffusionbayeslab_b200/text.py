"""Prompt conditioning for the pipelines (``encode_prompt``, /root/reference/src/models.py:139-149).

Once per batch and outside the timed region; SURVEY.md section 8(f) row 3.  The tower itself runs on the
native engine (``clip_engine.ClipTextEngine``); this module only supplies its WEIGHTS -- the state dict of a
``transformers.CLIPTextModel`` (CLIP-L/14 text tower, 77 tokens, final LayerNorm), loaded from a local
diffusers ``text_encoder`` directory or random-initialised -- and the tokenizer.  No CLIP BPE vocabulary
exists offline, so unless a local tokenizer directory is given prompts are tokenised by a deterministic
hash tokenizer with CLIP's framing (BOS 49406, EOS/pad 49407, length 77, ids in [0, 49405]).
"""
from __future__ import annotations

import os
import re
import zlib

import torch

BOS, EOS, VOCAB, MAX_LEN = 49406, 49407, 49408, 77


class HashTokenizer:
    """Whitespace/punctuation split + CRC32 -> id.  Deterministic across processes and ranks."""

    model_max_length = MAX_LEN

    def __call__(self, prompts, max_length=MAX_LEN, **_):
        if isinstance(prompts, str):
            prompts = [prompts]
        ids = torch.full((len(prompts), max_length), EOS, dtype=torch.long)
        mask = torch.zeros((len(prompts), max_length), dtype=torch.long)
        for i, p in enumerate(prompts):
            words = re.findall(r"[a-z0-9]+|[^\sa-z0-9]", p.lower())[: max_length - 2]
            toks = [BOS] + [zlib.crc32(w.encode()) % (VOCAB - 3) for w in words] + [EOS]
            ids[i, : len(toks)] = torch.tensor(toks)
            mask[i, : len(toks)] = 1
        return ids, mask


def load_tokenizer(path=None):
    if path and os.path.isdir(path):
        from transformers import CLIPTokenizer

        tok = CLIPTokenizer.from_pretrained(path)

        def call(prompts, max_length=MAX_LEN, **_):
            out = tok(prompts, padding="max_length", max_length=max_length, truncation=True, return_tensors="pt")
            return out.input_ids, out.attention_mask

        return call
    return HashTokenizer()


class TextWeights:
    """What the pipelines keep in ``pipe.text_encoder``: the CLIP text tower's state dict and config
    (transformers key names); the module that produced them is discarded -- there is no library forward path."""

    def __init__(self, state_dict, config):
        self._sd = state_dict
        self.config = config

    def state_dict(self):
        return self._sd

    def to(self, *_a, **_k):
        return self


def make_text_encoder(seed: int = 29, path=None, dtype=torch.bfloat16, device="cpu", hidden=768, layers=12,
                      heads=12) -> TextWeights:
    """CLIP-L text tower weights from ``path`` (diffusers ``text_encoder`` dir) or seeded random-init."""
    from transformers import CLIPTextConfig, CLIPTextModel

    if path and os.path.isdir(path):
        model = CLIPTextModel.from_pretrained(path)
    else:
        cfg = CLIPTextConfig(vocab_size=VOCAB, hidden_size=hidden, intermediate_size=4 * hidden,
                             num_hidden_layers=layers, num_attention_heads=heads,
                             max_position_embeddings=MAX_LEN, hidden_act="quick_gelu", projection_dim=hidden,
                             bos_token_id=BOS, eos_token_id=EOS, pad_token_id=1)
        state = torch.random.get_rng_state()
        torch.manual_seed(seed + 2)
        try:
            model = CLIPTextModel(cfg)
        finally:
            torch.random.set_rng_state(state)
    sd = {k: v.detach().to(dtype).clone() for k, v in model.state_dict().items()}
    return TextWeights(sd, model.config)
