"""B200 CLIP towers: the transformer encoders behind the CLIP-score metric and the prompt encoder.

Reference call sites: ``ClipScoreMetric`` (/root/reference/src/metrics/metrics.py:25-41, torchmetrics ``CLIPScore`` on
CLIP ViT-B/16) and ``encode_prompt`` (/root/reference/src/models.py:139-149, CLIP-L text tower) -- rows (f)-2 and (f)-3
of SURVEY.md section 8.  One engine class records a pre-LN transformer encoder (LayerNorm -> fused QKV GEMM -> flash
attention, optionally causal -> out-proj GEMM + residual -> LayerNorm -> fc1 GEMM with a QuickGELU epilogue -> fc2
GEMM + residual) into a native launch plan from the same kernels as the UNet; state-dict keys are the
``transformers`` names (``vision_model.*`` / ``text_model.*``), so real CLIP checkpoints load unchanged.

What stays in PyTorch is data movement only: cutting the image into 16x16 patches, the token-embedding gather and
picking the pooled row.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import kernels as K
from ._lib import check, lib
from .unet_engine import Arena, UNetEngine, _Plan


class _Arch:
    norm_num_groups = 32


class ClipEncoderEngine(UNetEngine):
    """``layers`` pre-LN transformer blocks over a resident token buffer ``x`` [n*seq, width] (bf16)."""

    def __init__(self, state_dict, prefix, *, n, seq, width, heads, layers, mlp, causal, pre_ln=None, final_ln=None,
                 device="cuda"):
        # deliberately not calling UNetEngine.__init__: only its recording helpers are reused
        self.sd = state_dict
        self.dev = torch.device(device)
        self.n, self.seq, self.width, self.heads = n, seq, width, heads
        self.arch = _Arch()
        self.arena = Arena(self.dev)
        self.fuse_gn_stats = False
        self._w = {}
        self.x = torch.zeros(n * seq, width, device=self.dev, dtype=torch.bfloat16)
        self.pre_plan = _Plan()                      # recorded by subclasses (patch embedding)
        with torch.no_grad():
            self.plan = _Plan()
            h = self.x
            if pre_ln:
                h = self._ln(self.plan, self.x, pre_ln)
            for i in range(layers):
                self._layer(self.plan, f"{prefix}.encoder.layers.{i}", h, causal)
            self.out = self._ln(self.plan, h, final_ln) if final_ln else h

    def _layer(self, plan, L, x, causal):
        w, d = self.width, self.width // self.heads
        ln = self._ln(plan, x, L + ".layer_norm1")
        key = ("qkv", L)
        if key not in self._w:
            self._w[key] = (
                torch.cat([self._p(f"{L}.self_attn.{n}_proj.weight") for n in "qkv"], 0).to(torch.bfloat16).contiguous(),
                torch.cat([self._p(f"{L}.self_attn.{n}_proj.bias") for n in "qkv"], 0).float().contiguous())
        wqkv, bqkv = self._w[key]
        qkv = self._gemm(plan, ln, wqkv, 3 * w, bias=bqkv)
        self.arena.release(ln)
        ao = self.arena.alloc((self.n * self.seq, w))
        a = K.attention_args(qkv[:, :w], qkv[:, w:2 * w], qkv[:, 2 * w:], ao, batch=self.n, heads=self.heads,
                             seq_q=self.seq, seq_k=self.seq, head_dim=d, causal=causal)
        check(lib().sonic_plan_add_attention(plan.h, C.byref(a)), "sonic_plan_add_attention")
        plan.log.append(f"attention B={self.n} H={self.heads} S={self.seq} d={d} causal={int(causal)}")
        self.arena.release(qkv)
        self._gemm(plan, ao, self._lin(L + ".self_attn.out_proj.weight"), w, bias=self._f32(L + ".self_attn.out_proj.bias"),
                   residual=x, out=x)
        self.arena.release(ao)
        ln = self._ln(plan, x, L + ".layer_norm2")
        h = self._gemm(plan, ln, self._lin(L + ".mlp.fc1.weight"), self.sd[L + ".mlp.fc1.weight"].shape[0],
                       bias=self._f32(L + ".mlp.fc1.bias"), epilogue=K.EPI_QUICK_GELU)
        self.arena.release(ln)
        self._gemm(plan, h, self._lin(L + ".mlp.fc2.weight"), w, bias=self._f32(L + ".mlp.fc2.bias"), residual=x, out=x)
        self.arena.release(h)

    def run(self):
        self.plan.run(K.stream_ptr())
        return self.out


class ClipVisionEngine(ClipEncoderEngine):
    """CLIP ViT image tower: pixel_values (n, 3, S, S) -> projected image embeddings (n, proj)."""

    def __init__(self, state_dict, *, n, image_size=224, patch=16, width=768, heads=12, layers=12, mlp=3072,
                 device="cuda"):
        self.grid = image_size // patch
        self.patch = patch
        seq = self.grid * self.grid + 1
        super().__init__(state_dict, "vision_model", n=n, seq=seq, width=width, heads=heads, layers=layers, mlp=mlp,
                         causal=False, pre_ln="vision_model.pre_layrnorm", device=device)
        np_ = self.grid * self.grid
        self.patches = torch.zeros(n * np_, 3 * patch * patch, device=self.dev, dtype=torch.bfloat16)
        pos = self._p("vision_model.embeddings.position_embedding.weight").float()
        self.cls_pos = (self._p("vision_model.embeddings.class_embedding").float() + pos[0]).to(torch.bfloat16)
        self.pos_patch = pos[1:].to(torch.bfloat16).contiguous()
        wpe = self._p("vision_model.embeddings.patch_embedding.weight")
        self._w["patch"] = wpe.reshape(wpe.shape[0], -1).to(torch.bfloat16).contiguous()
        with torch.no_grad():
            for i in range(n):                       # patch embedding + position, straight into the token rows
                self._gemm(self.pre_plan, self.patches[i * np_:(i + 1) * np_], self._w["patch"], width,
                           residual=self.pos_patch, out=self.x[i * seq + 1:(i + 1) * seq])

    @torch.no_grad()
    def image_features(self, pixel_values):
        """Already-normalised pixel values (n, 3, S, S) -> projected image embeddings."""
        n, g, p_ = self.n, self.grid, self.patch
        assert pixel_values.shape == (n, 3, g * p_, g * p_), pixel_values.shape
        x = pixel_values.to(self.dev).view(n, 3, g, p_, g, p_).permute(0, 2, 4, 1, 3, 5).reshape(n * g * g, -1)
        self.patches.copy_(x)
        return self.features_from_patches()

    @torch.no_grad()
    def image_features_from_images(self, images, rescale_twice=False):
        """uint8 (or float [0,1]) images (n, 3, H, W) -> embeddings: the fused preprocessing kernel
        (``sonic_clip_preprocess``) writes the patch rows of the embedding GEMM directly."""
        K.clip_preprocess(images.contiguous(), size=self.grid * self.patch, patches_out=self.patches, patch=self.patch,
                          rescale_twice=rescale_twice)
        return self.features_from_patches()

    @torch.no_grad()
    def features_from_patches(self):
        n = self.n
        self.x.view(n, self.seq, self.width)[:, 0] = self.cls_pos
        self.pre_plan.run(K.stream_ptr())
        out = self.run()
        pooled = out.view(n, self.seq, self.width)[:, 0].contiguous()
        pooled = K.layernorm(pooled, self._f32("vision_model.post_layernorm.weight"),
                             self._f32("vision_model.post_layernorm.bias"))
        return K.conv_gemm(pooled, self._lin("visual_projection.weight"), self.sd["visual_projection.weight"].shape[0])


class ClipTextEngine(ClipEncoderEngine):
    """CLIP text tower: input_ids (n, 77) -> last hidden state (n, 77, width) / projected text embeddings."""

    def __init__(self, state_dict, *, n, seq=77, width=512, heads=8, layers=12, mlp=2048, device="cuda"):
        super().__init__(state_dict, "text_model", n=n, seq=seq, width=width, heads=heads, layers=layers, mlp=mlp,
                         causal=True, final_ln="text_model.final_layer_norm", device=device)
        self.tok = self._p("text_model.embeddings.token_embedding.weight")
        self.pos = self._p("text_model.embeddings.position_embedding.weight")[:seq]

    @torch.no_grad()
    def last_hidden_state(self, input_ids):
        n = self.n
        assert tuple(input_ids.shape) == (n, self.seq), input_ids.shape
        emb = self.tok[input_ids.to(self.dev)].float() + self.pos.float()
        self.x.copy_(emb.reshape(n * self.seq, self.width))
        return self.run().view(n, self.seq, self.width)

    @torch.no_grad()
    def text_features(self, input_ids):
        h = self.last_hidden_state(input_ids)
        eos = input_ids.to(self.dev).argmax(dim=-1)           # CLIP pools at the (highest-id) EOS token
        pooled = h[torch.arange(self.n, device=self.dev), eos].contiguous()
        return K.conv_gemm(pooled, self._lin("text_projection.weight"), self.sd["text_projection.weight"].shape[0])
