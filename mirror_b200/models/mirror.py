"""B200-native MIRROR model — drop-in for the reference ``models/mirror.py``.

Same public API (class names, constructor arguments and defaults, attribute
names, ``state_dict`` keys and shapes, output tuple order) with
all device arithmetic routed through the hand-written sm_100a kernels
(``mirror_b200.ops`` / ``mirror_b200.kernels``).  ``nn.Module`` objects are used
as *parameter containers* so that ``state_dict()`` is key-compatible with
reference checkpoints (``tools/split_weights.py``, ``load_checkpoint``); their
``forward`` methods are never called on the hot path.

Reference lines are cited per class (paths relative to the reference checkout).
Not supported (the reference never enables them for MIRROR): qk_norm,
LayerScale init_values, DropPath > 0, pre_norm, attention dropout > 0.
"""
import logging
import math
from typing import Optional, Tuple

import numpy as np
import torch
from torch import nn

from .. import kernels as K
from .. import ops

_logger = logging.getLogger(__name__)

try:  # the trainers resolve the model through timm's registry (train_mirror.py:689)
    from timm.models import register_model  # type: ignore
except Exception:  # timm absent (e.g. this image): keep a local registry with the same decorator shape
    _REGISTRY = {}

    def register_model(fn):
        _REGISTRY[fn.__name__] = fn
        return fn


def _norm_eps(norm_layer) -> float:
    """timm semantics (models/mirror.py:210): get_norm_layer(None) -> LayerNorm(eps=1e-6); 'layernorm' -> timm LayerNorm (1e-6)."""
    if norm_layer is None or (isinstance(norm_layer, str) and norm_layer.replace("_", "").lower() == "layernorm"):
        return 1e-6
    raise NotImplementedError(f"norm layer {norm_layer!r}: only LayerNorm is implemented in the B200 path")


def _check_gelu(act_layer):
    if act_layer is None or (isinstance(act_layer, str) and act_layer.lower() == "gelu"):
        return
    raise NotImplementedError(f"activation {act_layer!r}: only (exact) GELU is implemented in the B200 path")


class _Mlp(nn.Module):
    """Parameter container with timm ``Mlp`` key names: fc1, norm (optional), fc2."""

    def __init__(self, in_features, hidden_features, out_features, norm_eps: Optional[float] = None, drop: float = 0.0):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.norm = nn.LayerNorm(hidden_features, eps=norm_eps) if norm_eps is not None else nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = drop

    def forward(self, x, x16=None, row_res=None, residual=None):
        """[residual +] drop(fc2(norm(drop(GELU(fc1(x)))))) [+ row_res]  (timm Mlp order: fc1, act, drop1, norm, fc2, drop2)."""
        p = self.drop if self.training else 0.0
        y = ops.linear(x, self.fc1.weight, self.fc1.bias, x16=x16)
        y = ops.activation(y, K.ACT_GELU, p)
        y16 = None
        if isinstance(self.norm, nn.LayerNorm):
            y, y16 = ops.layer_norm(y, self.norm.weight, self.norm.bias, self.norm.eps)
        return ops.linear(y, self.fc2.weight, self.fc2.bias, row_res=row_res, x16=y16, res=residual, drop_p=p)


# ===========================================
#  Transformer for Transcriptomics Data
# ===========================================
class Attention(nn.Module):
    """models/mirror.py:50-102 (parameters qkv, proj)."""

    def __init__(self, dim, num_heads=12, qkv_bias=True, proj_drop=0.0):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        assert num_heads == 12, "the RNA attention kernel is written for the reference's 12 chunks"
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = proj_drop

    def forward(self, x, x16=None, residual=None):
        """[residual +] proj_drop(proj(SDPA(qkv(x))))"""
        qkv = ops.linear(x, self.qkv.weight, self.qkv.bias, x16=x16)
        o = ops.RnaAttnFn.apply(qkv)
        return ops.linear(o, self.proj.weight, self.proj.bias, res=residual, drop_p=self.proj_drop if self.training else 0.0)


class Block(nn.Module):
    """models/mirror.py:105-152: x += attn(LN(x)); x += mlp(LN(x))."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=True, proj_drop=0.0, norm_eps=1e-6):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=norm_eps)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, proj_drop=proj_drop)
        self.norm2 = nn.LayerNorm(dim, eps=norm_eps)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio), dim, drop=proj_drop)

    def forward(self, x):
        y, y16 = ops.layer_norm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        x = self.attn(y, y16, residual=x)
        y, y16 = ops.layer_norm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        return self.mlp(y, y16, residual=x)


class TransFormer(nn.Module):
    """models/mirror.py:155-289."""

    def __init__(self, input_dim, embed_dim=768, depth=2, num_heads=12, mlp_ratio=4.0, qkv_bias=True, qk_norm=False,
                 init_values=None, gene_embed="learn", pre_norm=False, final_norm=True, embed_drop_rate=0.0,
                 pos_drop_rate=0.0, proj_drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, weight_init="",
                 fix_init=False, norm_layer=None, act_layer=None):
        super().__init__()
        assert gene_embed in ("", "none", "learn")
        if qk_norm or init_values or pre_norm or attn_drop_rate > 0 or drop_path_rate > 0 or pos_drop_rate > 0 or not final_norm:
            raise NotImplementedError("qk_norm / LayerScale / pre_norm / attn-drop / drop-path / pos-drop are not part of the MIRROR hot path")
        eps = _norm_eps(norm_layer)
        _check_gelu(act_layer)
        self.num_features = self.head_hidden_size = self.embed_dim = embed_dim
        self.embedding = _Mlp(input_dim, embed_dim * 2, embed_dim, norm_eps=eps, drop=embed_drop_rate)
        self.gene_embed = None if not gene_embed or gene_embed == "none" else nn.Parameter(torch.randn(1, embed_dim) * 0.02)
        self.blocks = nn.Sequential(*[
            Block(embed_dim, num_heads, mlp_ratio, qkv_bias, proj_drop_rate, eps) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=eps)
        if self.gene_embed is not None:
            nn.init.trunc_normal_(self.gene_embed, std=0.02)
        if fix_init:
            for layer_id, layer in enumerate(self.blocks):
                layer.attn.proj.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))
                layer.mlp.fc2.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))

    def forward(self, x):
        x = self.embedding(x.float(), row_res=self.gene_embed)
        for blk in self.blocks:
            x = blk(x)
        return ops.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)[0]


# ===========================================
#  TransMIL for Histopathology Data
# ===========================================
class _NystromParams(nn.Module):
    """Parameter container with the key names of ``nystrom_attention.NystromAttention`` (heads 8, landmarks dim//2,
    6 pinv iterations, 33-tap residual conv, dropout 0.1; models/mirror.py:299-309)."""

    def __init__(self, dim, heads=8, dropout=0.1, kernel=33):
        super().__init__()
        assert dim % heads == 0 and dim % 2 == 0
        if heads != ops.WSI_HEADS or kernel != 33:
            raise NotImplementedError("the Nystrom kernels are written for the reference's 8 heads / 33-tap residual (models/mirror.py:299-309)")
        self.heads, self.dropout = heads, dropout
        self.to_qkv = nn.Linear(dim, dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(dim, dim), nn.Dropout(dropout))
        self.res_conv = nn.Conv2d(heads, heads, (kernel, 1), padding=(kernel // 2, 0), groups=heads, bias=False)


class TransLayer(nn.Module):
    """models/mirror.py:295-314."""

    def __init__(self, dim=512):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.attn = _NystromParams(dim)

    def forward(self, x, per_slide_scale=False):
        """per_slide_scale: the Moore-Penrose initial scale per slide instead of over the batch (variable-length bags)."""
        a = self.attn
        return ops.nystrom_layer(x, self.norm.weight, self.norm.bias, a.to_qkv.weight, a.to_out[0].weight, a.to_out[0].bias,
                                 a.res_conv.weight, a.to_out[1].p if self.training else 0.0, self.norm.eps, per_slide_scale)


class PPEG(nn.Module):
    """models/mirror.py:317-331."""

    def __init__(self, dim=512):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim, 7, 1, 7 // 2, groups=dim)
        self.proj1 = nn.Conv2d(dim, dim, 5, 1, 5 // 2, groups=dim)
        self.proj2 = nn.Conv2d(dim, dim, 3, 1, 3 // 2, groups=dim)

    def forward(self, x, H=None, W=None):
        return ops.PpegFn.apply(x, self.proj.weight, self.proj.bias, self.proj1.weight, self.proj1.bias, self.proj2.weight,
                                self.proj2.bias)


class FeatureTransMIL(nn.Module):
    """models/mirror.py:334-380."""

    def __init__(self, input_dim=1024, embed_dim=512):
        super().__init__()
        self.input_dim, self.embed_dim = input_dim, embed_dim
        self.pos_layer = PPEG(dim=embed_dim)
        self._fc1 = nn.Sequential(nn.Linear(input_dim, embed_dim), nn.ReLU())
        self.cls_token = nn.Parameter(torch.randn(1, 1, embed_dim))
        self.layer1 = TransLayer(dim=embed_dim)
        self.layer2 = TransLayer(dim=embed_dim)
        self.norm = nn.LayerNorm(embed_dim)

    def _tokens(self, h, drop_wrap=False, per_slide_scale=False):
        """fc1+ReLU, wrap pad, cls, layer1, PPEG, layer2, final norm -> ([B,S,E] f32, bf16 copy, add_length).
        ``drop_wrap``: the final norm returns only the N+1 real tokens (contiguous [B,N+1,E]); add_length is then 0."""
        N = h.shape[1]
        Hs = int(np.ceil(np.sqrt(N)))
        add = Hs * Hs - N
        x = ops.WsiEmbedFn.apply(h.float(), self._fc1[0].weight, self._fc1[0].bias, self.cls_token)
        x = self.layer1(x, per_slide_scale)
        x = self.pos_layer(x)
        x = self.layer2(x, per_slide_scale)
        if drop_wrap:
            y, y16 = ops.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps, keep=N + 1)
            return y, y16, 0
        y, y16 = ops.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        return y, y16, add

    def forward(self, h):
        return self._tokens(h)[0][:, 0]


# ===========================================
#  TransFormer for Pre-training
# ===========================================
class TransFormerHybrid(TransFormer):
    """models/mirror.py:386-569."""

    def __init__(self, input_dim, embed_dim=768, depth=2, num_heads=12, mlp_ratio=4.0, qkv_bias=True, gene_embed="learn",
                 pos_drop_rate=0.0, proj_drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, norm_layer=None,
                 act_layer=None, retention_decoder_depth=1, **kw):
        super().__init__(input_dim=input_dim, embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio,
                         qkv_bias=qkv_bias, gene_embed=gene_embed, pos_drop_rate=pos_drop_rate, proj_drop_rate=proj_drop_rate,
                         attn_drop_rate=attn_drop_rate, drop_path_rate=drop_path_rate, norm_layer=norm_layer,
                         act_layer=act_layer, **kw)
        eps = _norm_eps(norm_layer)
        self.alignment_head = nn.Linear(embed_dim, embed_dim)
        self.retention_embed = nn.Linear(embed_dim, embed_dim)
        self.mask_token = nn.Parameter(torch.zeros(1, 1))
        self.retention_gene_embed = nn.Parameter(torch.randn(1, embed_dim) * 0.02)
        self.retention_blocks = nn.ModuleList([
            Block(embed_dim, num_heads, mlp_ratio, qkv_bias, proj_drop_rate, eps) for _ in range(retention_decoder_depth)])
        self.retention_norm = nn.LayerNorm(embed_dim, eps=eps)
        self.retention_head = nn.Linear(embed_dim, embed_dim)
        nn.init.normal_(self.mask_token, std=0.02)
        nn.init.trunc_normal_(self.retention_gene_embed, std=0.02)
        for layer_id, layer in enumerate(self.retention_blocks):
            layer.attn.proj.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))
            layer.mlp.fc2.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))

    def random_masking(self, x, mask_ratio, noise=None):
        """models/mirror.py:510-533; returns (masked x + retention_gene_embed, mask)."""
        B, N = x.shape
        keep = int(N * (1 - mask_ratio))
        if noise is None:
            noise = torch.rand(B, N, device=x.device)
        mask = K.rank_mask(noise.float().contiguous(), keep)
        r = ops.MaskPosFn.apply(x.view(B, N, 1), mask, self.mask_token, self.retention_gene_embed.view(N, 1), 0)
        return r.view(B, N), mask

    def forward_encoder(self, x):
        return super().forward(x)

    def forward_alignment_head(self, x):
        eps = 1e-6 if x.dtype == torch.float16 else 1e-12
        return ops.linear(ops.l2_normalize(x, eps), self.alignment_head.weight, self.alignment_head.bias)

    def forward_retention_head(self, x, mask_ratio, noise=None):
        r = ops.linear(x, self.retention_embed.weight, self.retention_embed.bias)
        r, mask = self.random_masking(r, mask_ratio, noise)  # includes "+ retention_gene_embed" (:549)
        for blk in self.retention_blocks:
            r = blk(r)
        r, r16 = ops.layer_norm(r, self.retention_norm.weight, self.retention_norm.bias, self.retention_norm.eps)
        return ops.linear(r, self.retention_head.weight, self.retention_head.bias, x16=r16), mask

    def forward_decoders(self, x, mask_ratio, noise=None):
        a = self.forward_alignment_head(x)
        r, mask = self.forward_retention_head(x, mask_ratio, noise)
        return a, r, mask

    def forward(self, x, mask_ratio=0.75):
        x = self.forward_encoder(x)
        a, r, mask = self.forward_decoders(x, mask_ratio)
        return a, r, x, mask


# ===========================================
#  TransMIL for Pre-training
# ===========================================
class FeatureTransMILHybrid(FeatureTransMIL):
    """models/mirror.py:575-714."""

    def __init__(self, input_dim=1024, embed_dim=512, num_tokens=2048, retention_decoder_depth=1):
        super().__init__(input_dim, embed_dim)
        self.num_tokens = num_tokens
        self.alignment_head = nn.Linear(embed_dim, embed_dim)
        self.retention_embed = nn.Linear(embed_dim, embed_dim)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.retention_gene_embed = nn.Parameter(torch.randn(1, num_tokens + 1, embed_dim) * 0.02)
        self.retention_blocks = nn.ModuleList([TransLayer(dim=embed_dim) for _ in range(retention_decoder_depth)])
        self.retention_norm = nn.LayerNorm(embed_dim)
        self.retention_head = nn.Linear(embed_dim, embed_dim)
        self.init_weights()

    def init_weights(self):
        nn.init.normal_(self.mask_token, std=0.02)
        nn.init.normal_(self.cls_token, std=0.02)
        nn.init.trunc_normal_(self.retention_gene_embed, std=0.02)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward_encoder(self, h, per_slide_scale=False):
        y, y16, _ = self._tokens(h, drop_wrap=True, per_slide_scale=per_slide_scale)  # LayerNorm + `h[:, :-add_length]` (models/mirror.py:372) in one pass
        # bf16 copy for the retention decoder, attached to the TENSOR it mirrors (with its version counter), not to the
        # module: a different or modified embedding of the same shape cannot pick it up, and nothing stays pinned on the
        # module (deepcopy / ModelEma) once the caller drops the embedding
        y._mirror_side = (y16, y._version)
        return y

    def forward_alignment_head(self, h, cls=None):
        eps = 1e-6 if h.dtype == torch.float16 else 1e-12
        cls = h[:, 0, :] if cls is None else cls
        return ops.linear(ops.l2_normalize(cls, eps), self.alignment_head.weight, self.alignment_head.bias)

    def forward_retention_head(self, h, mask_ratio, noise=None, per_slide_scale=False):
        B, T, E = h.shape  # T = N + 1
        N = T - 1
        if T > self.retention_gene_embed.shape[1]:
            raise ValueError(f"{N} patches exceed the position table (wsi_num_tokens = {self.retention_gene_embed.shape[1] - 1})")
        keep = int(N * (1 - mask_ratio))
        if noise is None:
            noise = torch.rand(B, N, device=h.device)
        mask = K.rank_mask(noise.float().contiguous(), keep)
        side = getattr(h, "_mirror_side", None)
        x16 = None
        if side is not None and side[1] == h._version and side[0].shape == h.shape and h.is_contiguous() and h.dtype == torch.float32:
            x16 = side[0]
        r = ops.linear(h, self.retention_embed.weight, self.retention_embed.bias, x16=x16)
        pos = self.retention_gene_embed if T == self.retention_gene_embed.shape[1] else self.retention_gene_embed[:, :T]
        r = ops.MaskPosFn.apply(r, mask, self.mask_token, pos, 1)
        for blk in self.retention_blocks:
            r = blk(r, per_slide_scale)
        r, r16 = ops.layer_norm(r, self.retention_norm.weight, self.retention_norm.bias, self.retention_norm.eps)
        r = ops.linear(r, self.retention_head.weight, self.retention_head.bias, x16=r16)
        return ops.drop_first_token(r), mask

    def forward_decoders(self, h, mask_ratio, noise=None, cls=None, per_slide_scale=False):
        a = self.forward_alignment_head(h, cls)
        r, mask = self.forward_retention_head(h, mask_ratio, noise, per_slide_scale)
        return a, r, mask

    def forward(self, h, mask_ratio=0.75):
        h = self.forward_encoder(h)
        a, r, mask = self.forward_decoders(h, mask_ratio)
        return a, r, h[:, 1:, :], mask


# ===========================================
#  MIRROR for Pre-training
# ===========================================
class MIRROR(nn.Module):
    """models/mirror.py:720-915.  ``noise`` (optional, not in the reference signature) injects the four random draws of
    the forward (``wsi_mask``, ``rna_mask``, ``wsi_eps``, ``rna_eps``) for parity tests; by default they are drawn with
    ``torch.rand`` / ``torch.randn`` in the reference's call order."""

    def __init__(self, wsi_embed_dim: int, rna_embed_dim: int, embed_dim: int, wsi_num_tokens: int = 2048,
                 wsi_retention_decoder_depth: int = 1, rna_encoder_depth: int = 2, rna_gene_embed: str = "learn",
                 rna_mlp_ratio: float = 2.572, rna_pos_drop_rate: float = 0.0, rna_proj_drop_rate: float = 0.1,
                 rna_attn_drop_rate: float = 0.0, rna_drop_path_rate: float = 0.0, rna_norm_layer=None, rna_act_layer=None,
                 rna_retention_decoder_depth: int = 1, init_logit_scale: float = np.log(1 / 0.07),
                 style_mlp_hidden_dim: int = 512, style_mlp_out_dim: int = 256, style_norm_layer=None, style_act_layer=None,
                 style_latent_dim: int = 128, num_prototypes: int = 3000) -> None:
        super().__init__()
        self.wsi_embed_dim, self.rna_embed_dim, self.embed_dim = wsi_embed_dim, rna_embed_dim, embed_dim
        self.wsi_num_tokens = wsi_num_tokens
        self.wsi_retention_decoder_depth = wsi_retention_decoder_depth
        self.rna_encoder_depth, self.rna_gene_embed, self.rna_mlp_ratio = rna_encoder_depth, rna_gene_embed, rna_mlp_ratio
        self.rna_pos_drop_rate, self.rna_proj_drop_rate = rna_pos_drop_rate, rna_proj_drop_rate
        self.attn_drop_rate, self.drop_path_rate = rna_attn_drop_rate, rna_drop_path_rate
        self.rna_norm_layer, self.rna_act_layer = rna_norm_layer, rna_act_layer
        self.rna_retention_decoder_depth = rna_retention_decoder_depth
        if embed_dim % 24:
            raise ValueError("embed_dim must be a multiple of 24 (8 Nystrom heads and 12 RNA chunks, models/mirror.py:64,301)")
        if style_norm_layer is not None:
            raise NotImplementedError("style_norm_layer is not used by the reference configuration")
        _check_gelu(style_act_layer)

        self.logit_scale = nn.Parameter(torch.ones([]) * init_logit_scale)
        self.wsi_encoder = FeatureTransMILHybrid(input_dim=wsi_embed_dim, embed_dim=embed_dim, num_tokens=wsi_num_tokens,
                                                 retention_decoder_depth=wsi_retention_decoder_depth)
        self.rna_encoder = TransFormerHybrid(input_dim=rna_embed_dim, embed_dim=embed_dim, depth=rna_encoder_depth,
                                             gene_embed=rna_gene_embed, mlp_ratio=rna_mlp_ratio, pos_drop_rate=rna_pos_drop_rate,
                                             proj_drop_rate=rna_proj_drop_rate, attn_drop_rate=rna_attn_drop_rate,
                                             drop_path_rate=rna_drop_path_rate, norm_layer=rna_norm_layer,
                                             act_layer=rna_act_layer, retention_decoder_depth=rna_retention_decoder_depth)
        self.style_encoder_mlp = _Mlp(embed_dim, style_mlp_hidden_dim, style_mlp_out_dim, drop=0.0)
        self.style_mu = nn.Linear(style_mlp_out_dim, style_latent_dim)
        self.style_logstd = nn.Linear(style_mlp_out_dim, style_latent_dim)
        self.style_decoder = nn.Linear(style_latent_dim, embed_dim)
        self.prototypes = nn.Linear(embed_dim, num_prototypes, bias=False)
        nn.init.orthogonal_(self.prototypes.weight)

    @torch.no_grad()
    def normalize_prototypes(self):
        """The trainer's per-step `prototypes.weight.data = F.normalize(w, dim=1, p=2)` (train_mirror.py:1133-1136) as one
        kernel launch, in place."""
        w = self.prototypes.weight
        y, _ = K.l2norm_fwd(w.data, 1e-12)
        K.copy_rows_(y, w.data)

    def reparameterize(self, mu, logstd, eps=None):
        if eps is None:
            eps = torch.randn_like(mu)
        return ops.ReparamFn.apply(mu, logstd, eps.float())

    def forward_style_clustering(self, wsi_emb, rna_emb, wsi_eps=None, rna_eps=None):
        """models/mirror.py:835-858; both modalities share the weights, so they run as ONE stacked batch of 2B rows."""
        B = wsi_emb.shape[0]
        x = ops.stack_rows(wsi_emb, rna_emb)  # [2B,E]
        x = self.style_encoder_mlp(x)
        mu = ops.linear(x, self.style_mu.weight, self.style_mu.bias)
        logstd = ops.linear(x, self.style_logstd.weight, self.style_logstd.bias)
        if wsi_eps is None:
            wsi_eps = torch.randn(B, mu.shape[1], device=mu.device)
        if rna_eps is None:
            rna_eps = torch.randn(B, mu.shape[1], device=mu.device)
        z = self.reparameterize(mu, logstd, torch.cat([wsi_eps.float(), rna_eps.float()], 0))
        z = ops.linear(z, self.style_decoder.weight, self.style_decoder.bias)
        score = ops.linear(z, self.prototypes.weight)
        return score[:B], mu[:B], logstd[:B], score[B:], mu[B:], logstd[B:]

    def forward(self, wsi_emb, rna_emb, wsi_mask_ratio: float = 0.75, rna_mask_ratio: float = 0.75, noise=None) -> Tuple[torch.Tensor, ...]:
        noise = noise or {}
        # the encoder output is read whole (retention decoder), as its cls row (alignment head, style encoder) and as its
        # patch rows (retention target): one fan-out node, so that the backward merges the three gradients in one pass
        enc = self.wsi_encoder.forward_encoder(wsi_emb)
        wsi_cls, wsi_emb, wsi_retention_target = ops.token_fanout(enc)
        side = getattr(enc, "_mirror_side", None)
        if side is not None and wsi_emb.data_ptr() == enc.data_ptr():  # the fan-out's full view aliases the encoder output
            wsi_emb._mirror_side = (side[0], wsi_emb._version)
        wa, wr, wm = self.wsi_encoder.forward_decoders(wsi_emb, wsi_mask_ratio, noise.get("wsi_mask"), cls=wsi_cls)
        rna_emb = self.rna_encoder.forward_encoder(rna_emb)
        ra, rr, rm = self.rna_encoder.forward_decoders(rna_emb, rna_mask_ratio, noise.get("rna_mask"))
        ws, wmu, wls, rs, rmu, rls = self.forward_style_clustering(wsi_cls, rna_emb, noise.get("wsi_eps"), noise.get("rna_eps"))
        return (wa, wr, wsi_retention_target, wm, ws, wmu, wls, ra, rr, rna_emb, rm, rs, rmu, rls, self.logit_scale.exp())


def _length_groups(bags):
    """{N: [slide indices]} in first-appearance order."""
    groups = {}
    for i, b in enumerate(bags):
        if b.dim() != 2:
            raise ValueError("every bag must be a [N_i, Dw] tensor")
        groups.setdefault(int(b.shape[0]), []).append(i)
    return groups


def _forward_varlen(self, bags, rna_emb, wsi_mask_ratio=0.75, rna_mask_ratio=0.75, noise=None):
    """Variable-length bags: ``bags`` is a list of [N_i, Dw] feature tensors, one per slide, instead of the
    reference's resample-every-slide-to-N batch (datasets/dataset_pretrain.py:157-161).  Every slide keeps its own geometry
    -- H_i = ceil(sqrt(N_i)), S_i = H_i^2 + 1, landmark group size ceil(S_i / m), Moore-Penrose scale, int(N_i (1 - r)) kept
    tokens, position table sliced to N_i + 1 rows (N_i <= wsi_num_tokens) -- i.e. the result is the reference at B = 1 per
    slide with wsi_num_tokens = N_i.  Slides of EQUAL length run as one dense batch through the fixed-length kernels
    (bucket the dataset by length to keep the launch count down); the ragged retention outputs are packed along the token
    axis, [1, sum N_i, E] with mask [1, sum N_i], which MIRRORLoss's masked MSE consumes unchanged.  Returns the 15-tuple of
    ``forward`` with the batch-level entries in the order of ``bags``.
    ``noise["wsi_mask"]``: optional list of [1, N_i] tensors (parity tests)."""
    noise = noise or {}
    B = len(bags)
    enc = self.wsi_encoder
    cls_rows, wa_rows, wr_parts, wt_parts, wm_parts = [None] * B, [None] * B, [None] * B, [None] * B, [None] * B
    for N, idx in _length_groups(bags).items():
        x = torch.stack([bags[i] for i in idx]) if len(idx) > 1 else bags[idx[0]][None]
        h = enc.forward_encoder(x, per_slide_scale=True)
        cls, full, tgt = ops.token_fanout(h)
        side = getattr(h, "_mirror_side", None)
        if side is not None and full.data_ptr() == h.data_ptr():
            full._mirror_side = (side[0], full._version)
        nz = None
        if noise.get("wsi_mask") is not None:
            nz = torch.cat([noise["wsi_mask"][i].reshape(1, N) for i in idx])
        wa, wr, wm = enc.forward_decoders(full, wsi_mask_ratio, nz, cls=cls, per_slide_scale=True)
        for j, i in enumerate(idx):
            cls_rows[i], wa_rows[i] = cls[j:j + 1], wa[j:j + 1]
            wr_parts[i], wt_parts[i], wm_parts[i] = wr[j], tgt[j], wm[j]
    wsi_cls, wa = torch.cat(cls_rows), torch.cat(wa_rows)
    wr, wt, wm = torch.cat(wr_parts)[None], torch.cat(wt_parts)[None], torch.cat(wm_parts)[None]
    rna_emb = self.rna_encoder.forward_encoder(rna_emb)
    ra, rr, rm = self.rna_encoder.forward_decoders(rna_emb, rna_mask_ratio, noise.get("rna_mask"))
    ws, wmu, wls, rs, rmu, rls = self.forward_style_clustering(wsi_cls, rna_emb, noise.get("wsi_eps"), noise.get("rna_eps"))
    return (wa, wr, wt, wm, ws, wmu, wls, ra, rr, rna_emb, rm, rs, rmu, rls, self.logit_scale.exp())


MIRROR.forward_varlen = _forward_varlen


class MIRRORDualEncoder(nn.Module):
    """The 2-output model ``train_pretrain.py:1119-1122`` unpacks (the reference registers none):
    FeatureTransMIL cls embedding (models/mirror.py:352-380) and TransFormer embedding (:283-289)."""

    def __init__(self, wsi_embed_dim, rna_embed_dim, embed_dim, rna_encoder_depth=2, rna_gene_embed="learn", rna_mlp_ratio=2.572,
                 rna_proj_drop_rate=0.1, rna_norm_layer=None, rna_act_layer=None):
        super().__init__()
        self.wsi_encoder = FeatureTransMIL(input_dim=wsi_embed_dim, embed_dim=embed_dim)
        self.rna_encoder = TransFormer(input_dim=rna_embed_dim, embed_dim=embed_dim, depth=rna_encoder_depth,
                                       gene_embed=rna_gene_embed, mlp_ratio=rna_mlp_ratio, proj_drop_rate=rna_proj_drop_rate,
                                       norm_layer=rna_norm_layer, act_layer=rna_act_layer)

    def forward(self, wsi_emb, rna_emb):
        return self.wsi_encoder(wsi_emb), self.rna_encoder(rna_emb)


class MIRRORClassifier(nn.Module):
    """models/mirror.py:921-1015: the downstream (sub-typing / survival) model -- FeatureTransMIL cls embedding, optional
    TransFormer RNA embedding, "add" or "concat" fusion, linear head -- on the same kernels, with the reference's attribute
    names and state_dict keys (``wsi_encoder.*``, ``rna_encoder.*``, ``head.*``) so that a split pre-training checkpoint
    (tools/split_weights.py) loads with ``load_checkpoint(strict=False)``."""

    def __init__(self, wsi_embed_dim: int, rna_embed_dim: int, embed_dim: int, num_classes: int, rna_encoder_depth: int = 2,
                 rna_gene_embed: str = "learn", rna_mlp_ratio: float = 2.572, rna_pos_drop_rate: float = 0.0,
                 rna_proj_drop_rate: float = 0.1, rna_attn_drop_rate: float = 0.0, rna_drop_path_rate: float = 0.0,
                 rna_norm_layer=None, rna_act_layer=None, fusion: str = "concat") -> None:
        super().__init__()
        self.wsi_embed_dim, self.rna_embed_dim, self.embed_dim = wsi_embed_dim, rna_embed_dim, embed_dim
        self.rna_encoder_depth, self.rna_gene_embed, self.rna_mlp_ratio = rna_encoder_depth, rna_gene_embed, rna_mlp_ratio
        self.rna_pos_drop_rate, self.rna_proj_drop_rate = rna_pos_drop_rate, rna_proj_drop_rate
        self.attn_drop_rate, self.drop_path_rate = rna_attn_drop_rate, rna_drop_path_rate
        self.rna_norm_layer, self.rna_act_layer = rna_norm_layer, rna_act_layer
        self.num_classes, self.fusion = num_classes, fusion
        assert self.fusion in ["add", "concat"], "Fusion must be either add or concat"
        self.wsi_encoder = FeatureTransMIL(input_dim=wsi_embed_dim, embed_dim=embed_dim)
        self.rna_encoder = TransFormer(input_dim=rna_embed_dim, embed_dim=embed_dim, depth=rna_encoder_depth,
                                       gene_embed=rna_gene_embed, mlp_ratio=rna_mlp_ratio, pos_drop_rate=rna_pos_drop_rate,
                                       proj_drop_rate=rna_proj_drop_rate, attn_drop_rate=rna_attn_drop_rate,
                                       drop_path_rate=rna_drop_path_rate, norm_layer=rna_norm_layer, act_layer=rna_act_layer)
        self.head = nn.Linear(embed_dim * (2 if fusion == "concat" else 1), num_classes)

    def forward(self, wsi_emb: torch.Tensor, rna_emb: Optional[torch.Tensor] = None) -> torch.Tensor:
        w = self.wsi_encoder(wsi_emb)
        if rna_emb is None:
            return ops.linear(w, self.head.weight, self.head.bias)
        r = self.rna_encoder(rna_emb)
        if self.fusion == "add":
            # W (w + r) + b as two products sharing the bias: no separate add kernel, both gradients from the same GEMMs
            return ops.linear(w, self.head.weight, self.head.bias, res=ops.linear(r, self.head.weight))
        E = self.embed_dim  # concat: [w | r] W^T = w W[:, :E]^T + r W[:, E:]^T
        return ops.linear(w, self.head.weight[:, :E], self.head.bias, res=ops.linear(r, self.head.weight[:, E:]))


_MIRROR_ARGS = {
    "wsi_embed_dim", "rna_embed_dim", "embed_dim", "wsi_num_tokens", "wsi_retention_decoder_depth", "rna_encoder_depth",
    "rna_gene_embed", "rna_mlp_ratio", "rna_pos_drop_rate", "rna_proj_drop_rate", "rna_attn_drop_rate", "rna_drop_path_rate",
    "rna_norm_layer", "rna_act_layer", "rna_retention_decoder_depth", "init_logit_scale", "style_mlp_hidden_dim",
    "style_mlp_out_dim", "style_norm_layer", "style_act_layer", "style_latent_dim", "num_prototypes",
}
_DUAL_ARGS = {"wsi_embed_dim", "rna_embed_dim", "embed_dim", "rna_encoder_depth", "rna_gene_embed", "rna_mlp_ratio",
              "rna_proj_drop_rate", "rna_norm_layer", "rna_act_layer"}


def _filtered(kwargs, accepted):
    dropped = [k for k in kwargs if k not in accepted]
    if dropped:  # timm injects pretrained / pretrained_cfg / ...: filter with a warning like models/mirror.py:1045-1053
        _logger.warning("Filtered model kwargs: %s", ", ".join(dropped))
    return {k: v for k, v in kwargs.items() if k in accepted}


@register_model
def mirror(**kwargs):
    return MIRROR(**_filtered(kwargs, _MIRROR_ARGS))


_CLS_ARGS = {"wsi_embed_dim", "rna_embed_dim", "embed_dim", "rna_encoder_depth", "rna_gene_embed", "rna_mlp_ratio", "rna_pos_drop_rate",
             "rna_proj_drop_rate", "rna_attn_drop_rate", "rna_drop_path_rate", "rna_norm_layer", "rna_act_layer", "num_classes", "fusion"}


@register_model
def mirror_classifier(**kwargs):
    return MIRRORClassifier(**_filtered(kwargs, _CLS_ARGS))


@register_model
def mirror_dual_encoder(**kwargs):
    return MIRRORDualEncoder(**_filtered(kwargs, _DUAL_ARGS))
