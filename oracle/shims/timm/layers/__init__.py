"""timm.layers subset: Mlp, DropPath, LayerType, get_act_layer, get_norm_layer,
trunc_normal_, use_fused_attn — semantics of timm 1.0.15."""
from functools import partial
from typing import Callable, Optional, Type, Union

import torch
from torch import nn
from torch.nn.init import trunc_normal_  # noqa: F401  (timm re-exports torch's)

LayerType = Union[str, Callable, Type[nn.Module]]


class LayerNorm(nn.LayerNorm):
    """timm's LayerNorm: eps defaults to 1e-6."""

    def __init__(self, num_channels, eps=1e-6, affine=True):
        super().__init__(num_channels, eps=eps, elementwise_affine=affine)


class GELU(nn.GELU):
    def __init__(self, inplace: bool = False):
        super().__init__()


_NORMS = {"layernorm": LayerNorm}
_ACTS = {"gelu": GELU, "relu": nn.ReLU}


def get_norm_layer(norm_layer: Optional[LayerType]):
    if norm_layer is None:
        return None
    if isinstance(norm_layer, str):
        return _NORMS[norm_layer.replace("_", "").lower()] if norm_layer else None
    return norm_layer


def get_act_layer(name: Optional[LayerType] = "relu"):
    if name is None:
        return None
    if isinstance(name, str):
        return _ACTS[name] if name else None
    return name


def use_fused_attn(experimental: bool = False) -> bool:
    return True


class DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        r = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            r.div_(keep)
        return x * r


class Mlp(nn.Module):
    """fc1 -> act -> drop1 -> norm(hidden) -> fc2 -> drop2."""

    def __init__(self, in_features, hidden_features=None, out_features=None,
                 act_layer=nn.GELU, norm_layer=None, bias=True, drop=0.0,
                 use_conv=False):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.norm = norm_layer(hidden_features) if norm_layer is not None else nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))
