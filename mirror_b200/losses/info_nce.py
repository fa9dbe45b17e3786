"""B200-native InfoNCE — drop-in for the reference ``losses/info_nce.py``.

Same constructor / ``forward(query, positive_key, negative_keys=None)`` signature and the same
``ValueError`` checks (reference :85-120).  The implicit-negatives branch (:144-164), which is the
one ``train_pretrain.py:874`` uses, runs on the fused contrastive kernels (L2-normalise kernel,
tcgen05 logits GEMM, row/column LSE); by SURVEY.md §3.5 ``InfoNCE(T, symmetric=True)(q,k)`` equals
``ClipLoss(normalize(q), normalize(k), 1/T)``.  The explicit-negatives branches of the reference
never assign ``loss`` (UnboundLocalError at :166) and are rejected here with a clear error.
"""
import torch
from torch import nn

from .. import ops

__all__ = ["InfoNCE"]


class InfoNCE(nn.Module):
    def __init__(self, temperature=0.1, reduction="mean", negative_mode="unpaired", symmetric=False, global_negatives=False, group=None):
        super().__init__()
        self.global_negatives, self.group = global_negatives, group  # extension, default off (see losses/mirror_loss.ClipLoss)
        self.temperature = temperature
        self.reduction = reduction
        self.negative_mode = negative_mode
        self.symmetric = symmetric

    def forward(self, query, positive_key, negative_keys=None):
        return self.info_nce(query, positive_key, negative_keys, temperature=self.temperature, reduction=self.reduction,
                             negative_mode=self.negative_mode, symmetric=self.symmetric)

    def info_nce(self, query, positive_key, negative_keys=None, temperature=0.1, reduction="mean", negative_mode="unpaired",
                 symmetric=False):
        if query.dim() != 2:
            raise ValueError("<query> must have 2 dimensions.")
        if positive_key.dim() != 2:
            raise ValueError("<positive_key> must have 2 dimensions.")
        if negative_keys is not None:
            if negative_mode == "unpaired" and negative_keys.dim() != 2:
                raise ValueError("<negative_keys> must have 2 dimensions if <negative_mode> == 'unpaired'.")
            if negative_mode == "paired" and negative_keys.dim() != 3:
                raise ValueError("<negative_keys> must have 3 dimensions if <negative_mode> == 'paired'.")
        if len(query) != len(positive_key):
            raise ValueError("<query> and <positive_key> must must have the same number of samples.")
        if negative_keys is not None:
            if negative_mode == "paired" and len(query) != len(negative_keys):
                raise ValueError("If negative_mode == 'paired', then <negative_keys> must have the same number of samples as <query>.")
        if query.shape[-1] != positive_key.shape[-1]:
            raise ValueError("Vectors of <query> and <positive_key> should have the same number of components.")
        if negative_keys is not None:
            if query.shape[-1] != negative_keys.shape[-1]:
                raise ValueError("Vectors of <query> and <negative_keys> should have the same number of components.")
            raise NotImplementedError("explicit negative_keys: the reference branch never assigns `loss` "
                                      "(losses/info_nce.py:126-143,166); only implicit negatives are implemented")
        if reduction not in ("mean", "sum", "none"):  # F.cross_entropy's own check (losses/info_nce.py:155-164)
            raise ValueError(f"{reduction} is not a valid value for reduction")
        q = ops.l2_normalize(query.float(), 1e-12)  # F.normalize default eps
        k = ops.l2_normalize(positive_key.float(), 1e-12)
        scale = torch.full((), 1.0 / temperature, device=query.device, dtype=torch.float32)
        from .mirror_loss import _negatives_group
        group = _negatives_group(getattr(self, "global_negatives", False), getattr(self, "group", None))
        return ops.clip_loss(q, k, scale, 0.5, 0.5, reduction, group) if symmetric else ops.clip_loss(q, k, scale, 1.0, 0.0, reduction, group)
