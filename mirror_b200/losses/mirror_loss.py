"""B200-native MIRROR pre-training loss — drop-in for the reference ``losses/mirror_loss.py``.

``ClipLoss`` (reference :16-52) and ``MIRRORLoss`` (:55-135) keep their constructor arguments, call
signatures and return values; the arithmetic (logits GEMM, row/column log-sum-exp, masked MSE,
Gaussian KL, symmetric KL and all of their gradients) runs in the kernels of ``csrc/loss.cu``.
"""
import torch
from torch import nn

from .. import ops


def _negatives_group(global_negatives, group):
    """None (rank-local negatives, the reference's behaviour under DDP: SURVEY.md fact 5) or the process group whose ranks
    pool their embeddings as negatives."""
    if not global_negatives:
        return None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None
    group = group if group is not None else dist.group.WORLD
    return group if dist.get_world_size(group) > 1 else None


class ClipLoss(nn.Module):
    """``global_negatives`` / ``group`` are extensions (default off = reference behaviour): the contrastive term is taken
    against the embeddings of every rank (embedding + log-sum-exp all-gathers, see ops.ClipLossFn)."""

    def __init__(self, cache_labels=False, global_negatives=False, group=None):
        super().__init__()
        self.cache_labels = cache_labels  # labels are implicit (the diagonal) in the fused kernels
        self.prev_num_logits = 0
        self.labels = {}
        self.global_negatives, self.group = global_negatives, group

    def forward(self, wsi_features, rna_features, logit_scale, output_dict=False):
        if not torch.is_tensor(logit_scale):
            logit_scale = torch.tensor(float(logit_scale), device=wsi_features.device)
        loss = ops.clip_loss(wsi_features, rna_features, logit_scale.float(), 0.5, 0.5,
                             group=_negatives_group(self.global_negatives, self.group))
        return {"contrastive_loss": loss} if output_dict else loss


class MIRRORLoss(nn.Module):
    def __init__(self, clip_loss_cache_labels=True, alignment_loss_weight=0.5, wsi_retention_loss_weight=0.1,
                 rna_retention_loss_weight=0.1, style_loss_weight=0.1, cluster_loss_weight=0.2, global_negatives=False, group=None):
        super().__init__()
        self.clip_loss = ClipLoss(cache_labels=clip_loss_cache_labels, global_negatives=global_negatives, group=group)
        self.alignment_loss_weight = alignment_loss_weight
        self.wsi_retention_loss_weight = wsi_retention_loss_weight
        self.rna_retention_loss_weight = rna_retention_loss_weight
        self.style_loss_weight = style_loss_weight
        self.cluster_loss_weight = cluster_loss_weight

    def forward(self, wsi_alignment_emb, wsi_retention_emb, wsi_retention_target, wsi_mask, wsi_score, wsi_mu, wsi_logstd,
                rna_alignment_emb, rna_retention_emb, rna_retention_target, rna_mask, rna_score, rna_mu, rna_logstd,
                logit_scale):
        B = wsi_alignment_emb.shape[0]
        alignment_loss = self.clip_loss(wsi_alignment_emb, rna_alignment_emb, logit_scale)
        wsi_retention_loss = ops.MaskedMseFn.apply(wsi_retention_emb, wsi_retention_target, wsi_mask)
        E = rna_retention_emb.shape[1]
        rna_retention_loss = ops.MaskedMseFn.apply(rna_retention_emb.reshape(B, E, 1), rna_retention_target.reshape(B, E, 1),
                                                   rna_mask)
        mu = ops.stack_rows(wsi_mu, rna_mu)
        logstd = ops.stack_rows(wsi_logstd, rna_logstd)
        style_loss = ops.GaussKlFn.apply(mu, logstd, B)
        cluster_loss = ops.SymKlFn.apply(ops.stack_rows(wsi_score, rna_score), B)
        terms = ops.StackScalarsFn.apply(alignment_loss, wsi_retention_loss, rna_retention_loss, style_loss, cluster_loss)
        total_loss = ops.CombineFn.apply(terms, (self.alignment_loss_weight, self.wsi_retention_loss_weight,
                                                 self.rna_retention_loss_weight, self.style_loss_weight,
                                                 self.cluster_loss_weight))
        return total_loss, alignment_loss, wsi_retention_loss, rna_retention_loss, style_loss, cluster_loss
