from .info_nce import InfoNCE
from .mirror_loss import ClipLoss, MIRRORLoss

__all__ = ["InfoNCE", "MIRRORLoss", "ClipLoss"]
