from .mirror import MIRROR, MIRRORDualEncoder, mirror, mirror_dual_encoder

__all__ = ["MIRROR", "MIRRORDualEncoder", "mirror", "mirror_dual_encoder"]
