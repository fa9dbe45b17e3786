from .mirror import MIRROR, MIRRORClassifier, MIRRORDualEncoder, mirror, mirror_classifier, mirror_dual_encoder

__all__ = ["MIRROR", "MIRRORClassifier", "MIRRORDualEncoder", "mirror", "mirror_classifier", "mirror_dual_encoder"]
