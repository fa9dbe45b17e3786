"""mirror_b200 — B200-native (sm_100a) implementation of the MIRROR pre-training step.

Drop-in for the reference's ``models/mirror.py`` model API and the
``losses/mirror_loss.py`` / ``losses/info_nce.py`` loss interfaces: same class
names, constructor arguments, output tuples and ``state_dict`` keys.  All device
arithmetic runs in the hand-written CUDA kernels of ``csrc/`` behind the C ABI
declared in ``include/mirror_b200.h``; PyTorch provides memory, streams,
autograd bookkeeping and ``torch.distributed`` only.
"""
__version__ = "0.1.0"
