"""Load the UNMODIFIED reference sources from the read-only mount.  TEST
INFRASTRUCTURE, container-only: ``/root/reference`` does not exist on the GPU
box, so nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this.

``models/mirror.py`` imports ``timm`` and ``nystrom_attention`` which are not
installed; ``oracle/shims`` provides stand-ins for exactly the
symbols used.  The reference files are exec'd where they lie — never copied.
"""
import importlib.util
import os
import sys

REFERENCE_ROOT = os.environ.get("MIRROR_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "mirror.py"))


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns (models.mirror, losses.mirror_loss, losses.info_nce) modules."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    return (_load("_ref_models_mirror", "models/mirror.py"),
            _load("_ref_losses_mirror_loss", "losses/mirror_loss.py"),
            _load("_ref_losses_info_nce", "losses/info_nce.py"))


def pin_noise(ref_mirror_module, model, noise):
    """Make the reference draw OUR noise: ``torch.rand`` inside the two
    random_masking methods and ``Normal.rsample`` in ``reparameterize``
    (models/mirror.py:516,630,830-833; call order SURVEY.md §3.3)."""
    import torch

    eps_queue = [noise["wsi_eps"], noise["rna_eps"]]
    state = {"i": 0}

    def reparameterize(self, mu, logstd):
        e = eps_queue[state["i"] % 2]
        state["i"] += 1
        return mu + torch.exp(0.5 * logstd) * e.to(mu.dtype)

    ref_mirror_module.MIRROR.reparameterize = reparameterize

    rand_queue = [noise["wsi_mask"], noise["rna_mask"]]
    rstate = {"i": 0}
    real_rand = torch.rand

    class _TorchProxy:
        def __getattr__(self, k):
            return getattr(torch, k)

        @staticmethod
        def rand(*shape, **kw):
            t = rand_queue[rstate["i"] % 2]
            rstate["i"] += 1
            assert tuple(t.shape) == tuple(shape), (t.shape, shape)
            return t.clone()

    ref_mirror_module.torch = _TorchProxy()

    def reset():
        state["i"] = 0
        rstate["i"] = 0

    return reset, real_rand
