"""MIRRORClassifier (models/mirror.py:921-1015): the downstream model on the same kernels.  Oracle: the
reference algorithm's encoders (oracle.cls_encoder / rna_encoder) + the fusion and linear head in torch."""
import pytest
import torch

from oracle import mirror_oracle as O
import emu_backend
import parity


def _build(cfg, sd, fusion, classes, device):
    from mirror_b200.models import MIRRORClassifier
    torch.manual_seed(1234)  # the head is not part of the checkpoint: fixed init
    model = MIRRORClassifier(cfg["Dw"], cfg["Dr"], cfg["E"], classes, rna_mlp_ratio=cfg["mlp_ratio"], rna_norm_layer="layernorm",
                             rna_act_layer="gelu", fusion=fusion)
    own = model.state_dict()
    enc = {k: v for k, v in sd.items() if k in own}
    missing, unexpected = model.load_state_dict(enc, strict=False)   # the split pre-training checkpoint carries no head
    assert sorted(missing) == ["head.bias", "head.weight"] and not unexpected
    return model.to(device).eval()


def _oracle(sd, head_w, head_b, wsi, rna, fusion):
    w = O.cls_encoder(sd, wsi)
    if rna is None:
        return w @ head_w.T + head_b
    r = O.rna_encoder(sd, rna)
    f = w + r if fusion == "add" else torch.cat((w, r), 1)
    return f @ head_w.T + head_b


def _case(device, over, B, classes, seed, logit_tol=1e-2):
    cfg = O.default_cfg(**over)
    sd = O.make_state_dict(cfg, seed)
    wsi, rna = O.make_inputs(B, cfg["N"], cfg["Dw"], cfg["Dr"], seed + 1)
    y = torch.randint(0, classes, (B,), generator=torch.Generator().manual_seed(seed))
    for fusion in ("concat", "add"):
        model = _build(cfg, sd, fusion, classes, device)
        for with_rna in (True, False):
            if not with_rna and fusion == "concat":
                continue  # the reference's head would not fit a lone WSI embedding either
            model.zero_grad(set_to_none=True)
            logits = model(wsi.to(device), rna.to(device) if with_rna else None)
            torch.nn.functional.cross_entropy(logits, y.to(device)).backward()
            sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
            hw = model.head.weight.detach().cpu().clone().requires_grad_(True)
            hb = model.head.bias.detach().cpu().clone().requires_grad_(True)
            ol = _oracle(sdo, hw, hb, wsi, rna if with_rna else None, fusion)
            torch.nn.functional.cross_entropy(ol, y).backward()
            assert logits.shape == (B, classes)
            assert parity.rel(logits.detach().cpu(), ol.detach()) <= logit_tol, (fusion, with_rna, parity.rel(logits.detach().cpu(), ol.detach()))
            # head gradients: sums of (softmax - onehot) over 3-4 samples, i.e. differences of O(1) terms: 3e-2
            assert parity.rel(model.head.weight.grad.cpu(), hw.grad) <= 3e-2 and parity.rel(model.head.bias.grad.cpu(), hb.grad) <= 3e-2
            keys = [k for k, p in model.named_parameters() if k in sdo and p.grad is not None and sdo[k].grad is not None]
            gp = torch.cat([dict(model.named_parameters())[k].grad.flatten().cpu() for k in keys])
            go = torch.cat([sdo[k].grad.flatten() for k in keys])
            assert parity.rel(gp, go) <= 1e-2, (fusion, with_rna, parity.rel(gp, go))


def test_classifier_matches_oracle_cpu():
    emu_backend.use()
    try:
        # miniature, degenerate Nystrom corner (61 tokens -> n = m = 96) and O(0.2) logits out of 192 cancelling products: the
        # embedding tolerance (cosine >= 0.999) allows ~1e-2 relative on them; the E = 768 GPU case holds 1e-2
        _case("cpu", dict(Dw=40, Dr=77, E=192, N=60), 3, 4, 61, logit_tol=3e-2)
    finally:
        emu_backend.release()


@pytest.mark.gpu
@pytest.mark.parametrize("classes", [2, 5])
def test_classifier_matches_oracle_gpu(classes):
    _case("cuda", dict(Dw=96, Dr=300, E=768, N=500), 4, classes, 62)


def test_factory_filters_unknown_kwargs():
    from mirror_b200.models import mirror_classifier
    m = mirror_classifier(wsi_embed_dim=8, rna_embed_dim=10, embed_dim=24, num_classes=3, pretrained=False, pretrained_cfg=None)
    assert m.head.weight.shape == (3, 48)
