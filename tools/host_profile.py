"""cProfile of the host side of a training step (where do the 60 ms of enqueue time go?):  python tools/host_profile.py [B]"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mirror_b200.losses import MIRRORLoss  # noqa: E402
from mirror_b200.models import MIRROR  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N, Dw, Dr = 2048, 768, 10234
dev = torch.device("cuda")
torch.manual_seed(0)
model = MIRROR(wsi_embed_dim=Dw, rna_embed_dim=Dr, embed_dim=768, wsi_num_tokens=N, rna_mlp_ratio=4.0, rna_norm_layer="layernorm",
               rna_act_layer="gelu").to(dev).train()
loss_fn = MIRRORLoss().to(dev)
wsi, rna = torch.randn(B, N, Dw, device=dev), torch.randn(B, Dr, device=dev)


def step():
    for p in model.parameters():
        p.grad = None
    losses = loss_fn(*model(wsi, rna, 0.75, 0.75))
    losses[0].backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(4):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(30)
