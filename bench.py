#!/usr/bin/env python
"""MIRROR pre-training step benchmark (forward + loss + backward, slides/s).

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels behind the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (CPU baseline arm)

    python bench.py --impl eager ...                         # the reference algorithm as PyTorch eager on the same GPU (the bar)

Default workload (config.workload): C3 of SURVEY.md §8 — train_mirror.py, full alignment + retention + style + cluster losses,
768-d Phikon-shaped patch features, 2048 patches per slide, 10 234-d RNA vector, embed 768, 64 slides per GPU,
synthetic N(0,1) inputs, random-init weights, train mode (dropout on), weak scaling over GPUs.
A "step" = MIRROR.forward + MIRRORLoss + backward on one batch (optimizer excluded, as in BASELINE.json's metric).

Other configurations of BASELINE.json (`--workload`; the driver only runs the default):
    c1  train_mirror, 1024-d ResNet-50-shaped features, 2048 patches, 4 slides (the reference's CPU-runnable case)
    c2  train_pretrain: dual encoder + InfoNCE(T=0.1), 768-d, 4096 patches, 32 slides
    c3  the default; `--patches 4096|16384` for the long bags (`--batch` to fit)
    c4  c3 at a fixed GLOBAL batch of 256 slides (`--global-batch`): B_local = 256 / N GPUs, strong scaling, global negatives
        (embedding + log-sum-exp all-gather, ops.ClipLossFn)
    c5  contrastive-loss microbench: fused logits + bidirectional softmax-CE, fwd+bwd, B = 256..8192, D = 512, against the
        reference InfoNCE / ClipLoss as PyTorch eager on the same GPU
The step runs through mirror_b200.step.GraphedStep by default (forward + loss + backward [+ the flat gradient all-reduce for
N > 1] replayed from ONE CUDA graph: one cudaGraphLaunch per step instead of ~630 kernel launches from Python); `--no-graph`
launches every kernel through the plain module API (DistributedDataParallel for N > 1), as the unchanged reference trainer would.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

F_STEP_GFLOP = {  # algorithmic fwd+bwd GFLOP per slide, SURVEY.md §8(d) closed form (Dw=768, E=768)
    2048: 417.3, 4096: 603.8, 16384: 1788.0}


def algorithmic_gflop_per_slide(N, Dw, E=768):
    import math
    H = math.ceil(math.sqrt(N))
    h, m = 8, E // 2

    def nys(S):
        n = m * math.ceil(S / m)
        return 6 * n * E * E + 2 * E * (2 * n * m + m * m) + 48 * h * m ** 3 + 2 * h * n * m * m + 2 * E * m * n + 2 * E * n * m + 2 * n * E * E + 66 * n * E

    fwd = 2 * N * Dw * E + 2 * nys(H * H + 1) + 166 * H * H * E + nys(N + 1) + 4 * (N + 1) * E * E + 0.09e9
    return 3 * fwd / 1e9


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p.get("bf16_tflops_sustained", 1400.8), p.get("bf16_tflops", 1697.1), p.get("hbm_gbs", 6539.2), "measured"
    except Exception:
        return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


WORKLOADS = {
    "c1": dict(batch=4, patches=2048, wsi_dim=1024, model="mirror",
               desc="C1 train_mirror full step (alignment+retention+style+cluster), ResNet-50-shaped 1024-d features"),
    "c2": dict(batch=32, patches=4096, wsi_dim=768, model="dual",
               desc="C2 train_pretrain step: dual encoder (FeatureTransMIL cls + TransFormer) + InfoNCE(T=0.1), Phikon-shaped 768-d features"),
    "c3": dict(batch=64, patches=2048, wsi_dim=768, model="mirror",
               desc="C3 train_mirror full step (alignment+retention+style+cluster), Phikon-shaped 768-d features"),
    "c4": dict(batch=None, patches=2048, wsi_dim=768, model="mirror",
               desc="C4 train_mirror full step at a fixed global batch, global-negative embedding all-gather, Phikon-shaped 768-d features"),
}


def resolve(args):
    w = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "c4":
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not a multiple of {world} ranks")
        args.batch = args.global_batch // world
    args.batch = args.batch if args.batch is not None else w["batch"]
    args.patches = args.patches if args.patches is not None else w["patches"]
    args.wsi_dim = args.wsi_dim if args.wsi_dim is not None else w["wsi_dim"]
    args.model = w["model"]
    return args


def gflop_per_slide(args):
    """algorithmic fwd+bwd GFLOP per slide, SURVEY.md §8(d) closed forms"""
    import math
    N, Dw, E = args.patches, args.wsi_dim, 768
    if args.model == "mirror":
        return algorithmic_gflop_per_slide(N, Dw)
    H = math.ceil(math.sqrt(N))
    h, m = 8, E // 2
    n = m * math.ceil((H * H + 1) / m)
    nys = 6 * n * E * E + 2 * E * (2 * n * m + m * m) + 48 * h * m ** 3 + 2 * h * n * m * m + 2 * E * m * n + 2 * E * n * m + 2 * n * E * E + 66 * n * E
    return 3 * (2 * N * Dw * E + 2 * nys + 166 * H * H * E + 0.06e9) / 1e9


def oracle_step_fn(args, device, B, dropout=False):
    """(step() -> loss, description) of the reference algorithm (oracle/mirror_oracle.py, pinned against the reference sources)
    for this workload on `device`; used by the CPU baseline arm and by the eager-GPU arm."""
    import torch
    from oracle import mirror_oracle as O
    N, Dw, Dr = args.patches, args.wsi_dim, args.rna_dim
    cfg = O.default_cfg(Dw=Dw, Dr=Dr, N=N)
    sd = {k: v.to(device).requires_grad_(True) for k, v in O.make_state_dict(cfg, 0).items()}
    wsi, rna = (t.to(device) for t in O.make_inputs(B, N, Dw, Dr, 1234))
    noise = {k: v.to(device) for k, v in O.make_noise(B, N, cfg["E"], cfg["latent"], 4321).items()}

    def step():
        if args.model == "dual":
            we, re_ = O.dual_encoder_forward(sd, wsi, rna)
            total = O.info_nce(we, re_, 0.1, False)
        else:
            total = O.mirror_loss(O.mirror_forward(sd, wsi, rna, noise))[0]
        O.grads_of(total, sd)
        return total
    return step


def cpu_reference_step(args, B, steps, warmup, threads):
    """The reference algorithm on the host cores: forward + loss + backward in fp32, dropout off (the masks cost the CPU
    nothing measurable).  Returns (slides/s, seconds per step)."""
    import torch
    torch.set_num_threads(threads)
    step = oracle_step_fn(args, "cpu", B)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return B * len(times) / sum(times), sum(times) / len(times)


METRIC = "MIRROR pretrain slides/s fwd+bwd"


def run_eager(args):
    """The GPU bar (SURVEY.md §8d last row): the reference algorithm (oracle restatement, pinned against the reference
    sources) as PyTorch eager on the same B200 -- cuBLAS / cuDNN / ATen kernels under torch.autocast(bf16), TF32 matmuls
    enabled like train_mirror.py:650-652.  Dropout off (favours this arm).  Baseline only: none of this repo's kernels run."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "c5":
        return run_c5(args)
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    B = args.batch
    adt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": None}[args.eager_dtype]
    inner = oracle_step_fn(args, dev, B)

    def step():
        with torch.autocast("cuda", dtype=adt, enabled=adt is not None):
            return inner()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    sustained = peaks()[0]
    f_slide = gflop_per_slide(args)
    val = B / (ms / 1e3)
    print(json.dumps({"impl": "eager", "metric": METRIC, "value": val, "unit": "slides/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
                      "dtype": args.eager_dtype, "data": "synthetic",
                      "config": workload_config(args, B) | {"mode": "dropout off, fwd+loss+bwd, PyTorch eager (cuBLAS/cuDNN/ATen), autocast " + args.eager_dtype},
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30, "last_loss": float(loss.detach()),
                      "step_algorithmic_tflops": val * f_slide / 1e3, "step_algorithmic_frac": val * f_slide / 1e3 / sustained}), flush=True)


def run_reference(args):
    """CPU baseline arm: the reference's algorithm for this workload on the box's host cores.  /root/reference is Python and
    does not exist on the GPU box, so what runs is the oracle port (kind "port": oracle/mirror_oracle.py, pinned against the
    unmodified reference sources by oracle/pin_against_reference.py), fp32, all host threads, dropout off."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    B = args.cpu_batch
    warm = max(1, min(args.warmup, 1))
    val, sec = cpu_reference_step(args, B, args.steps, warm, cores)
    cfg = workload_config(args, B) | {"mode": "dropout off, fwd+loss+bwd, fp32 oracle port of the reference on the host cores, optimizer excluded"}
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "slides/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": "slides/s", "cores": cores, "kind": "port",
                         "sample": f"{B} slides per step of the same workload (N={args.patches}, Dw={args.wsi_dim}), fp32 oracle port, dropout off"},
        "e2e": {"value": val, "unit": "slides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, B):
    cfg = {"workload": WORKLOADS[args.workload]["desc"],
           "slides_per_gpu": B, "patches_per_slide": args.patches, "wsi_dim": args.wsi_dim, "rna_dim": args.rna_dim, "embed_dim": 768,
           "mode": ("train (dropout on)" if not args.eval_mode else "eval (dropout off)") + ", fwd+loss+bwd, optimizer excluded",
           "parallelism": f"dp{args.gpus}",
           "l2_policy": "inputs+activations (GBs per step) far exceed the 126 MB L2; no explicit flush"}
    if args.workload == "c4":
        cfg["global_batch"] = args.global_batch
        cfg["negatives"] = "global (embedding + LSE all-gather)"
    if args.impl == "ours":
        cfg["launch"] = "one CUDA graph per step (GraphedStep)" if args.graph else "per-kernel launches from Python (module API)"
    return cfg


# ------------------------------------------------------------------------------------------------------------------
def run_c5(args):
    """C5: contrastive loss fwd+bwd at B = 256..8192, D = 512: the fused kernels (logits only in TMEM) against the reference
    losses as PyTorch eager on the same GPU (fp32 with TF32 matmuls, and bf16 autocast).  Roofline denominator 6 B^2 D."""
    import torch
    import torch.nn.functional as F
    from mirror_b200.losses import InfoNCE
    from mirror_b200 import ops, kernels as K
    from oracle import mirror_oracle as O
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    torch.backends.cuda.matmul.allow_tf32 = True
    sustained = peaks()[0]
    D = 512
    rows = []
    g = torch.Generator(device=dev).manual_seed(1234)

    def timeit(fn, iters):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / iters)
        return statistics.median(ts)

    for B in (256, 512, 1024, 2048, 4096, 8192):
        q = torch.randn(B, D, device=dev, generator=g, requires_grad=True)
        k = torch.randn(B, D, device=dev, generator=g, requires_grad=True)
        scale = torch.tensor(1 / 0.07, device=dev, requires_grad=True)
        ours_nce = InfoNCE(temperature=0.1, symmetric=True)

        def ours_infonce():
            q.grad = k.grad = None
            ours_nce(q, k).backward()

        qn, kn = F.normalize(q.detach(), dim=-1).requires_grad_(True), F.normalize(k.detach(), dim=-1).requires_grad_(True)

        def ours_clip():
            qn.grad = kn.grad = None
            ops.clip_loss(qn, kn, scale).backward()

        def ref_infonce():
            q.grad = k.grad = None
            O.info_nce(q, k, 0.1, True).backward()

        def ref_clip():
            qn.grad = kn.grad = None
            O.clip_loss(qn, kn, scale).backward()

        def ref_clip_bf16():
            qn.grad = kn.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16):
                l = O.clip_loss(qn, kn, scale)
            l.backward()

        iters = 50 if B <= 2048 else 20
        l0 = K.LAUNCHES[0]
        ours_clip()
        nl = K.LAUNCHES[0] - l0
        r = {"B": B, "D": D, "algorithmic_gflop": 6.0 * B * B * D / 1e9, "launches": nl,
             "ours_clip_ms": timeit(ours_clip, iters), "ours_infonce_ms": timeit(ours_infonce, iters),
             "eager_clip_tf32_ms": timeit(ref_clip, iters), "eager_clip_bf16_ms": timeit(ref_clip_bf16, iters),
             "eager_infonce_tf32_ms": timeit(ref_infonce, iters),
             "operands": "split-3 bf16 (fp32-grade)" if B <= ops.PRECISE_ROWS else "bf16"}
        r["ours_clip_tflops"] = r["algorithmic_gflop"] / r["ours_clip_ms"]
        r["ours_clip_frac_of_sustained"] = r["ours_clip_tflops"] / sustained
        r["speedup_vs_eager_tf32"] = r["eager_clip_tf32_ms"] / r["ours_clip_ms"]
        r["speedup_vs_eager_bf16"] = r["eager_clip_bf16_ms"] / r["ours_clip_ms"]
        rows.append(r)
        del q, k, qn, kn
    big = rows[-1]
    print(json.dumps({"metric": "contrastive loss fwd+bwd TFLOP/s (6 B^2 D) at B=8192, D=512", "value": big["ours_clip_tflops"], "unit": "TFLOP/s",
                      "n_gpus": 1, "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": "C5 contrastive-loss microbench: fused logits GEMM + bidirectional softmax-CE fwd+bwd, B=256..8192, D=512",
                                 "timing": "10 warm-up, median of 5 x (20-50) iterations, CUDA events"},
                      "roofline": {"bound": "tensor", "achieved": big["ours_clip_tflops"], "peak": sustained, "unit": "TFLOP/s",
                                   "frac": big["ours_clip_tflops"] / sustained, "traffic": None},
                      "rows": rows}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--eager-dtype", default="bf16", choices=["bf16", "fp16", "fp32"], help="autocast dtype of --impl eager")
    ap.add_argument("--batch", type=int, default=None, help="slides per GPU (default: the workload's)")
    ap.add_argument("--global-batch", type=int, default=256, help="c4: slides per step over all GPUs")
    ap.add_argument("--patches", type=int, default=None)
    ap.add_argument("--wsi-dim", type=int, default=None)
    ap.add_argument("--rna-dim", type=int, default=10234)
    ap.add_argument("--cpu-batch", type=int, default=8, help="slides per step of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eval-mode", action="store_true", help="dropout off")
    ap.add_argument("--graph", dest="graph", action="store_true", default=None,
                    help="replay the step from one CUDA graph (mirror_b200.step.GraphedStep); the default")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch every kernel from Python through the plain module API (under DistributedDataParallel for N > 1)")
    ap.add_argument("--bucket-mb", type=int, default=25, help="DDP gradient bucket size (bucket_cap_mb)")
    args = ap.parse_args()
    if args.graph is None:
        args.graph = args.impl == "ours"
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload == "c5":
        if int(os.environ.get("RANK", "0")) == 0:
            run_c5(args)
        return
    resolve(args)
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "eager":
        return run_eager(args)

    import torch
    import torch.distributed as dist

    from mirror_b200 import kernels as K
    from mirror_b200.losses import InfoNCE, MIRRORLoss
    from mirror_b200.models import MIRROR, MIRRORDualEncoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N, Dw, Dr = args.batch, args.patches, args.wsi_dim, args.rna_dim

    torch.manual_seed(0)
    if args.model == "dual":
        model = MIRRORDualEncoder(wsi_embed_dim=Dw, rna_embed_dim=Dr, embed_dim=768, rna_mlp_ratio=4.0,
                                  rna_norm_layer="layernorm", rna_act_layer="gelu").to(dev)
        loss_fn = InfoNCE(temperature=0.1).to(dev)
    else:
        model = MIRROR(wsi_embed_dim=Dw, rna_embed_dim=Dr, embed_dim=768, wsi_num_tokens=N, rna_mlp_ratio=4.0,
                       rna_norm_layer="layernorm", rna_act_layer="gelu").to(dev)
        loss_fn = MIRRORLoss(global_negatives=args.workload == "c4").to(dev)
    model.train(not args.eval_mode)
    net = model
    if world > 1 and not args.graph:  # the reference's own data-parallel mechanism (train_mirror.py:811-813): bucketed gradient all-reduce overlapped with backward
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True, bucket_cap_mb=args.bucket_mb)

    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 2  # two pinned host batches, alternated, so that every step really moves fresh bytes
    host = [(torch.randn(B, N, Dw, generator=g).pin_memory(), torch.randn(B, Dr, generator=g).pin_memory()) for _ in range(n_host)]
    wsi_d, rna_d = host[0][0].to(dev), host[0][1].to(dev)

    def eager_step(wsi, rna):
        for p in model.parameters():
            p.grad = None
        if args.model == "dual":
            we, re_ = net(wsi, rna)
            loss = loss_fn(we, re_)
        else:
            loss = loss_fn(*net(wsi, rna, 0.75, 0.75))[0]
        loss.backward()
        return loss

    step = eager_step
    if args.graph:
        from mirror_b200.step import GraphedStep
        gs = GraphedStep(model, loss_fn, (wsi_d, rna_d), dual=args.model == "dual",
                         group=dist.group.WORLD if world > 1 else None)
        step = gs.step
        wsi_d, rna_d = gs.wsi, gs.rna  # kernel-resident arm: the inputs already sit in the graph's static buffers

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-resident arm: inputs already in HBM
    for _ in range(args.warmup):
        step(wsi_d, rna_d)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = K.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step(wsi_d, rna_d)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = K.LAUNCHES[0] - l0
    if args.graph:
        launches = gs.kernels_per_replay * args.steps
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t) / args.steps
    value = world * B / (ms / 1e3)

    # ---- end-to-end arm: pinned host inputs -> H2D every step (prefetched on a copy stream), loss read back to the host
    copy_stream = torch.cuda.Stream()
    bufs = [(torch.empty(B, N, Dw, device=dev), torch.empty(B, Dr, device=dev)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            bufs[s][0].copy_(host[i % n_host][0], non_blocking=True)
            bufs[s][1].copy_(host[i % n_host][1], non_blocking=True)
            ready[s].record(copy_stream)

    for s in range(2):
        consumed[s].record()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    prefetch(0)
    host_loss = 0.0
    for i in range(args.steps):
        if i + 1 < args.steps:
            prefetch(i + 1)
        s = i % 2
        torch.cuda.current_stream().wait_event(ready[s])
        loss = step(bufs[s][0], bufs[s][1])
        consumed[s].record()
        host_loss = float(loss.detach())  # D2H read of the step's result (the trainer's loss.item(), train_mirror.py:1258)
    e3.record()
    barrier()
    t = torch.tensor([e2.elapsed_time(e3)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t) / args.steps
    e2e_val = world * B / (e2e_ms / 1e3)

    kernels_per_replay = gs.kernels_per_replay if args.graph else 0
    if args.graph:  # the timed arms are done: release the graph and its private memory pool (at 16 384 patches x 64 slides it holds
        gs.close()  # 87 GB, and the eagerly launched instrumented step below needs as much again)
        for p_ in model.parameters():
            p_.grad = None
        del gs, step, loss
        import gc
        gc.collect()
        torch.cuda.empty_cache()

    # ---- roofline: one instrumented (eager-launched) step with CUDA events around every tcgen05 GEMM launch
    sustained, burst, hbm, src = peaks()
    rec = []
    real_gemm = K.gemm

    def timed_gemm(a, b, **kw):
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        real_gemm(a, b, **kw)
        e_.record()
        bt = 1
        for d in a.shape[:-2]:
            bt *= d
        sig = (bt, a.shape[-2], b.shape[-2], a.shape[-1], "T" if a.stride(-1) != 1 else "N", "N" if b.stride(-1) != 1 else "T",
               "f32" if kw.get("out_f32") is not None else "", "b16" if kw.get("out_bf16") is not None else "", kw.get("split_k", 1))
        flop = 2.0 * bt * a.shape[-2] * a.shape[-1] * b.shape[-2]
        for (am, bm) in (kw.get("more") or ()):  # multi-term launches: every product counts
            flop += 2.0 * bt * am.shape[-2] * am.shape[-1] * bm.shape[-2]
        sig = sig + (1 + len(kw.get("more") or ()),)
        rec.append((s_, e_, flop, sig))

    K.gemm = timed_gemm
    import mirror_b200.ops as _ops
    _ops.K.gemm = timed_gemm
    ei0, ei1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eager_step(wsi_d, rna_d)      # allocator / first-call effects of the eagerly launched path (the timed arms may have replayed a graph)
    torch.cuda.synchronize()
    rec.clear()
    torch.cuda._sleep(int(1.2e8))  # ~60 ms of device spin: the host enqueues the whole instrumented step behind it, so no CUDA-event
    ei0.record()                   # interval contains host launch latency (the GPU never waits for the host inside the step)
    eager_step(wsi_d, rna_d)
    ei1.record()
    torch.cuda.synchronize()
    K.gemm = real_gemm
    gemm_ms = sum(s_.elapsed_time(e_) for s_, e_, _, _ in rec)
    gemm_flop = sum(f for _, _, f, _ in rec)
    agg = {}
    for s_, e_, f, sig in rec:
        t_ = agg.setdefault(sig, [0, 0.0, 0.0])
        t_[0] += 1
        t_[1] += s_.elapsed_time(e_)
        t_[2] += f
    if os.environ.get("MIRROR_BENCH_VERBOSE") and rank == 0:
        print(f"GEMM launches of one step: {len(rec)}, {gemm_ms:.2f} ms of {ei0.elapsed_time(ei1):.2f} ms", file=sys.stderr)
        print("batch      M      N      K  AB  outs      sk terms  count       ms   TFLOP/s", file=sys.stderr)
        for sig, (c, ms_, f) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            print(f"{sig[0]:5d} {sig[1]:6d} {sig[2]:6d} {sig[3]:6d}  {sig[4]}{sig[5]}  {sig[6]:3s} {sig[7]:3s} {sig[8]:3d} {sig[9]:5d} {c:6d} {ms_:8.3f} {f / ms_ / 1e9:9.0f}",
                  file=sys.stderr)
    inst_ms = ei0.elapsed_time(ei1)
    f_slide = gflop_per_slide(args)
    # the dominant single instantiation: the one-term batched 384^3 product of the Moore-Penrose chain
    # (gemm_tcgen05_kernel<192,0,1,true,2>: 128x192 tiles, bf16 in/out, fp32 TMEM accumulator; 72 launches per step)
    dom_sig, dom = max(agg.items(), key=lambda kv: kv[1][1])
    dom_c, dom_ms, dom_f = dom
    dom_ach = dom_f / (dom_ms / 1e3) / 1e12
    is_pinv = dom_sig[1:4] == (384, 384, 384) and dom_sig[9] == 1
    roofline = {"bound": "tensor",
                "kernel": ("gemm_tcgen05_kernel<192,0,1,true,2> " if is_pinv else "gemm_tcgen05_kernel ") +
                          f"[batch {dom_sig[0]} x {dom_sig[1]}x{dom_sig[2]}x{dom_sig[3]} {dom_sig[4]}{dom_sig[5]}, {dom_sig[9]} term(s)]: the launch signature with the largest share of the step",
                "achieved": dom_ach, "peak": sustained, "unit": "TFLOP/s", "frac": dom_ach / sustained,
                "traffic": 261.5e6 if is_pinv and dom_sig[0] == 512 else None,
                "traffic_source": "profiles/r2_ncu_pinv_gemm.md: dram read+write per launch of this instantiation (ncu --set full), vs 453 MB operands+result" if is_pinv and dom_sig[0] == 512 else None,
                "flop_per_launch": dom_f / dom_c, "avg_launch_ms": dom_ms / dom_c, "launches_per_step": dom_c,
                "share_of_step": dom_ms / inst_ms,
                "peak_source": f"{src} sustained bf16 (MEASURED_PEAKS.json)",
                "all_gemm": {"achieved": gemm_flop / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0,
                             "frac": (gemm_flop / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0) / sustained,
                             "launches_per_step": len(rec), "share_of_step": gemm_ms / inst_ms,
                             "executed_gflop_per_slide": gemm_flop / B / 1e9},
                "algorithmic_gflop_per_slide": f_slide,
                "step_algorithmic_tflops": value / world * f_slide / 1e3, "step_algorithmic_frac": value / world * f_slide / 1e3 / sustained}

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "slides/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong" if args.workload == "c4" else "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": workload_config(args, B), "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "slides/s", "h2d_bytes_per_step": (B * N * Dw + B * Dr) * 4, "d2h_bytes_per_step": 4,
                        "ms_per_step": e2e_ms, "last_loss": host_loss},
                "gpu_launches": launches, "roofline": roofline, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            t0 = time.time()
            cv, csec = cpu_reference_step(args, args.cpu_batch, 3, 1, cores)
            line["cpu_baseline"] = {"value": cv, "unit": "slides/s", "cores": cores, "kind": "port",
                                    "sample": f"1 warm-up + 3 timed steps of {args.cpu_batch} slides (same N/Dw/Dr, fp32 oracle port, dropout off, "
                                              f"{time.time() - t0:.0f} s of CPU work)"}
        print(json.dumps(line), flush=True)
    if world > 1:  # (the graph, which holds NCCL kernels, was released above: before the communicator goes away)
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
