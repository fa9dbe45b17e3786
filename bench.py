#!/usr/bin/env python
"""MIRROR pre-training step benchmark (forward + loss + backward, slides/s).

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels behind the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (CPU baseline arm)

Workload (config.workload): C3 of SURVEY.md §8 — train_mirror.py, full alignment + retention + style + cluster losses,
768-d Phikon-shaped patch features, 2048 patches per slide, 10 234-d RNA vector, embed 768, 64 slides per GPU,
synthetic N(0,1) inputs, random-init weights, train mode (dropout on), weak scaling over GPUs.
A "step" = MIRROR.forward + MIRRORLoss + backward on one batch (optimizer excluded, as in BASELINE.json's metric).
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


def cpu_reference_step(B, N, Dw, Dr, steps, warmup, threads):
    """The reference algorithm (oracle/mirror_oracle.py: a restatement pinned against the reference sources) on the host
    cores: forward + MIRRORLoss + backward in fp32 (dropout off: the masks cost the CPU nothing measurable).  Returns
    (slides/s, seconds per step)."""
    import torch
    from oracle import mirror_oracle as O
    torch.set_num_threads(threads)
    cfg = O.default_cfg(Dw=Dw, Dr=Dr, N=N)
    sd = {k: v.requires_grad_(True) for k, v in O.make_state_dict(cfg, 0).items()}
    wsi, rna = O.make_inputs(B, N, Dw, Dr, 1234)
    noise = O.make_noise(B, N, cfg["E"], cfg["latent"], 4321)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = O.mirror_forward(sd, wsi, rna, noise)
        total = O.mirror_loss(out)[0]
        O.grads_of(total, sd)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return B * len(times) / sum(times), sum(times) / len(times)


def run_eager(args):
    """The GPU bar (SURVEY.md §8d last row): the reference algorithm (oracle restatement, pinned against the reference
    sources) as PyTorch eager on the same B200 -- cuBLAS / cuDNN / ATen kernels under torch.autocast(bf16), TF32 matmuls
    enabled like train_mirror.py:650-652.  Dropout off (favours this arm).  Baseline only: none of this repo's kernels run."""
    import torch
    from oracle import mirror_oracle as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    B, N, Dw, Dr = args.batch, args.patches, args.wsi_dim, args.rna_dim
    cfg = O.default_cfg(Dw=Dw, Dr=Dr, N=N)
    sd = {k: v.to(dev).requires_grad_(True) for k, v in O.make_state_dict(cfg, 0).items()}
    wsi, rna = (t.to(dev) for t in O.make_inputs(B, N, Dw, Dr, 1234))
    noise = {k: v.to(dev) for k, v in O.make_noise(B, N, cfg["E"], cfg["latent"], 4321).items()}
    adt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": None}[args.eager_dtype]

    def step():
        with torch.autocast("cuda", dtype=adt, enabled=adt is not None):
            out = O.mirror_forward(sd, wsi, rna, noise)
            total = O.mirror_loss(out)[0]
        O.grads_of(total, sd)
        return total

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
    f_slide = algorithmic_gflop_per_slide(N, Dw)
    val = B / (ms / 1e3)
    print(json.dumps({"impl": "eager", "metric": "MIRROR pretrain slides/s fwd+bwd", "value": val, "unit": "slides/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
                      "dtype": args.eager_dtype, "data": "synthetic", "config": workload_config(args, B) | {"mode": "dropout off, fwd+loss+bwd, PyTorch eager (cuBLAS/cuDNN/ATen), autocast " + args.eager_dtype},
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30, "last_loss": float(loss),
                      "step_algorithmic_tflops": val * f_slide / 1e3, "step_algorithmic_frac": val * f_slide / 1e3 / sustained}), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    B = args.cpu_batch
    val, sec = cpu_reference_step(B, args.patches, args.wsi_dim, args.rna_dim, args.steps, max(1, min(args.warmup, 1)), cores)
    line = {
        "impl": "reference", "metric": "MIRROR pretrain slides/s fwd+bwd", "value": val, "unit": "slides/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(1, min(args.warmup, 1)), "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, B),
        "cpu_baseline": {"value": val, "unit": "slides/s", "cores": cores, "kind": "port",
                         "sample": f"{B} slides per step of the same workload (N={args.patches}, Dw={args.wsi_dim}), fp32 oracle port, dropout off"},
        "e2e": {"value": val, "unit": "slides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, B):
    return {"workload": "C3 train_mirror full step (alignment+retention+style+cluster), Phikon-shaped 768-d features",
            "slides_per_gpu": B, "patches_per_slide": args.patches, "wsi_dim": args.wsi_dim, "rna_dim": args.rna_dim, "embed_dim": 768,
            "mode": "train (dropout on), fwd+loss+bwd, optimizer excluded", "parallelism": f"dp{args.gpus}",
            "l2_policy": "inputs+activations (>10 GB per step) far exceed the 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager"])
    ap.add_argument("--eager-dtype", default="bf16", choices=["bf16", "fp16", "fp32"], help="autocast dtype of --impl eager")
    ap.add_argument("--batch", type=int, default=64, help="slides per GPU")
    ap.add_argument("--patches", type=int, default=2048)
    ap.add_argument("--wsi-dim", type=int, default=768)
    ap.add_argument("--rna-dim", type=int, default=10234)
    ap.add_argument("--cpu-batch", type=int, default=8, help="slides per step of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eval-mode", action="store_true", help="dropout off")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "eager":
        return run_eager(args)

    import torch
    import torch.distributed as dist

    from mirror_b200 import kernels as K
    from mirror_b200.losses import MIRRORLoss
    from mirror_b200.models import MIRROR

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
    model = MIRROR(wsi_embed_dim=Dw, rna_embed_dim=Dr, embed_dim=768, wsi_num_tokens=N, rna_mlp_ratio=4.0,
                   rna_norm_layer="layernorm", rna_act_layer="gelu").to(dev)
    model.train(not args.eval_mode)
    loss_fn = MIRRORLoss().to(dev)
    net = model
    if world > 1:  # the reference's own data-parallel mechanism (train_mirror.py:811-813): bucketed gradient all-reduce overlapped with backward
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)

    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 2  # two pinned host batches, alternated, so that every step really moves fresh bytes
    host = [(torch.randn(B, N, Dw, generator=g).pin_memory(), torch.randn(B, Dr, generator=g).pin_memory()) for _ in range(n_host)]
    wsi_d, rna_d = host[0][0].to(dev), host[0][1].to(dev)

    def step(wsi, rna):
        for p in model.parameters():
            p.grad = None
        out = net(wsi, rna, 0.75, 0.75)
        losses = loss_fn(*out)
        losses[0].backward()
        return losses[0]

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
        host_loss = float(loss)  # D2H read of the step's result (the trainer's loss.item(), train_mirror.py:1258)
    e3.record()
    barrier()
    t = torch.tensor([e2.elapsed_time(e3)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t) / args.steps
    e2e_val = world * B / (e2e_ms / 1e3)

    # ---- roofline of the dominant kernel (the tcgen05 GEMM): one instrumented step with CUDA events around every launch
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
    ei0.record()
    step(wsi_d, rna_d)
    ei1.record()
    torch.cuda.synchronize()
    K.gemm = real_gemm
    gemm_ms = sum(s_.elapsed_time(e_) for s_, e_, _, _ in rec)
    gemm_flop = sum(f for _, _, f, _ in rec)
    if os.environ.get("MIRROR_BENCH_VERBOSE") and rank == 0:
        agg = {}
        for s_, e_, f, sig in rec:
            t_ = agg.setdefault(sig, [0, 0.0, 0.0])
            t_[0] += 1
            t_[1] += s_.elapsed_time(e_)
            t_[2] += f
        print(f"GEMM launches of one step: {len(rec)}, {gemm_ms:.2f} ms of {ei0.elapsed_time(ei1):.2f} ms", file=sys.stderr)
        print("batch      M      N      K  AB  outs      sk terms  count       ms   TFLOP/s", file=sys.stderr)
        for sig, (c, ms_, f) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            print(f"{sig[0]:5d} {sig[1]:6d} {sig[2]:6d} {sig[3]:6d}  {sig[4]}{sig[5]}  {sig[6]:3s} {sig[7]:3s} {sig[8]:3d} {sig[9]:5d} {c:6d} {ms_:8.3f} {f / ms_ / 1e9:9.0f}",
                  file=sys.stderr)
    inst_ms = ei0.elapsed_time(ei1)
    achieved = gemm_flop / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    f_slide = algorithmic_gflop_per_slide(N, Dw)
    roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all launches of one step)", "achieved": achieved, "peak": sustained,
                "unit": "TFLOP/s", "frac": achieved / sustained, "traffic": None, "peak_source": f"{src} sustained bf16 (MEASURED_PEAKS.json)",
                "launches_per_step": len(rec), "avg_launch_ms": gemm_ms / max(len(rec), 1), "gemm_share_of_step": gemm_ms / inst_ms,
                "executed_gemm_gflop_per_slide": gemm_flop / B / 1e9, "algorithmic_gflop_per_slide": f_slide,
                "step_algorithmic_tflops": value / world * f_slide / 1e3, "step_algorithmic_frac": value / world * f_slide / 1e3 / sustained}

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {"metric": "MIRROR pretrain slides/s fwd+bwd", "value": value, "unit": "slides/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": workload_config(args, B), "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "slides/s", "h2d_bytes_per_step": (B * N * Dw + B * Dr) * 4, "d2h_bytes_per_step": 4,
                        "ms_per_step": e2e_ms, "last_loss": host_loss},
                "gpu_launches": launches, "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            t0 = time.time()
            cv, csec = cpu_reference_step(args.cpu_batch, N, Dw, Dr, 3, 1, cores)
            line["cpu_baseline"] = {"value": cv, "unit": "slides/s", "cores": cores, "kind": "port",
                                    "sample": f"1 warm-up + 3 timed steps of {args.cpu_batch} slides (same N/Dw/Dr, fp32 oracle port, dropout off, "
                                              f"{time.time() - t0:.0f} s of CPU work)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
