"""The training step of ``train_mirror.py`` / ``train_pretrain.py`` as ONE replayable unit.

``GraphedStep`` runs, in the reference trainer's order (train_mirror.py:1133-1136, 1144-1230, 1254-1274):

    prototype row-normalisation -> forward (autocast-free: the kernels own their precision plan) -> loss -> backward
    -> gradient averaging over the data-parallel group (bucketed all-reduces of a flat buffer on a side stream, launched by
    parameter hooks while the backward is still running) -> clip_grad (mode "norm") -> Adam
    -> logit_scale clamp -> the six loss scalars + exp(logit_scale) + gradient norm packed into one tensor (one D2H copy
    instead of the trainer's seven ``.item()`` syncs)

and captures the whole sequence into one CUDA graph, so a step costs one ``cudaGraphLaunch`` on the host instead of ~630
kernel launches (36 ms of Python / driver time per step: smaller per-GPU batches were host-bound).  What makes the capture
valid:
  * parameters and Adam moments live in flat fp32 buffers (``FlatParams``; ``p.data`` are views) and the gradients are gathered
    into one with a multi-tensor copy, so the optimizer is one HBM-bound kernel (csrc/optim.cu) and the gradient exchange one
    NCCL call;
  * every per-step scalar (learning rate, step count, clip coefficient) is read from device memory;
  * dropout masks are counter-based hashes of (seed, element); a device-side epoch counter is mixed into the seeds of the
    captured launches and bumped once per replay (``kernels.set_dropout_epoch``), torch's own generator (mask noise, the
    reparameterisation draws) is graph-aware already;
  * inputs are copied into static buffers before each replay.
``graph=False`` runs the same sequence eagerly (debugging, CPU unit tests with the emulated kernels).
"""
import math

import torch

from . import kernels as K

F32 = torch.float32
_ALIGN = 64  # elements: every parameter starts on a 256-byte boundary inside the flat buffer (TMA / vector loads)


class FlatParams:
    """Re-homes the trainable parameters of ``model`` into one flat fp32 buffer (and their gradients into another).
    ``state_dict()`` / ``load_state_dict()`` keep working: the parameters are the same objects, only their storage moved."""

    def __init__(self, model):
        self.params = [p for p in model.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("model has no trainable parameters")
        dev = self.params[0].device
        if any(p.dtype != F32 or p.device != dev for p in self.params):
            raise ValueError("FlatParams needs fp32 parameters on one device")
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = n
        self.data = torch.zeros(n, device=dev, dtype=F32)
        self.grad = torch.zeros(n, device=dev, dtype=F32)
        self.grad_views = []
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                v = self.data[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                self.grad_views.append(self.grad[o:o + p.numel()].view(p.shape))

    def gather_grads(self):
        """flat.grad <- the parameters' .grad tensors, as one multi-tensor copy (parameters without a gradient: zeros).  Autograd
        ASSIGNS a fresh gradient when .grad is None; accumulating into flat views instead would cost one add kernel per parameter
        (~230 launches per step)."""
        with torch.no_grad():
            dst = [v for v, p in zip(self.grad_views, self.params) if p.grad is not None]
            src = [p.grad for p in self.params if p.grad is not None]
            if len(dst) != len(self.params):
                self.grad.zero_()
            torch._foreach_copy_(dst, src)

    def view_of(self, p, flat):
        i = next(i for i, q in enumerate(self.params) if q is p)
        return flat[self.offsets[i]:self.offsets[i] + p.numel()].view(p.shape)


class GraphedStep:
    """fwd + loss + bwd (+ gradient all-reduce, clip, Adam, clamp) of MIRROR / the dual encoder, captured in a CUDA graph.

    model      : ``MIRROR`` (loss_fn = ``MIRRORLoss``) or ``MIRRORDualEncoder`` with ``dual=True`` (loss_fn = ``InfoNCE``)
    example    : (wsi [B,N,Dw], rna [B,Dr]) device tensors fixing the static shapes
    group      : data-parallel process group (None = single process); gradients are averaged over it like DDP does: buckets of
                 ``bucket_mb`` MB of the flat gradient buffer (last parameters first) are all-reduced on a side stream as soon as
                 the backward has produced their gradients, overlapping the rest of the backward -- also inside the graph
    optimizer  : None (forward + backward only, the benchmark's "optimizer excluded" step) or a dict
                 ``lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False`` (torch.optim.Adam / AdamW semantics)
    clip_grad  : max gradient norm (clip_mode "norm") or None
    noise      : optional dict of STATIC noise tensors (parity tests); by default the model draws its own
    """

    STATS = ("total", "alignment", "wsi_retention", "rna_retention", "style", "cluster", "logit_scale_exp", "grad_norm")

    def __init__(self, model, loss_fn, example, dual=False, group=None, optimizer=None, clip_grad=None, mask_ratios=(0.75, 0.75),
                 noise=None, graph=True, warmup=3, bucket_mb=50):
        self.model, self.loss_fn, self.dual, self.group = model, loss_fn, dual, group
        self.opt, self.clip_grad, self.ratios, self.noise = optimizer, clip_grad, mask_ratios, noise
        wsi, rna = example
        dev = wsi.device
        self.flat = FlatParams(model)
        self.wsi = torch.empty_like(wsi, dtype=F32)
        self.rna = torch.empty_like(rna, dtype=F32)
        self.stats = torch.zeros(len(self.STATS), device=dev, dtype=F32)
        self.world = 1
        if group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(group)
        if optimizer is not None:
            o = dict(betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False)
            o.update(optimizer)
            self.opt = o
            self.m = torch.zeros_like(self.flat.data)
            self.v = torch.zeros_like(self.flat.data)
            self.lr = torch.full((1,), float(o["lr"]), device=dev, dtype=F32)
            self.t = torch.zeros(1, device=dev, dtype=F32)
            self.coef = torch.ones(1, device=dev, dtype=F32)
        self.epoch = torch.zeros(1, device=dev, dtype=torch.int64) if dev.type == "cuda" else None
        self.graph = None
        self.kernels_per_replay = 0
        self._hooks = []
        if group is not None:
            self._make_buckets(bucket_mb)
        if graph:
            if dev.type != "cuda":
                raise RuntimeError("CUDA graph capture needs CUDA tensors")
            self._capture(example, warmup)

    # ------------------------------------------------------------------------------------------------------------
    def _make_buckets(self, bucket_mb):
        """Contiguous slices of the flat gradient buffer, filled from the LAST parameter backwards (the order the backward produces
        gradients in, roughly), each reduced as soon as all of its parameters have their gradient."""
        import torch.distributed as dist
        flat = self.flat
        cap = max(1, int(bucket_mb)) * (1 << 20) // 4
        self.buckets = []  # dicts: lo, hi (element range of flat.grad), idx (parameter indices)
        hi = flat.numel
        idx = []
        for i in range(len(flat.params) - 1, -1, -1):
            idx.append(i)
            lo = flat.offsets[i]
            if hi - lo >= cap or i == 0:
                self.buckets.append(dict(lo=lo, hi=hi, idx=idx))
                hi, idx = lo, []
        self._bucket_of = {}
        for b, bk in enumerate(self.buckets):
            for i in bk["idx"]:
                self._bucket_of[i] = b
        self._pending = [0] * len(self.buckets)
        self._done = [True] * len(self.buckets)
        dev = flat.data.device
        self._side = torch.cuda.Stream(dev) if dev.type == "cuda" else None
        self._avg = dist.get_backend(self.group) == "nccl"  # ReduceOp.AVG exists for NCCL only
        for i, p in enumerate(flat.params):
            self._hooks.append(p.register_post_accumulate_grad_hook(lambda _p, i=i: self._grad_ready(i)))

    def _grad_ready(self, i):
        b = self._bucket_of[i]
        if self._done[b]:
            return  # a backward outside step(): nothing to do
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._reduce_bucket(b)

    def _reduce_bucket(self, b):
        """gather the bucket's gradients into the flat buffer (one multi-tensor copy on the backward's stream), then all-reduce the
        slice on the side stream"""
        import torch.distributed as dist
        bk, flat = self.buckets[b], self.flat
        with torch.no_grad():
            have = [i for i in bk["idx"] if flat.params[i].grad is not None]
            if len(have) != len(bk["idx"]):
                flat.grad[bk["lo"]:bk["hi"]].zero_()
            if have:
                torch._foreach_copy_([flat.grad_views[i] for i in have], [flat.params[i].grad for i in have])
            sl = flat.grad[bk["lo"]:bk["hi"]]
            op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
            if self._side is None:
                dist.all_reduce(sl, op=op, group=self.group)
            else:
                cur = torch.cuda.current_stream()
                ev = torch.cuda.Event()
                ev.record(cur)
                with torch.cuda.stream(self._side):
                    self._side.wait_event(ev)
                    dist.all_reduce(sl, op=op, group=self.group)
        self._done[b] = True

    def _finish_reduction(self):
        for b in range(len(self.buckets)):
            if not self._done[b]:  # parameters without a gradient in this step
                self._reduce_bucket(b)
        if self._side is not None:
            ev = torch.cuda.Event()
            ev.record(self._side)
            torch.cuda.current_stream().wait_event(ev)
        if not self._avg:
            self.flat.grad.mul_(1.0 / self.world)

    def set_lr(self, lr: float):
        if self.opt is None:
            raise RuntimeError("GraphedStep was built without an optimizer")
        self.lr.fill_(float(lr))  # device scalar: the captured Adam launch reads it at replay time

    def _body(self):
        model, flat = self.model, self.flat
        with torch.no_grad():
            if self.opt is not None and not self.dual and hasattr(model, "normalize_prototypes"):
                model.normalize_prototypes()  # train_mirror.py:1133-1136
        for p in flat.params:
            p.grad = None  # autograd then assigns (no accumulate kernels); under capture the new tensors come from the graph's pool
        if self.dual:
            we, re_ = model(self.wsi, self.rna)
            losses = (self.loss_fn(we, re_),)
        else:
            out = model(self.wsi, self.rna, self.ratios[0], self.ratios[1], noise=self.noise) if self.noise is not None else \
                model(self.wsi, self.rna, self.ratios[0], self.ratios[1])
            losses = self.loss_fn(*out)
        if self.group is not None:  # arm the buckets: the parameters' post-accumulate hooks launch the all-reduces during the backward
            for b, bk in enumerate(self.buckets):
                self._pending[b] = len(bk["idx"])
                self._done[b] = False
        losses[0].backward()
        with torch.no_grad():
            if self.group is not None:
                self._finish_reduction()
            elif self.opt is not None:
                flat.gather_grads()
            if self.group is not None:
                if self.opt is None:  # the caller's own optimizer reads p.grad: hand back the averaged gradient, like DDP
                    torch._foreach_copy_([p.grad for p in flat.params if p.grad is not None],
                                         [v for v, p in zip(flat.grad_views, flat.params) if p.grad is not None])
            sumsq = None
            if self.opt is not None:
                if self.clip_grad is not None:
                    sumsq = K.grad_sumsq(flat.grad)
                ls = getattr(model, "logit_scale", None)
                # clip coefficient, step count += 1 (before Adam, which needs the 1-based count)
                K.tail_scalars_(sumsq, float(self.clip_grad or 0.0), self.coef, self.t)
                o = self.opt
                K.adam_step_(flat.data, flat.grad, self.m, self.v, self.lr, o["betas"][0], o["betas"][1], o["eps"], o["weight_decay"],
                             o["decoupled"], self.t, self.coef if self.clip_grad is not None else None)
                if ls is not None and not self.dual:
                    K.tail_scalars_(clamp_param=ls.data.view(1), lo=0.0, hi=math.log(100.0))  # train_mirror.py:1254-1256
            vals = [l.detach().reshape(()) for l in losses]
            vals += [torch.zeros((), device=vals[0].device)] * (6 - len(vals))
            ls = getattr(model, "logit_scale", None)
            vals.append(ls.detach().exp().reshape(()) if ls is not None else torch.zeros((), device=vals[0].device))
            vals.append(sumsq.sqrt() if sumsq is not None else torch.zeros((), device=vals[0].device))
            self.stats.copy_(torch.stack(vals))
            if self.epoch is not None:
                self.epoch += 1

    def _capture(self, example, warmup):
        self.wsi.copy_(example[0])
        self.rna.copy_(example[1])
        K.set_dropout_epoch(self.epoch)
        try:
            # warm-up on a side stream (torch.cuda.graph's rule): first-call attribute setup, allocator pools, NCCL channels.
            # The warm-up steps are real steps when an optimizer is attached; their effect is undone below.
            snap = self.flat.data.clone() if self.opt is not None else None
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(max(warmup, 1)):
                    self._body()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            if snap is not None:
                self.flat.data.copy_(snap)
                self.m.zero_(), self.v.zero_(), self.t.zero_()
                del snap
            self.graph = torch.cuda.CUDAGraph()
            l0 = K.LAUNCHES[0]
            with torch.cuda.graph(self.graph):
                self._body()
            self.kernels_per_replay = K.LAUNCHES[0] - l0
        finally:
            K.set_dropout_epoch(None)  # launches outside the graph use their seeds as passed

    # ------------------------------------------------------------------------------------------------------------
    def step(self, wsi, rna):
        """One training step on (wsi, rna); returns the total loss as a 0-d device tensor (a view of ``self.stats``)."""
        if wsi.data_ptr() != self.wsi.data_ptr():
            self.wsi.copy_(wsi, non_blocking=True)
        if rna.data_ptr() != self.rna.data_ptr():
            self.rna.copy_(rna, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        return self.stats[0]

    def close(self):
        """Drop the captured graph.  Call it before ``destroy_process_group``: a live graph that holds NCCL kernels keeps the
        communicator busy and the teardown waits for it forever."""
        for h in self._hooks:
            h.remove()
        self._hooks = []
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None

    def read_stats(self):
        """{name: float} of the last step: ONE device->host copy (the trainer's seven .item() calls, train_mirror.py:1256-1274)."""
        return dict(zip(self.STATS, self.stats.tolist()))
