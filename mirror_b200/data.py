"""Input side of the training step: what ``datasets/dataset_pretrain.py:150-167`` and the trainer's
``.to(device)`` (``train_mirror.py:1138-1139``) do per step, arranged so that the GPU never waits for it.

  * ``resample_indices``  -- the dataset's per-slide patch sampling (``np.random.choice(M_i, N, replace=M_i < N)``) as one
    index matrix for the whole batch (host side: a few thousand integers per slide);
  * ``kernels.gather_rows`` -- the gather itself on the device, over the PACKED patch features of the batch (each slide's
    M_i x Dw matrix once, fp32 or bf16 storage), instead of materialising ``[B, N, Dw]`` on the host: a slide with fewer
    than N patches is not duplicated before the copy, one with more is not copied whole;
  * ``SlidePrefetcher``   -- a ring of pinned host buffers and device buffers fed by a copy stream: batch i+1 travels
    (H2D) while step i computes; events order copy -> use -> reuse, no host synchronisation.

At 2 700 slides/s x 6.3 MB (fp32, N = 2048) the feature stream is 17 GB/s per GPU -- within PCIe 5 x16 only when the
copies are pinned, asynchronous and overlapped, which is what this module provides.
"""
import numpy as np
import torch

from . import kernels as K


def resample_indices(lengths, num_tokens, rng=None):
    """[B, num_tokens] int64 row indices into the packed features of the batch (slide i's rows start at sum(lengths[:i])):
    per slide ``rng.choice(M_i, num_tokens, replace=M_i < num_tokens)`` -- datasets/dataset_pretrain.py:157-161."""
    rng = rng if rng is not None else np.random
    out = np.empty((len(lengths), num_tokens), dtype=np.int64)
    base = 0
    for i, m in enumerate(lengths):
        out[i] = base + rng.choice(int(m), num_tokens, replace=not int(m) >= num_tokens)
        base += int(m)
    return torch.from_numpy(out)


def gather_bags(packed, index, out=None):
    """packed: [sum M_i, Dw] device features (fp32 / bf16), index: [B, N] device int64 -> [B, N, Dw] fp32 (one launch)."""
    return K.gather_rows(packed, index, out)


class SlidePrefetcher:
    """Iterate device batches from a host iterable of ``(wsi, rna)`` pairs (``wsi`` either the ready ``[B, N, Dw]`` batch or a
    ``(packed [sum M_i, Dw], index [B, N])`` pair for the device-side gather) with the H2D copies of the NEXT batch overlapping
    the compute of the current one.  ``depth`` buffers (>= 2) are recycled; the consumer's stream is the current stream."""

    def __init__(self, loader, device, depth=2):
        self.loader, self.device, self.depth = loader, torch.device(device), max(2, depth)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.slots = [dict(pinned={}, dev={}, ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(self.depth)]
        self.h2d_bytes = 0

    def _stage(self, slot, name, t):
        """host tensor -> pinned buffer -> device buffer (async on the copy stream); buffers are re-used when the shape allows."""
        pin, dev = slot["pinned"].get(name), slot["dev"].get(name)
        if pin is None or pin.numel() < t.numel() or pin.dtype != t.dtype:
            pin = slot["pinned"][name] = torch.empty(t.numel(), dtype=t.dtype).pin_memory()
            dev = slot["dev"][name] = torch.empty(t.numel(), dtype=t.dtype, device=self.device)
        p = pin[:t.numel()].view(t.shape)
        if t.is_pinned():
            p = t  # already page-locked (e.g. a cached dataset): no staging copy
        else:
            p.copy_(t)
        d = dev[:t.numel()].view(t.shape)
        d.copy_(p, non_blocking=True)
        self.h2d_bytes += t.numel() * t.element_size()
        return d

    def _issue(self, k, item):
        slot = self.slots[k % self.depth]
        wsi, rna = item
        slot["ready"].synchronize()  # host: the previous H2D copy out of this slot's pinned buffers has completed
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(slot["free"])  # the step that used this slot's buffers has finished with them
            if isinstance(wsi, (tuple, list)):
                packed, index = wsi
                wsi_d = (self._stage(slot, "packed", packed), self._stage(slot, "index", index))
            else:
                wsi_d = self._stage(slot, "wsi", wsi)
            rna_d = self._stage(slot, "rna", rna)
            slot["ready"].record(self.copy_stream)
        return slot, wsi_d, rna_d

    def __iter__(self):
        it = iter(self.loader)
        for s in self.slots:
            s["free"].record(torch.cuda.current_stream(self.device))
        pending = []
        k = 0
        try:
            pending.append(self._issue(k, next(it)))
            k += 1
        except StopIteration:
            return
        while pending:
            try:
                pending.append(self._issue(k, next(it)))  # next batch travels while the current one is consumed
                k += 1
            except StopIteration:
                pass
            slot, wsi_d, rna_d = pending.pop(0)
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(slot["ready"])
            if isinstance(wsi_d, tuple):
                out = slot["dev"].get("gathered")
                shape = (*wsi_d[1].shape, wsi_d[0].shape[1])
                if out is None or tuple(out.shape) != shape:
                    out = slot["dev"]["gathered"] = torch.empty(shape, device=self.device, dtype=torch.float32)
                wsi_d = gather_bags(wsi_d[0], wsi_d[1], out)
            yield wsi_d, rna_d
            slot["free"].record(cur)  # everything the consumer enqueued on these buffers precedes this event


def length_bucketed_batches(lengths, batch_size, rng=None, drop_last=False):
    """Batches of slide indices for ``MIRROR.forward_varlen``: slides of EQUAL length are placed next to each other (the varlen path
    runs every group of equal length as one dense batch, so the number of distinct lengths per batch is the number of passes through
    the fixed-length kernels), batches are then shuffled.  ``lengths``: patches per slide (e.g. after rounding the bags down to a
    multiple of 256 patches, which keeps every patch but the remainder).  Returns a list of index lists."""
    rng = rng if rng is not None else np.random
    order = np.lexsort((rng.permutation(len(lengths)), np.asarray(lengths)))  # by length, ties in random order
    batches = [order[i:i + batch_size].tolist() for i in range(0, len(order), batch_size)]
    if drop_last and batches and len(batches[-1]) < batch_size:
        batches.pop()
    perm = rng.permutation(len(batches))
    return [batches[i] for i in perm]
