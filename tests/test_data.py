"""Input pipeline (mirror_b200/data.py): the dataset's resampling semantics, the device-side gather, and the
prefetch ring (data integrity across buffer re-use)."""
import numpy as np
import pytest
import torch


def test_resample_indices_follow_dataset_semantics():
    from mirror_b200.data import resample_indices
    rng = np.random.RandomState(0)
    lengths = [50, 7, 20]
    idx = resample_indices(lengths, 20, rng).numpy()
    assert idx.shape == (3, 20) and idx.dtype == np.int64
    base = np.cumsum([0] + lengths)
    for i, m in enumerate(lengths):
        loc = idx[i] - base[i]
        assert loc.min() >= 0 and loc.max() < m
        if m >= 20:   # without replacement (datasets/dataset_pretrain.py:157-160: replace = not M >= N)
            assert len(set(loc.tolist())) == 20
    # the same generator state gives the reference's own draw
    rng2 = np.random.RandomState(0)
    want0 = rng2.choice(50, 20, replace=False)
    assert np.array_equal(idx[0], want0)


def test_gather_bags_emulated():
    import emu_backend
    from mirror_b200.data import gather_bags
    emu_backend.use()
    try:
        src = torch.randn(30, 12)
        idx = torch.randint(0, 30, (2, 5))
        assert torch.equal(gather_bags(src, idx), src[idx])
    finally:
        emu_backend.release()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,cols", [(torch.float32, 768), (torch.bfloat16, 768), (torch.float32, 37), (torch.bfloat16, 1024), (torch.bfloat16, 13)])
def test_gather_rows_kernel(dtype, cols):
    from mirror_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(1)
    src = torch.randn(1000, cols, device="cuda", generator=g).to(dtype)
    idx = torch.randint(0, 1000, (4, 300), device="cuda", generator=g)
    out = K.gather_rows(src, idx)
    assert out.dtype == torch.float32 and torch.equal(out, src[idx].float())


@pytest.mark.gpu
def test_prefetcher_overlaps_copies_and_keeps_batches_intact():
    from mirror_b200.data import SlidePrefetcher, resample_indices
    g = torch.Generator().manual_seed(2)
    rng = np.random.RandomState(3)
    batches, want = [], []
    for i in range(7):
        if i % 2 == 0:  # ready-made [B, N, Dw] batch
            wsi, rna = torch.randn(3, 64, 96, generator=g), torch.randn(3, 30, generator=g)
            batches.append((wsi, rna))
            want.append((wsi.clone(), rna.clone()))
        else:           # packed bf16 features + sampled indices -> device-side gather
            lengths = [40 + i, 100, 64]
            packed = torch.randn(sum(lengths), 96, generator=g).to(torch.bfloat16)
            index = resample_indices(lengths, 64, rng)
            rna = torch.randn(3, 30, generator=g)
            batches.append(((packed, index), rna))
            want.append((packed[index].float(), rna.clone()))
    pf = SlidePrefetcher(batches, "cuda", depth=2)
    got = []
    for wsi_d, rna_d in pf:
        assert wsi_d.is_cuda and wsi_d.dtype == torch.float32
        y = wsi_d * 2.0  # consumer work on the current stream
        got.append((y.cpu() / 2.0, rna_d.cpu()))
    assert len(got) == 7 and pf.h2d_bytes > 0
    for (a, b), (wa, wb) in zip(got, want):
        assert torch.equal(a, wa) and torch.equal(b, wb)


def test_length_bucketed_batches_minimise_distinct_lengths():
    from mirror_b200.data import length_bucketed_batches
    rng = np.random.RandomState(0)
    lengths = (rng.randint(4, 12, size=203) * 256).tolist()
    batches = length_bucketed_batches(lengths, 16, np.random.RandomState(1))
    flat = sorted(i for b in batches for i in b)
    assert flat == list(range(203))                                    # a partition of the dataset
    distinct = [len({lengths[i] for i in b}) for b in batches]
    assert max(distinct) <= 3 and np.mean(distinct) < 2.0              # vs ~6.5 distinct lengths in a random batch of 16
    naive = [len({lengths[i] for i in rng.permutation(203)[:16]}) for _ in range(20)]
    assert np.mean(naive) > 2 * np.mean(distinct)
    assert all(len(b) == 16 for b in length_bucketed_batches(lengths, 16, np.random.RandomState(2), drop_last=True))
