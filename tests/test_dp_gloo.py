"""N>1 host logic on CPU (gloo, world_size 2): batch sharding, flat gradient buckets and the mean
all-reduce reproduce the single-process full-batch gradient (for a model without batch statistics)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maskplanner_b200.train_step import FlatGradBuckets, all_reduce_mean_, shard_range


def _model():
    torch.manual_seed(0)
    m = torch.nn.Sequential()
    m.add_module("sa1", torch.nn.Linear(6, 16))
    m.add_module("act", torch.nn.Tanh())
    m.add_module("fc1", torch.nn.Linear(16, 4))
    return m


def _data():
    g = torch.Generator().manual_seed(1)
    return torch.randn(10, 6, generator=g), torch.randn(10, 4, generator=g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _model()
    b = FlatGradBuckets(m)
    x, y = _data()
    lo, hi = shard_range(10, rank, world)
    b.zero()
    # per-rank mean loss; ranks own equal-sized shards here so the mean of means is the global mean
    ((m(x[lo:hi]) - y[lo:hi]) ** 2).mean().backward()
    all_reduce_mean_(b.heads, world)
    all_reduce_mean_(b.encoder, world)
    if rank == 0:
        torch.save(b.flat.clone(), out)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_the_batch():
    for gb, w in [(64, 8), (64, 1), (10, 4), (7, 8)]:
        r = [shard_range(gb, i, w) for i in range(w)]
        assert r[0][0] == 0 and r[-1][1] == gb
        assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_flat_buckets_are_views_of_param_grads():
    m = _model()
    b = FlatGradBuckets(m)
    assert b.flat.numel() == sum(p.numel() for p in m.parameters())
    assert b.heads.numel() == 16 * 4 + 4 and b.encoder.numel() == 6 * 16 + 16
    x, y = _data()
    ((m(x) - y) ** 2).mean().backward()
    for p in m.parameters():
        assert p.grad.data_ptr() >= b.flat.data_ptr() and p.grad.abs().sum() > 0
    b.zero()
    assert all(float(p.grad.abs().sum()) == 0 for p in m.parameters())


def test_two_rank_gloo_allreduce_equals_full_batch_gradient(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    m = _model()
    b = FlatGradBuckets(m)
    x, y = _data()
    ((m(x) - y) ** 2).mean().backward()
    assert torch.allclose(got, b.flat, rtol=1e-5, atol=1e-7)


def test_flat_bucket_views_are_16_byte_aligned_and_disjoint():
    """Every .grad view starts on a 16-byte boundary of the flat buffer (vector path of the one-launch Adam kernel),
    views do not overlap, heads come first, and the padding between views stays zero through backward."""
    torch.manual_seed(3)
    m = torch.nn.Sequential()
    m.add_module("sa1", torch.nn.Linear(5, 7))          # 35 + 7 elements: not multiples of 4
    m.add_module("act", torch.nn.Tanh())
    m.add_module("fc1", torch.nn.Linear(7, 3))          # 21 + 3
    b = FlatGradBuckets(m)
    spans = []
    for p in b.params:
        off = (p.grad.data_ptr() - b.flat.data_ptr()) // 4
        assert (p.grad.data_ptr() - b.flat.data_ptr()) % 16 == 0
        assert p.grad.shape == p.shape
        spans.append((off, off + p.numel()))
    spans.sort()
    assert all(a_end <= b_start for (_, a_end), (b_start, _) in zip(spans, spans[1:]))
    n_heads = sum((p.numel() + 3) // 4 * 4 for p in b.head_params)
    assert b.heads.numel() == n_heads and b.heads.numel() + b.encoder.numel() == b.flat.numel()
    assert all(any(p is q for q in b.head_params) for p in [m.fc1.weight, m.fc1.bias])
    b.zero()
    (m(torch.randn(4, 5)) ** 2).sum().backward()
    used = torch.zeros_like(b.flat, dtype=torch.bool)
    for lo, hi in spans:
        used[lo:hi] = True
    assert float(b.flat[~used].abs().max()) == 0.0 and float(b.flat[used].abs().max()) > 0.0


def test_narrow_first_layer_eligibility(monkeypatch):
    """Host-side switch of the on-the-fly first SA layer: <= 8 channels per grouped row, no gradient into the point
    features, and the MPB_NARROW_FIRST escape hatch."""
    from maskplanner_b200 import shared_mlp as SM
    monkeypatch.delenv("MPB_NARROW_FIRST", raising=False)
    assert SM.narrow_rows_supported(None, 32)
    assert SM.narrow_rows_supported(torch.zeros(2, 10, 3), 32)
    assert SM.narrow_rows_supported(torch.zeros(2, 10, 5), 32)
    assert not SM.narrow_rows_supported(torch.zeros(2, 10, 6), 32)              # 3 + 6 > 8 channels
    assert not SM.narrow_rows_supported(torch.zeros(2, 10, 3, requires_grad=True), 32)
    monkeypatch.setenv("MPB_NARROW_FIRST", "0")
    assert not SM.narrow_rows_supported(None, 32)
    monkeypatch.delenv("MPB_NARROW_FIRST")
    assert SM.narrow_rows_supported(None, 32)
