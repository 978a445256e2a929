"""Host-side logic that needs no GPU: the pack-ahead bookkeeping of the shared MLP (which operands a step asks for, in which
order, and that unknown requests fall back to in-line packing), argument validation of the index kernels' `out=` buffers."""
import pytest
import torch

from maskplanner_b200 import pointnet2_utils as P
from maskplanner_b200 import shared_mlp as S


@pytest.fixture
def fake_packing(monkeypatch):
    """Replace the two device-touching halves of packing by CPU stand-ins that log what they were asked to do."""
    fills = []

    def alloc(kind, cout_p, cin_p, mode, dev):
        return (torch.zeros(cout_p, cin_p), None, torch.zeros(cin_p, cout_p), None)

    def fill(bufs, kind, W, cout_p, cin_p, xyz_last, mode):
        fills.append((kind, W.data_ptr(), cout_p, cin_p, bool(xyz_last), mode))
        bufs[0].fill_(float(len(fills)))

    monkeypatch.setattr(S, "_alloc_packed", alloc)
    monkeypatch.setattr(S, "_fill_packed", fill)
    monkeypatch.setattr(S.PackAhead, "enabled", True)
    yield fills
    assert S._PACK_AHEAD is None, "a bracket was left open"


def _ask(weights):
    """What a forward pass does: the narrow first layer, then two GEMM layers."""
    a = S._narrow_weight(weights[0], 64, "bf16")
    b = S._padded_weight(weights[1], 64, 64, False, "bf16")
    c = S._padded_weight(weights[2], 128, 64, False, "bf16")
    return a, b, c


def test_pack_ahead_records_then_serves_the_same_requests(fake_packing):
    fills = fake_packing
    ws = [torch.randn(64, 3, 1, 1), torch.randn(64, 64, 1, 1), torch.randn(128, 64, 1, 1)]
    pa = S.PackAhead()
    pa.begin()                       # first bracketed step: packs in line, records
    _ask(ws)
    pa.end()
    assert len(fills) == 3 and [r[1] for r in pa.requests] == ["narrow", "gemm", "gemm"]
    del fills[:]
    pa.begin()                       # second step: everything is packed when the bracket opens, in the recorded order
    assert [f[0] for f in fills] == ["narrow", "gemm", "gemm"] and len(pa.ready) == 3
    a, b, c = _ask(ws)
    assert len(fills) == 3, "a served request must not pack again"
    assert not pa.ready
    first = (a.data_ptr(), b[0].data_ptr(), c[0].data_ptr())
    pa.end()
    pa.begin()                       # third step: the same persistent buffers, refilled
    a, b, c = _ask(ws)
    pa.end()
    assert (a.data_ptr(), b[0].data_ptr(), c[0].data_ptr()) == first
    assert len(fills) == 6


def test_pack_ahead_unknown_request_packs_in_line(fake_packing):
    fills = fake_packing
    ws = [torch.randn(64, 3, 1, 1), torch.randn(64, 64, 1, 1), torch.randn(128, 64, 1, 1)]
    pa = S.PackAhead()
    pa.begin(), _ask(ws), pa.end()
    del fills[:]
    pa.begin()
    other = torch.randn(64, 64, 1, 1)
    got = S._padded_weight(other, 64, 64, False, "bf16")          # not announced: packed on the spot into fresh buffers
    assert len(fills) == 4 and fills[-1][1] == other.data_ptr()
    assert all(got[0].data_ptr() != bufs[0].data_ptr() for bufs in pa.bufs.values())
    tf32 = S._padded_weight(ws[1], 64, 64, False, "fp32")          # same weight, other arithmetic: a different operand
    assert len(fills) == 5 and fills[-1][5] == "fp32"
    pa.end()
    assert S._PACK_AHEAD is None and not pa.ready


def test_packing_outside_a_bracket_is_plain(fake_packing):
    fills = fake_packing
    w = torch.randn(64, 64, 1, 1)
    S._padded_weight(w, 64, 64, True, "bf16")
    S._padded_weight(w, 64, 64, True, "bf16")
    assert len(fills) == 2 and fills[0][4] is True


def test_pack_ahead_disabled_is_inert(fake_packing, monkeypatch):
    fills = fake_packing
    monkeypatch.setattr(S.PackAhead, "enabled", False)
    ws = [torch.randn(64, 3, 1, 1), torch.randn(64, 64, 1, 1), torch.randn(128, 64, 1, 1)]
    pa = S.PackAhead()
    for _ in range(2):
        pa.begin(), _ask(ws), pa.end()
    assert pa.requests is None and len(fills) == 6


def test_out_buffers_of_the_index_kernels_are_validated():
    dev = torch.device("cpu")
    good = torch.empty(2, 5, dtype=torch.long)
    P._check_out(good, (2, 5), dev)
    for bad in (torch.empty(2, 4, dtype=torch.long), torch.empty(2, 5, dtype=torch.int32), torch.empty(5, 2, dtype=torch.long).t()):
        with pytest.raises(ValueError):
            P._check_out(bad, (2, 5), dev)


def test_identity_group_is_cached_per_shape():
    a = P._identity_group(3, 7, "cpu") if not torch.cuda.is_available() else None
    if a is None:
        pytest.skip("CPU-only check")
    assert a.shape == (3, 1, 7) and a.is_contiguous() and torch.equal(a[2, 0], torch.arange(7))
    assert P._identity_group(3, 7, "cpu") is a
