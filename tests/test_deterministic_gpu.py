"""Deterministic backwards: inverse index + segmented reduction instead of the reference's atomicAdd
scatter (sampling_gpu.cu:46-83, group_points_gpu.cu:47-83, interpolate_gpu.cu:120-161).  Same gradients within fp32
summation order, bit-identical from run to run."""
import numpy as np
import pytest
import torch

from pn2_b200 import pointnet2_utils as pu

pytestmark = pytest.mark.gpu

REL = 1e-5  # fp32 sums of <= a few thousand terms, compared with an fp64 accumulation


@pytest.fixture
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


@pytest.fixture
def deterministic():
    prev = pu.set_deterministic(True)
    yield
    pu.set_deterministic(prev)


def ref_scatter(grad_out, idx, n, weight=None):
    """fp64 accumulation on the host: grad[b, c, idx[b, p]] += w[b, p] * grad_out[b, c, p // div]"""
    B, C = grad_out.shape[:2]
    g = grad_out.reshape(B, C, -1).astype(np.float64)
    flat = idx.reshape(B, -1)
    out = np.zeros((B, C, n), dtype=np.float64)
    for b in range(B):
        if weight is None:
            src = g[b]
        else:
            w = weight.reshape(B, -1)[b].astype(np.float32)
            src = (g[b][:, np.arange(flat.shape[1]) // 3].astype(np.float32) * w[None, :]).astype(np.float64)  # rn(go * w) terms
        for c in range(C):
            np.add.at(out[b, c], flat[b], src[c])
    return out


def close(got, want):
    scale = max(np.abs(want).max(), 1e-30)
    assert np.abs(got.astype(np.float64) - want).max() <= REL * scale


@pytest.mark.parametrize("B,C,N,M,K,hot", [(2, 19, 700, 128, 16, False), (1, 64, 4096, 512, 32, False), (2, 8, 50, 64, 32, True),
                                           (3, 3, 1, 10, 4, True)])
def test_group_points_grad_deterministic(cuda, deterministic, B, C, N, M, K, hot):
    g = torch.Generator().manual_seed(B * 1000 + C)
    idx = torch.randint(0, max(1, N // (8 if hot else 1)), (B, M, K), generator=g, dtype=torch.int32)  # hot: many collisions
    feat = torch.randn(B, C, N, generator=g)
    grad_out = torch.randn(B, C, M, K, generator=g)
    outs = []
    for _ in range(3):
        f = feat.to(cuda).requires_grad_(True)
        pu.grouping_operation(f, idx.to(cuda)).backward(grad_out.to(cuda))
        outs.append(f.grad.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])  # bit-identical run to run
    close(outs[0].cpu().numpy(), ref_scatter(grad_out.numpy(), idx.numpy(), N))
    pu.set_deterministic(False)  # the atomics path computes the same sums in another order
    f = feat.to(cuda).requires_grad_(True)
    pu.grouping_operation(f, idx.to(cuda)).backward(grad_out.to(cuda))
    close(f.grad.cpu().numpy(), ref_scatter(grad_out.numpy(), idx.numpy(), N))


@pytest.mark.parametrize("B,C,N,M", [(2, 3, 8192, 1024), (1, 33, 100, 400), (2, 5, 1, 7)])
def test_gather_points_grad_deterministic(cuda, deterministic, B, C, N, M):
    g = torch.Generator().manual_seed(N + M)
    idx = torch.randint(0, N, (B, M), generator=g, dtype=torch.int32)
    feat = torch.randn(B, C, N, generator=g)
    grad_out = torch.randn(B, C, M, generator=g)
    outs = []
    for _ in range(2):
        f = feat.to(cuda).requires_grad_(True)
        pu.gather_operation(f, idx.to(cuda)).backward(grad_out.to(cuda))
        outs.append(f.grad.clone())
    assert torch.equal(outs[0], outs[1])
    close(outs[0].cpu().numpy(), ref_scatter(grad_out.numpy(), idx.numpy(), N))


@pytest.mark.parametrize("B,C,n,m", [(2, 128, 8192, 1024), (1, 7, 300, 3), (2, 16, 64, 1)])
def test_three_interpolate_grad_deterministic(cuda, deterministic, B, C, n, m):
    g = torch.Generator().manual_seed(n + m)
    idx = torch.randint(0, m, (B, n, 3), generator=g, dtype=torch.int32)
    w = torch.rand(B, n, 3, generator=g)
    w = w / w.sum(-1, keepdim=True)
    feat = torch.randn(B, C, m, generator=g)
    grad_out = torch.randn(B, C, n, generator=g)
    outs = []
    for _ in range(3):
        f = feat.to(cuda).requires_grad_(True)
        pu.three_interpolate(f, idx.to(cuda), w.to(cuda)).backward(grad_out.to(cuda))
        outs.append(f.grad.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    close(outs[0].cpu().numpy(), ref_scatter(grad_out.numpy(), idx.numpy(), m, weight=w.numpy()))


def test_follows_torch_deterministic_switch(cuda):
    assert pu.set_deterministic(None) is None
    assert not pu._deterministic()
    torch.use_deterministic_algorithms(True, warn_only=True)
    try:
        assert pu._deterministic()
    finally:
        torch.use_deterministic_algorithms(False)
    assert pu.set_deterministic(True) is None and pu._deterministic()
    pu.set_deterministic(None)


def test_inverse_index_layout(cuda):
    """seg_start / pos as documented in include/pn2_abi.h: segment k of cloud b lists its source positions ascending."""
    from pn2_b200 import _lib
    B, n, J = 3, 11, 40
    idx = torch.randint(0, n, (B, J), generator=torch.Generator().manual_seed(5), dtype=torch.int32)
    d = idx.to(cuda)
    seg = torch.empty(B * n + 1, dtype=torch.int32, device=cuda)
    pos = torch.empty(B, J, dtype=torch.int32, device=cuda)
    _lib.call("pn2_inverse_index", B, n, J, _lib.ptr(d), _lib.ptr(seg), _lib.ptr(pos), _lib.stream_ptr(cuda))
    seg, pos = seg.cpu().numpy(), pos.cpu().numpy().reshape(-1)
    assert seg[0] == 0 and seg[-1] == B * J and (np.diff(seg) >= 0).all()
    for b in range(B):
        for k in range(n):
            want = np.nonzero(idx[b].numpy() == k)[0]
            np.testing.assert_array_equal(pos[seg[b * n + k]:seg[b * n + k + 1]], want)
