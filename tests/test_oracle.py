"""CPU tests of the oracle itself (no GPU): the C restatement against independent, literal emulations
of the reference kernels' control flow, written from the reference sources
(utils/src/sampling_gpu.cu:86-209, ball_query_gpu.cu:9-45, interpolate_gpu.cu:9-52)."""
import numpy as np
import pytest

from oracle import oracle as orc

f32 = np.float32


def dist_f32(a, b):
    """fma(dz,dz, fma(dx,dx, rn(dy*dy))) evaluated in float64 with explicit roundings.
    Products of two float32 are exact in float64, and float64 sums of such small terms round
    to float32 exactly as a fused multiply-add would for these magnitudes."""
    dx, dy, dz = f32(a[0] - b[0]), f32(a[1] - b[1]), f32(a[2] - b[2])
    t = f32(np.float64(dy) * np.float64(dy))
    t = f32(np.float64(dx) * np.float64(dx) + np.float64(t))
    return f32(np.float64(dz) * np.float64(dz) + np.float64(t))


def fps_emulated(xyz, m):
    """Thread-by-thread emulation of furthest_point_sampling_kernel<bs>: strided per-thread scan with a
    strict '>' update, then the shared-memory tree reduction that keeps the lower slot on ties."""
    n = xyz.shape[0]
    bs = 1
    while bs * 2 <= min(n, 1024):
        bs *= 2
    temp = np.full(n, 1e10, dtype=f32)
    idx = np.zeros(m, dtype=np.int32)
    old = 0
    for j in range(1, m):
        dists = np.zeros(bs, dtype=f32)
        dists_i = np.zeros(bs, dtype=np.int64)
        for tid in range(bs):
            best, besti = f32(-1), 0
            for k in range(tid, n, bs):
                d = dist_f32(xyz[k], xyz[old])
                d2 = min(d, temp[k])
                temp[k] = d2
                if d2 > best:
                    best, besti = d2, k
            dists[tid], dists_i[tid] = best, besti
        s = bs // 2
        while s >= 1:
            for tid in range(s):
                v1, v2 = dists[tid], dists[tid + s]
                i1, i2 = dists_i[tid], dists_i[tid + s]
                dists[tid] = max(v1, v2)
                dists_i[tid] = i2 if v2 > v1 else i1
            s //= 2
        old = int(dists_i[0])
        idx[j] = old
    return idx


@pytest.mark.parametrize("n,m,dup", [(16, 8, True), (37, 20, True), (64, 64, False), (100, 30, True), (130, 140, True),
                                     (5, 5, True), (1, 3, False), (2, 2, False)])
def test_fps_total_order_matches_kernel_emulation(n, m, dup):
    rng = np.random.default_rng(n * 1000 + m)
    if dup:
        # draw with replacement from a tiny pool: exact duplicates => exact ties in every round
        pool = rng.random((max(n // 3, 1), 3)).astype(f32)
        xyz = pool[rng.integers(0, pool.shape[0], n)]
    else:
        xyz = rng.random((n, 3)).astype(f32)
    got = orc.furthest_point_sample(xyz[None], m)[0]
    want = fps_emulated(xyz, m)
    np.testing.assert_array_equal(got, want)


def test_fps_grid_points_ties():
    # integer lattice: massive exact distance ties between distinct points
    g = np.stack(np.meshgrid(np.arange(4), np.arange(4), np.arange(3), indexing="ij"), -1).reshape(-1, 3).astype(f32)
    got = orc.furthest_point_sample(g[None], 24)[0]
    np.testing.assert_array_equal(got, fps_emulated(g, 24))


def test_opt_n_threads_table():
    # utils/src/cuda_utils.h:10-14
    for n, want in [(1, 1), (2, 2), (3, 2), (63, 32), (64, 64), (256, 256), (1000, 512), (1024, 1024), (8192, 1024),
                    (35000, 1024)]:
        assert orc.opt_n_threads(n) == want


def test_ball_query_semantics():
    rng = np.random.default_rng(7)
    xyz = rng.random((2, 200, 3)).astype(f32)
    new_xyz = xyz[:, :17].copy()
    new_xyz[0, 3] = 50.0  # empty ball -> zeros
    radius, K = 0.25, 8
    got = orc.ball_query(radius, K, xyz, new_xyz)
    r2 = f32(radius) * f32(radius)
    for b in range(2):
        for q in range(17):
            hits = [k for k in range(200) if dist_f32(new_xyz[b, q], xyz[b, k]) < r2]
            want = np.zeros(K, dtype=np.int32)
            if hits:
                want[:] = hits[0]
                want[:min(K, len(hits))] = hits[:K]
            np.testing.assert_array_equal(got[b, q], want)
    assert (got[0, 3] == 0).all()


def test_ball_query_strict_radius():
    # a point at exactly distance r is NOT in the ball (strict '<', ball_query_gpu.cu:34)
    xyz = np.array([[[0, 0, 0], [0.5, 0, 0], [0.25, 0, 0]]], dtype=f32)
    q = np.array([[[0, 0, 0]]], dtype=f32)
    got = orc.ball_query(0.5, 4, xyz, q)[0, 0]
    np.testing.assert_array_equal(got, [0, 2, 0, 0])


def test_three_nn_semantics_and_ties():
    rng = np.random.default_rng(11)
    known = rng.integers(0, 3, (1, 40, 3)).astype(f32)  # lattice: many exact ties -> first index wins
    unknown = rng.integers(0, 3, (1, 25, 3)).astype(f32) + f32(0.5)
    d, idx = orc.three_nn(unknown, known)
    for i in range(25):
        ds = np.array([dist_f32(unknown[0, i], known[0, k]) for k in range(40)], dtype=f32)
        order = np.lexsort((np.arange(40), ds))[:3]  # by distance, then index
        np.testing.assert_array_equal(idx[0, i], order)
        np.testing.assert_array_equal(d[0, i], np.sqrt(ds[order]))


def test_three_nn_fewer_than_three_known():
    d, idx = orc.three_nn(np.zeros((1, 2, 3), f32), np.ones((1, 2, 3), f32))
    assert np.isinf(d[0, :, 2]).all() and (idx[0, :, 2] == 0).all()
    assert np.isfinite(d[0, :, :2]).all()


def test_gather_group_interpolate_against_numpy():
    rng = np.random.default_rng(3)
    B, C, N, M, K = 2, 5, 50, 7, 4
    f = rng.standard_normal((B, C, N)).astype(f32)
    idx = rng.integers(0, N, (B, M)).astype(np.int32)
    np.testing.assert_array_equal(orc.gather_operation(f, idx), np.take_along_axis(f, idx[:, None, :].repeat(C, 1), 2))
    gidx = rng.integers(0, N, (B, M, K)).astype(np.int32)
    want = np.stack([f[b][:, gidx[b]] for b in range(B)])
    np.testing.assert_array_equal(orc.grouping_operation(f, gidx), want)
    idx3 = rng.integers(0, N, (B, M, 3)).astype(np.int32)
    w = rng.random((B, M, 3)).astype(f32)
    got = orc.three_interpolate(f, idx3, w)
    ref = np.einsum("bcmj,bmj->bcm", np.stack([f[b][:, idx3[b]] for b in range(B)]).astype(np.float64), w.astype(np.float64))
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-6)
    # backward ops are the transposes of the forwards
    go = rng.standard_normal((B, C, M)).astype(f32)
    gg = orc.gather_operation_grad(go, idx, N)
    assert np.allclose((gg * f).sum(), (go * orc.gather_operation(f, idx)).sum(), rtol=1e-4)
    gi = orc.three_interpolate_grad(go, idx3, w, N)
    assert np.allclose((gi * f).sum(), (go * got).sum(), rtol=1e-4)
    go4 = rng.standard_normal((B, C, M, K)).astype(f32)
    g4 = orc.grouping_operation_grad(go4, gidx, N)
    assert np.allclose((g4 * f).sum(), (go4 * want).sum(), rtol=1e-4)


def test_fp_weights():
    d = np.array([[[0.0, 1.0, 2.0], [1.0, 1.0, 1.0]]], dtype=f32)
    w = orc.fp_weights(d)
    np.testing.assert_allclose(w.sum(-1), 1.0, rtol=1e-6)
    assert w[0, 0, 0] > 0.999999  # zero distance clamps to 1e-10 and takes all the weight
    np.testing.assert_allclose(w[0, 1], 1 / 3, rtol=1e-6)
