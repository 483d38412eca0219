"""GPU parity of the multi-view lifting (rows A12-A14) and the loader's frustum count (row N2).

Anchors, strongest first:
  * tests/golden/projection_r2.npz -- vectors produced by the REFERENCE's own utils/projection.py and the view reductions of
    model/pointnet2multiview.py run on the host (tests/golden/make_golden_projection.py).  The kernels are fed with the
    reference's per-view parameters and must reproduce its decisions (identical except listed rounding-boundary cases)
    and its lifted features bit for bit; the product's own parameter kernel is pinned to the reference's values separately.
  * the C oracle (pinned to the same vectors by tests/test_golden.py) on further shapes, bit-exact.
  * the torch restatement oracle/projection_ref.py (also pinned there) for the end-to-end agreement rate."""
import numpy as np
import pytest
import torch

from lifting_cases import PROJ, boundary_cases, mgj, ref_view_params
from oracle import oracle as orc
from oracle import projection_ref
from pn2_b200 import projection, scenes

pytestmark = pytest.mark.gpu

INTR = torch.from_numpy(scenes.SCANNET_INTRINSIC)
DMIN, DMAX = scenes.SCANNET_DEPTH_RANGE
ACC = scenes.SCANNET_ACCURACY
DIMS = scenes.SCANNET_IMAGE_DIMS


def make_batch(B, N, V, C, seed):
    xyz, feats, depth, poses = [], [], [], []
    for b in range(B):
        x, _ = scenes.scannet_scene(seed + b, N)
        f, d, p = scenes.multiview_inputs(seed + b, x, V, C)
        xyz.append(x); feats.append(f); depth.append(d); poses.append(p)
    return np.stack(xyz), np.stack(feats), np.stack(depth), np.stack(poses)


@pytest.mark.parametrize("reduce", ["max", "first"])
@pytest.mark.parametrize("B,N,V,C", [(2, 8192, 3, 128), (1, 3000, 5, 16)])
def test_lift_views_bit_exact_vs_oracle(cuda, reduce, B, N, V, C):
    xyz, feats, depth, poses = make_batch(B, N, V, C, 40)
    if reduce == "first":
        feats[:, 0, :, 5, :] = 0.0  # visible but all-zero columns in view 0 must be replaced by later views
    t = lambda a: torch.from_numpy(a).to(cuda)
    out, pix, count = projection.lift_views(t(xyz), t(feats), t(depth), t(poses), INTR, DMIN, DMAX, DIMS, ACC, reduce=reduce,
                                            return_pixels=True)
    # the per-view parameters come from the product's own setup kernel (pinned to the reference's torch formulas by
    # test_view_params_match_the_reference_formulas below); the oracle then checks projection, tests and gather bit for bit
    w2c, corner2, corner4, normals = projection.view_params(t(poses), INTR, DMIN, DMAX, DIMS)
    want, wpix = orc.lift_views(xyz, feats, depth, w2c.cpu().numpy().reshape(B, V, 16), corner2.cpu().numpy(),
                                corner4.cpu().numpy(), normals.cpu().numpy().reshape(B, V, 18),
                                np.array([INTR[0, 0], INTR[1, 1], INTR[0, 2], INTR[1, 2]], np.float32), DMIN, DMAX, ACC, reduce)
    np.testing.assert_array_equal(pix.cpu().numpy(), wpix)
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    np.testing.assert_array_equal(count.cpu().numpy(), (wpix >= 0).sum(-1))
    assert (wpix >= 0).mean() > 0.05, "the synthetic views must actually see the scene"


def test_view_params_match_the_reference_formulas(cuda):
    """pn2_lift_setup against the reference's op sequence (utils/projection.py:25-95, :178): torch.inverse, pose x corners,
    cross products.  One launch instead of ~25; agreement to a few ulp (the inverse is an fp64 adjugate rounded once)."""
    _, _, _, poses = make_batch(3, 64, 5, 4, 77)
    c2w = torch.from_numpy(poses).to(cuda)
    w2c, c2, c4, nrm = projection.view_params(c2w, INTR, DMIN, DMAX, DIMS)
    corners = projection.frustum_corners(INTR, DMIN, DMAX, DIMS, c2w)
    want_n = projection.frustum_normals(corners)
    want_w = torch.inverse(c2w.double())  # the exact inverse; torch.inverse in fp32 is itself only good to ~1e-6

    def close(a, b, rel):
        a, b = a.double().cpu(), b.double().cpu()
        assert (a - b).abs().max() <= rel * b.abs().max()

    close(w2c, want_w, 2e-7)
    close(torch.inverse(c2w), want_w, 2e-5)          # the reference's own fp32 inverse is further from the truth than ours
    close(c2, corners[..., 2, :3], 2e-7)
    close(c4, corners[..., 4, :3], 2e-7)
    close(nrm, want_n, 2e-6)
    assert w2c.shape == (3, 5, 4, 4) and nrm.shape == (3, 5, 6, 3)
    # the bottom row of an inverted rigid pose
    assert torch.allclose(w2c[..., 3, :].cpu(), torch.tensor([0.0, 0.0, 0.0, 1.0]).expand(3, 5, 4), atol=1e-7)


def test_lift_matches_reference_op_sequence(cuda):
    """Against the torch restatement of utils/projection.py (BLAS-ordered mm): pixels may differ only where a
    value sits on a rounding boundary.  Required: >= 99.9 % identical decisions, and identical features wherever
    the decisions agree."""
    B, N, V, C = 2, 8192, 3, 32
    xyz, feats, depth, poses = make_batch(B, N, V, C, 60)
    t = lambda a: torch.from_numpy(a).to(cuda)
    for reduce in ("max", "first"):
        out, pix, _ = projection.lift_views(t(xyz), t(feats), t(depth), t(poses), INTR, DMIN, DMAX, DIMS, ACC, reduce=reduce,
                                            return_pixels=True)
        pix = pix.cpu()
        agree = 0
        for b in range(B):
            ref_pix = [projection_ref.compute_projection_ref(torch.from_numpy(xyz[b]), torch.from_numpy(depth[b, v]),
                                                             torch.from_numpy(poses[b, v]), INTR, DMIN, DMAX, DIMS, ACC)
                       for v in range(V)]
            same = torch.stack([ref_pix[v] == pix[b, v].long() for v in range(V)]).all(0)
            agree += int(same.sum())
            ref_out = projection_ref.lift_ref(torch.from_numpy(feats[b]), ref_pix, reduce)
            np.testing.assert_array_equal(out[b].cpu().numpy()[:, same.numpy()], ref_out.numpy()[:, same.numpy()])
        assert agree / (B * N) >= 0.999, "index agreement %.5f" % (agree / (B * N))


def test_projection_helper_surface(cuda):
    xyz, feats, depth, poses = make_batch(1, 4096, 2, 8, 80)
    helper = projection.ProjectionHelper(INTR, DMIN, DMAX, DIMS, ACC)
    pts = torch.from_numpy(xyz[0]).to(cuda)
    res = helper.compute_projection(pts, torch.from_numpy(depth[0, 0]).to(cuda), torch.from_numpy(poses[0, 0]).to(cuda), 4096)
    assert res is not None
    ind3d, ind2d = res
    assert ind3d.dtype == torch.int64 and ind3d.shape == (4097,) and int(ind3d[0]) == int(ind2d[0]) > 0
    n = int(ind3d[0])
    ref = projection_ref.compute_projection_ref(torch.from_numpy(xyz[0]), torch.from_numpy(depth[0, 0]), torch.from_numpy(poses[0, 0]),
                                                INTR, DMIN, DMAX, DIMS, ACC)
    ref3d = torch.nonzero(ref >= 0).squeeze(1)
    inter = np.intersect1d(ind3d[1:1 + n].cpu().numpy(), ref3d.numpy())
    assert len(inter) >= 0.995 * max(n, len(ref3d))
    assert (ind3d[1:1 + n][1:] > ind3d[1:1 + n][:-1]).all()  # ascending point order, as the reference packs them
    # Projection.apply on the packed vectors reproduces the single-view map
    f = torch.from_numpy(feats[0, 0]).to(cuda)
    single = projection.Projection.apply(f, ind3d, ind2d, 4096)
    fused = projection.lift_views(pts[None], f[None, None], torch.from_numpy(depth[0, :1]).to(cuda)[None],
                                  torch.from_numpy(poses[0, :1]).to(cuda)[None], INTR, DMIN, DMAX, DIMS, ACC)
    np.testing.assert_array_equal(single.cpu().numpy(), fused[0].cpu().numpy())
    # a camera looking away sees nothing -> None (the reference skips the whole batch then)
    away = poses[0, 0].copy()
    away[:3, 2] *= -1
    away[:3, 0] *= -1
    assert helper.compute_projection(pts, torch.from_numpy(depth[0, 0]).to(cuda), torch.from_numpy(away).to(cuda), 4096) is None
    # frustum helpers keep the reference's shapes
    cc = helper.compute_frustum_corners(torch.from_numpy(poses[0, 0]))
    assert cc.shape == (8, 4, 1) and helper.compute_frustum_normals(cc).shape == (6, 3)
    cnt = helper.points_in_frustum_cpu(cc, helper.compute_frustum_normals(cc), torch.from_numpy(xyz[0]))
    assert int(cnt) > 0


def _lift_with_reference_params(cuda, name, reduce, gold):
    _, B, N, V, C = mgj.CASES[name]
    xyz, feats, depth, poses = mgj.lifting_inputs(name)
    w2c, c2, c4, nrm = ref_view_params(gold, name)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    out, pix, count = projection.lift_views(t(xyz), t(feats), t(depth), None, mgj.INTRINSIC, mgj.DEPTH_MIN, mgj.DEPTH_MAX,
                                            mgj.IMAGE_DIMS, mgj.ACCURACY, reduce=reduce, return_pixels=True,
                                            view_parameters=(t(w2c), t(c2), t(c4), t(nrm)))
    return (xyz, feats, depth, poses), (w2c, c2, c4, nrm), out.cpu().numpy(), pix.cpu().numpy(), count.cpu().numpy()


@pytest.mark.parametrize("reduce", ["max", "first"])
@pytest.mark.parametrize("name", list(mgj.CASES))
def test_lift_views_reproduces_the_reference(cuda, name, reduce):
    """pn2_lift_views fed with the reference's own per-view parameters (torch.inverse, compute_frustum_corners / _normals
    as stored in the fixture) against the reference's compute_projection decisions, Projection.forward maps and view
    reductions (utils/projection.py:166-256, model/pointnet2multiview.py:36-41 / 89-99)."""
    gold = np.load(PROJ)
    (xyz, feats, depth, _), (w2c, c2, c4, nrm), out, pix, count = _lift_with_reference_params(cuda, name, reduce, gold)
    want = gold[name + "/pix"]
    diff = np.argwhere(pix != want)
    assert len(diff) <= 1e-3 * pix.size, "decision agreement %.5f" % (1 - len(diff) / pix.size)
    assert boundary_cases(xyz, depth, w2c, c2, c4, nrm, diff).all(), "a decision differs away from any rounding boundary"
    np.testing.assert_array_equal(count, (pix >= 0).sum(-1))
    B = xyz.shape[0]
    for b in range(B):
        same = (pix[b] == want[b]).all(0)
        np.testing.assert_array_equal(out[b][:, :mgj.KEEP_POINTS][:, same[:mgj.KEEP_POINTS]],
                                      gold["%s/%s/head/%d" % (name, reduce, b)][:, same[:mgj.KEEP_POINTS]])
        if same.all():
            assert mgj.sha(out[b]) == str(gold["%s/%s/sha/%d" % (name, reduce, b)]), "lifted features differ from the reference's"
        else:
            np.testing.assert_allclose(out[b].astype(np.float64).sum(1), gold["%s/%s/chansum/%d" % (name, reduce, b)], atol=50.0)


@pytest.mark.parametrize("name", list(mgj.CASES))
def test_view_params_reproduce_the_reference(cuda, name):
    """pn2_lift_setup (one launch) against the reference's torch.inverse / compute_frustum_corners / compute_frustum_normals
    values: corners to 1 ulp-class error, normals to the cancellation the cross products allow, the inverse within the error
    of the reference's own fp32 LU."""
    gold = np.load(PROJ)
    _, _, _, poses = mgj.lifting_inputs(name)
    w2c, c2, c4, nrm = projection.view_params(torch.from_numpy(poses).to(cuda), mgj.INTRINSIC, mgj.DEPTH_MIN, mgj.DEPTH_MAX, mgj.IMAGE_DIMS)
    rw, rc2, rc4, rn = ref_view_params(gold, name)

    def close(a, b, rel):
        a, b = a.double().cpu().numpy().reshape(b.shape), b.astype(np.float64)
        assert np.abs(a - b).max() <= rel * np.abs(b).max(), np.abs(a - b).max() / np.abs(b).max()

    close(c2, rc2, 2.5e-7)
    close(c4, rc4, 2.5e-7)
    close(nrm, rn, 2e-6)
    close(w2c, rw, 2e-5)


@pytest.mark.parametrize("name", list(mgj.CASES))
def test_lift_views_end_to_end_agreement_with_the_reference(cuda, name):
    """The whole product path (own parameter kernel + lifting) against the reference's decisions: the parameters differ
    from the reference's in the last place, so a point may flip only on a rounding boundary."""
    gold = np.load(PROJ)
    _, B, N, V, C = mgj.CASES[name]
    xyz, feats, depth, poses = mgj.lifting_inputs(name)
    t = lambda a: torch.from_numpy(a).to(cuda)
    w2c, c2, c4, nrm = ref_view_params(gold, name)
    for reduce in ("max", "first"):
        out, pix, _ = projection.lift_views(t(xyz), t(feats), t(depth), t(poses), mgj.INTRINSIC, mgj.DEPTH_MIN, mgj.DEPTH_MAX,
                                            mgj.IMAGE_DIMS, mgj.ACCURACY, reduce=reduce, return_pixels=True)
        pix, out = pix.cpu().numpy(), out.cpu().numpy()
        want = gold[name + "/pix"]
        diff = np.argwhere(pix != want)
        assert len(diff) <= 1e-3 * pix.size
        assert boundary_cases(xyz, depth, w2c, c2, c4, nrm, diff).all()
        for b in range(B):
            same = (pix[b] == want[b]).all(0)[:mgj.KEEP_POINTS]
            np.testing.assert_array_equal(out[b][:, :mgj.KEEP_POINTS][:, same], gold["%s/%s/head/%d" % (name, reduce, b)][:, same])


def test_compute_projection_and_projection_apply_reproduce_the_reference(cuda):
    """The reference's call surface: ProjectionHelper.compute_projection -> ([count, idx...], [count, pix...]) int64 vectors
    and Projection.apply on them (utils/projection.py:166-256), against the fixture."""
    gold = np.load(PROJ)
    name = "v3_n8192"
    _, B, N, V, C = mgj.CASES[name]
    xyz, feats, depth, poses = mgj.lifting_inputs(name)
    helper = projection.ProjectionHelper(mgj.INTRINSIC, mgj.DEPTH_MIN, mgj.DEPTH_MAX, mgj.IMAGE_DIMS, mgj.ACCURACY)
    w2c, c2, c4, nrm = ref_view_params(gold, name)
    b, v = 0, V - 1
    pts = torch.from_numpy(xyz[b]).to(cuda)
    ind3d, ind2d = helper.compute_projection(pts, torch.from_numpy(depth[b, v]).to(cuda), torch.from_numpy(poses[b, v]).to(cuda), N)
    n = int(ind3d[0])
    got = np.full(N, -1, np.int64)
    got[ind3d[1:1 + n].cpu().numpy()] = ind2d[1:1 + n].cpu().numpy()
    want = gold[name + "/pix"][b, v]
    diff = np.argwhere(got != want)
    diff = np.concatenate([np.full((len(diff), 1), b), np.full((len(diff), 1), v), diff], 1)
    assert len(diff) <= 1e-3 * N and boundary_cases(xyz, depth, w2c, c2, c4, nrm, diff).all()
    # Projection.apply on the REFERENCE's index vectors -> the reference's single-view map, bit for bit
    keep = np.nonzero(want >= 0)[0]
    r3d = torch.zeros(N + 1, dtype=torch.int64)
    r2d = torch.zeros(N + 1, dtype=torch.int64)
    r3d[0] = r2d[0] = len(keep)
    r3d[1:1 + len(keep)] = torch.from_numpy(keep)
    r2d[1:1 + len(keep)] = torch.from_numpy(want[keep].astype(np.int64))
    single = projection.Projection.apply(torch.from_numpy(feats[b, v]).to(cuda), r3d.to(cuda), r2d.to(cuda), N)
    assert mgj.sha(single.cpu().numpy()) == str(gold["%s/single/sha/%d" % (name, b)])


def test_frustum_counts_and_best_views(cuda):
    """Row N2: the loader's best-view selection (data_utils/ScanNetDataLoader.py:87-105), all poses of a scene
    in one launch, fp64 as the reference.  Expected values: the reference's own points_in_frustum_cpu on the host
    (fixture).  With the reference's corners / normals the counts must be identical; with the product's own parameter
    kernel a count may differ only by points that sit on a rounding boundary of a plane test."""
    gold = np.load(PROJ)
    x, poses = mgj.frustum_inputs()
    want = gold["frustum/counts"]
    cor, nrm = gold["frustum/corners"], gold["frustum/normals"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    exact = projection.frustum_counts(t(x), None, mgj.INTRINSIC, mgj.DEPTH_MIN, mgj.DEPTH_MAX, mgj.IMAGE_DIMS,
                                      view_parameters=(t(cor[:, 2]), t(cor[:, 4]), t(nrm))).cpu().numpy()
    np.testing.assert_array_equal(exact, want)
    got = projection.frustum_counts(t(x), t(poses), mgj.INTRINSIC, mgj.DEPTH_MIN, mgj.DEPTH_MAX, mgj.IMAGE_DIMS).cpu().numpy()
    # points within fp32 noise of a plane's -0.005 threshold, per pose (fp64 evaluation of the reference's parameters)
    p = x.astype(np.float64)
    near = np.zeros(len(want), np.int64)
    for q in range(len(want)):
        close = np.zeros(len(p), bool)
        for k in range(6):
            s = 100.0 * ((p - cor[q, 2 if k < 3 else 4].astype(np.float64)) @ nrm[q, k].astype(np.float64))
            close |= np.abs(s + 0.5) < 1e-3
        near[q] = close.sum()
    assert (np.abs(got - want) <= near).all(), (got - want)[np.abs(got - want) > near]
    assert (got != want).mean() <= 0.05
    views = projection.best_views(t(x), t(poses), 5, mgj.INTRINSIC, mgj.DEPTH_MIN, mgj.DEPTH_MAX, mgj.IMAGE_DIMS)
    if (got == want).all():
        assert views == gold["frustum/best5"].tolist()
    assert views[0] == int(np.argmax(got))


def test_graphed_views_reproduce_the_eager_call(cuda):
    """GraphedViews (lifting + point branch as one CUDA graph) returns what the eager forward_views returns, also after
    new inputs are copied into its static buffers."""
    from pn2_b200 import pointnet_util
    from pn2_b200.models import GraphedViews, PointNet2Multiview2
    B, N, V = 2, 2048, 3
    prev = pointnet_util.set_mlp_precision("bf16")
    try:
        torch.manual_seed(1)
        net = PointNet2Multiview2(21).eval().to(cuda)
        args = (scenes.SCANNET_INTRINSIC, 0.1, 4.0, scenes.SCANNET_IMAGE_DIMS, 0.05)

        def inputs(seed):
            pts = scenes.scannet_batch(seed, B, N)[:, :, :3].astype(np.float32)
            mv = [scenes.multiview_inputs(seed + b, pts[b], V, 128) for b in range(B)]
            return (torch.from_numpy(pts).to(cuda).permute(0, 2, 1).contiguous(), torch.from_numpy(np.stack([m[0] for m in mv])).to(cuda),
                    torch.from_numpy(np.stack([m[1] for m in mv])).to(cuda), torch.from_numpy(np.stack([m[2] for m in mv]).astype(np.float32)).to(cuda))

        a, b = inputs(40), inputs(50)
        g = GraphedViews(net, *a, *args)
        with torch.no_grad():
            for x in (a, b, a):
                want = net.forward_views(*x, *args)
                got = g.run(*x).clone()
                assert torch.equal(got, want)
    finally:
        pointnet_util.set_mlp_precision(prev)
