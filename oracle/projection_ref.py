"""CPU restatement of the reference's projection index computation and view reduction, op for op with torch
on the host (TEST INFRASTRUCTURE ONLY).  Follows utils/projection.py:25-52 (corners), :54-95 (normals),
:97-130 (frustum mask), :166-230 (compute_projection), :237-256 (Projection.forward) and
model/pointnet2multiview.py:30-43 / 83-102 (view reductions).  Unlike oracle/pn2_oracle.c this keeps the
reference's torch.mm / elementwise op sequence, so it measures how often the fixed fma order of the kernel
lands on a different pixel than BLAS-ordered arithmetic would (tests/test_lifting_gpu.py)."""
import torch


def corners_ref(intrinsic, dmin, dmax, image_dims, c2w):
    pts = c2w.new_ones(8, 4, 1)
    W, H = image_dims

    def skel(ux, uy, d):
        x = (ux - intrinsic[0][2]) / intrinsic[0][0]
        y = (uy - intrinsic[1][2]) / intrinsic[1][1]
        return torch.Tensor([d * x, d * y, d])

    for i, (u, v, d) in enumerate([(0, 0, dmin), (W - 1, 0, dmin), (W - 1, H - 1, dmin), (0, H - 1, dmin),
                                   (0, 0, dmax), (W - 1, 0, dmax), (W - 1, H - 1, dmax), (0, H - 1, dmax)]):
        pts[i][:3] = skel(u, v, d).unsqueeze(1)
    return torch.bmm(c2w.repeat(8, 1, 1), pts)


def normals_ref(cc):
    c = cc.reshape(8, 4)[:, :3]
    pairs = [(0, 3, 0, 1), (1, 2, 1, 5), (2, 3, 2, 6), (3, 0, 3, 7), (0, 1, 0, 4), (5, 6, 5, 4)]
    return torch.stack([torch.linalg.cross(c[a1] - c[a0], c[b1] - c[b0]) for a0, a1, b0, b1 in pairs])


def compute_projection_ref(points, depth, c2w, intrinsic, dmin, dmax, image_dims, accuracy):
    """-> pix (N,) int64 with -1 where the point is not lifted (the reference's packed vectors carry the same
    information: indices_3d = nonzero(pix >= 0), indices_2d = pix[indices_3d])."""
    N = points.shape[0]
    w2c = torch.inverse(c2w)
    coords = c2w.new_empty(4, N)
    coords[:3, :] = points.t()
    coords[3, :].fill_(1)
    cc = corners_ref(intrinsic, dmin, dmax, image_dims, c2w)
    normals = normals_ref(cc)
    c = cc.reshape(8, 4)
    mask = torch.ones(N, dtype=torch.bool)
    for k in range(6):
        rel = points - (c[2][:3] if k < 3 else c[4][:3])
        mask &= (torch.round(torch.mm(rel, normals[k].unsqueeze(1)) * 100) / 100 < 0).squeeze(1)
    pix = torch.full((N,), -1, dtype=torch.int64)
    if not mask.any():
        return pix
    ind = torch.nonzero(mask).squeeze(1)
    cam = torch.mm(w2c, coords[:, ind])
    cam[0] = (cam[0] * intrinsic[0][0]) / cam[2] + intrinsic[0][2]
    cam[1] = (cam[1] * intrinsic[1][1]) / cam[2] + intrinsic[1][2]
    img = torch.round(cam).long()
    valid = (img[0] >= 0) & (img[1] >= 0) & (img[0] < image_dims[0]) & (img[1] < image_dims[1])
    if not valid.any():
        return pix
    p2 = img[1][valid] * image_dims[0] + img[0][valid]
    dv = depth.reshape(-1)[p2]
    dm = (dv >= dmin) & (dv <= dmax) & ((dv - cam[2][valid]).abs() <= accuracy)
    pix[ind[valid][dm]] = p2[dm]
    return pix


def lift_ref(feats, pix_per_view, reduce):
    """feats (V, C, H, W), pix_per_view list of (N,) -> (C, N)"""
    V, C = feats.shape[:2]
    N = pix_per_view[0].shape[0]
    maps = []
    for v in range(V):
        out = feats.new_zeros(C, N)
        keep = torch.nonzero(pix_per_view[v] >= 0).squeeze(1)
        if keep.numel():
            out[:, keep] = feats[v].reshape(C, -1)[:, pix_per_view[v][keep]]
        maps.append(out)
    stack = torch.stack(maps, dim=2)  # (C, N, V)
    if reduce == "max":
        return stack.max(dim=2)[0]
    final = stack[:, :, 0].clone()
    for j in range(1, V):
        m = torch.nonzero((final == 0).sum(0) == C).squeeze(1)
        final[:, m] = stack[:, m, j]
    return final
