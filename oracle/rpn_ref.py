"""numpy restatement of iou_spheres / nms (model/pointmaskrcnn.py:233-321), fp32 operation by operation
(TEST INFRASTRUCTURE ONLY)."""
import numpy as np

F = np.float32


def iou_spheres(a, b):
    a, b = np.asarray(a, F), np.asarray(b, F)
    d3 = a[:, None, :3] - b[None, :, :3]
    dist = np.sqrt((d3[..., 0] * d3[..., 0] + d3[..., 1] * d3[..., 1]) + d3[..., 2] * d3[..., 2]).astype(F)
    ra = np.broadcast_to(a[:, None, 3], dist.shape)
    rb = np.broadcast_to(b[None, :, 3], dist.shape)
    iou = np.zeros(dist.shape, F)
    inside = dist <= np.abs(ra - rb)                                  # :247-251
    q = (np.minimum(ra, rb) / np.maximum(ra, rb)).astype(F)
    iou[inside] = (q * q * q)[inside]
    part = (dist > np.abs(ra - rb)) & (dist < ra + rb)                 # :255-261
    with np.errstate(divide="ignore", invalid="ignore"):
        inter = ((ra + rb - dist) * (ra + rb - dist)).astype(F)
        inter = (inter * (dist * dist + F(2) * dist * (ra + rb) - F(3) * ((ra - rb) * (ra - rb)))).astype(F)
        inter = (inter * (F(np.pi) / (F(12) * dist))).astype(F)
        union = (F(4 / 3.0 * np.pi) * (ra * ra * ra + rb * rb * rb) - inter).astype(F)
        val = (inter / union).astype(F)
    iou[part] = val[part]
    return iou


def nms(spheres, scores, threshold=0.7):
    """:290-321 with equal scores ordered by index (torch.sort leaves it unspecified)."""
    scores = np.asarray(scores, F)
    order = np.lexsort((np.arange(len(scores)), -scores.astype(np.float64)))
    table = iou_spheres(spheres, spheres)
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        if order.size == 1:
            break
        rest = order[1:]
        order = rest[table[i, rest] <= threshold]
    return np.asarray(keep, np.int64)
