"""numpy restatement of utils/pc_util.py:39-51 and of the voxel-wise counters of train_scannet_semseg.py:225-239
(TEST INFRASTRUCTURE ONLY: tests/ compare pn2_b200.pc_util against it)."""
import numpy as np


def point_cloud_label_to_surface_voxel_label_fast(point_cloud, label, res=0.0484):
    """utils/pc_util.py:39-51, statement by statement (float32 in, float32 arithmetic as numpy performs it)."""
    coordmax = np.max(point_cloud, axis=0)
    coordmin = np.min(point_cloud, axis=0)
    nvox = np.ceil((coordmax - coordmin) / res)
    vidx = np.ceil((point_cloud - coordmin) / res)
    vidx = vidx[:, 0] + vidx[:, 1] * nvox[0] + vidx[:, 2] * nvox[0] * nvox[1]
    uvidx, vpidx = np.unique(vidx, return_index=True)
    uvlabel = label[vpidx] if label.ndim == 1 else label[vpidx, :]
    return uvidx, uvlabel, nvox


def voxel_accuracy_counts(points_np, target_np, pred_val, weights_np, num_classes, res=0.02):
    """train_scannet_semseg.py:225-239"""
    out = {"total_correct_vox": 0, "total_seen_vox": 0, "labelweights_vox": np.zeros(num_classes, dtype=np.int64),
           "total_seen_class_vox": np.zeros(num_classes, dtype=np.int64), "total_correct_class_vox": np.zeros(num_classes, dtype=np.int64),
           "total_union_class_vox": np.zeros(num_classes, dtype=np.int64)}
    for b in range(target_np.shape[0]):
        m = weights_np[b, :] > 0
        _, uvlabel, _ = point_cloud_label_to_surface_voxel_label_fast(
            points_np[b, m, :], np.concatenate((np.expand_dims(target_np[b, m], 1), np.expand_dims(pred_val[b, m], 1)), axis=1), res=res)
        out["total_correct_vox"] += np.sum((uvlabel[:, 0] == uvlabel[:, 1]) & (uvlabel[:, 0] > 0))
        out["total_seen_vox"] += np.sum(uvlabel[:, 0] > 0)
        tmp, _ = np.histogram(uvlabel[:, 0], range(num_classes + 1))
        out["labelweights_vox"] += tmp
        for l in range(num_classes):
            out["total_seen_class_vox"][l] += np.sum(uvlabel[:, 0] == l)
            out["total_correct_class_vox"][l] += np.sum((uvlabel[:, 0] == l) & (uvlabel[:, 1] == l))
            out["total_union_class_vox"][l] += np.sum((uvlabel[:, 0] == l) | (uvlabel[:, 1] == l))
    return out
