"""Caller-side helpers of the eval loops: checkpoint container, SO(3)
test-time rotation, accuracy / shape-IoU.  Mirrors what main_cls_dgcnn.py:218-251,
main_partseg_dgcnn.py:225-279 and utils.py:68-91,118-138 do around ``model(data)``.
"""
import numpy as np
import torch

from . import _native as nv
from .synthetic import strip_module_prefix


def load_checkpoint(path, map_location="cpu"):
    """utils.load_checkpoint (utils.py:118-138): returns the container dict
    {'epoch','state_dict','optimizer','scheduler','best_test_acc'|'best_test_iou'}."""
    state = torch.load(path, map_location=map_location)
    if "state_dict" not in state:       # a bare state_dict is accepted as well
        state = {"epoch": 0, "state_dict": state}
    return state


def load_weights(model, path_or_state):
    """model.load_state_dict(checkpoint['state_dict']) for a model that is NOT wrapped in DataParallel
    (strips the 'module.' prefix the reference's checkpoints carry, main_cls_dgcnn.py:125,210)."""
    state = load_checkpoint(path_or_state) if isinstance(path_or_state, str) else path_or_state
    sd = state["state_dict"] if "state_dict" in state else state
    model.load_state_dict(strip_module_prefix(sd))
    return state


def random_rotations(n, generator=None, device="cpu"):
    """Uniform random rotation matrices (n,3,3), det = +1 (what pytorch3d.random_rotations provides;
    pytorch3d is not installable here).  QR of a Gaussian matrix with the sign fix of Mezzadri (2007)."""
    a = torch.randn(n, 3, 3, generator=generator)
    q, r = torch.linalg.qr(a)
    q = q * torch.sign(torch.diagonal(r, dim1=1, dim2=2)).unsqueeze(1)
    det = torch.linalg.det(q)
    q[:, :, 2] = q[:, :, 2] * det.sign().view(n, 1)
    return q.to(device)


def rotate_points(data, R=None):
    """(B,N,3) points [, (B,3,3) rotations] -> (B,3,N) model input: transform_points + permute
    (main_cls_dgcnn.py:229-235) in one kernel."""
    return nv.rotate_permute(data, R)


def classification_metrics(true, pred, num_class=None):
    """accuracy and balanced (mean per-class) accuracy, as sklearn's accuracy_score /
    balanced_accuracy_score in main_cls_dgcnn.py:248-249."""
    true, pred = np.asarray(true), np.asarray(pred)
    acc = float((true == pred).mean())
    classes = np.unique(true)
    per = [float((pred[true == c] == c).mean()) for c in classes]
    return acc, float(np.mean(per))


SEG_NUM = [4, 2, 2, 4, 4, 3, 3, 2, 4, 2, 6, 2, 3, 3, 3, 3]
INDEX_START = [0, 4, 6, 8, 12, 16, 19, 22, 24, 28, 30, 36, 38, 41, 44, 47]


def calculate_shape_IoU(pred_np, seg_np, label, class_choice=None):
    """ShapeNet-part instance IoU (utils.py:68-91): mean over the parts of the shape's category of
    |pred==part & gt==part| / |pred==part | gt==part| (1 when the union is empty)."""
    label = np.asarray(label).reshape(-1)
    out = []
    for i in range(seg_np.shape[0]):
        cat = int(label[i]) if not class_choice else int(label[0])
        parts = range(INDEX_START[cat], INDEX_START[cat] + SEG_NUM[cat]) if not class_choice else range(SEG_NUM[cat])
        ious = []
        for part in parts:
            inter = np.sum((pred_np[i] == part) & (seg_np[i] == part))
            union = np.sum((pred_np[i] == part) | (seg_np[i] == part))
            ious.append(1.0 if union == 0 else inter / float(union))
        out.append(float(np.mean(ious)))
    return out
