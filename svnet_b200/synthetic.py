"""Synthetic checkpoints and point clouds for parity tests and the benchmark.

The 13 committed ``checkpoints/*.pth`` of the reference are stripped from the mount
(/root/reference/.MISSING_LARGE_BLOBS:1-13), so every parity/bench run uses a synthetic
``state_dict`` with the reference's exact key/shape layout.  The values
are drawn *per key* from a key-seeded generator, so the same dict is produced for the reference
model, for the oracle and for the svnet_b200 modules, independent of construction order.

Checkpoint container format mirrors main_cls_dgcnn.py:208-214 / utils.py:141-155
(``{'epoch','state_dict','optimizer','scheduler','best_test_acc'}`` with ``module.``-prefixed keys).
"""
import hashlib
import math
from types import SimpleNamespace

import torch


def _key_generator(seed, key):
    h = hashlib.sha256(("%d:%s" % (seed, key)).encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def synthetic_state_dict(template, seed=0, beta_zero=False):
    """Fill ``template`` (a state_dict: key -> tensor, only shapes/dtypes are used) with
    representative trained-like values:

    * ``*.beta``  ~ N(0, 0.1^2) and never exactly 0 (``beta_zero=True`` gives the untrained
      beta == 0 case that exercises ``sign(0) == 0``, sv_layers.py:25,39),
    * ``*.scale`` ~ U(0.5, 1.5)/sqrt(K),
    * BN ``running_mean`` ~ N(0, 0.1^2), ``running_var`` ~ U(0.5, 1.5), ``weight`` ~ U(0.5, 1.5),
      ``bias`` ~ N(0, 0.1^2),
    * every other weight/bias ~ U(-1, 1)/sqrt(fan_in)  (torch's default Linear/Conv init bound).
    """
    out = {}
    keys = list(template.keys())
    for key in keys:
        t = template[key]
        shape = tuple(t.shape)
        g = _key_generator(seed, key)
        leaf = key.rsplit(".", 1)[-1]
        parent = key.rsplit(".", 1)[0] + "." if "." in key else ""
        is_bn = (parent + "running_mean") in template
        if leaf == "num_batches_tracked":
            out[key] = torch.zeros(shape, dtype=t.dtype)
            continue
        if leaf == "beta":
            if beta_zero:
                v = torch.zeros(shape)
            else:
                v = torch.randn(shape, generator=g) * 0.1
                v = torch.where(v == 0, torch.full_like(v, 0.05), v)
        elif leaf == "scale":
            w = template[parent + "weight"]
            k_in = w.shape[1]
            v = (torch.rand(shape, generator=g) + 0.5) / math.sqrt(k_in)
        elif is_bn and leaf == "running_mean":
            v = torch.randn(shape, generator=g) * 0.1
        elif is_bn and leaf == "running_var":
            v = torch.rand(shape, generator=g) + 0.5
        elif is_bn and leaf == "weight":
            v = torch.rand(shape, generator=g) + 0.5
        elif is_bn and leaf == "bias":
            v = torch.randn(shape, generator=g) * 0.1
        elif leaf == "weight":
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            v = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(max(fan_in, 1))
            # sign(0) on a weight never happens after training; keep the synthetic weights away
            # from exact zero as well.
            v = torch.where(v == 0, torch.full_like(v, 1e-3), v)
        elif leaf == "bias":
            w = template.get(parent + "weight")
            fan_in = 1
            if w is not None:
                for d in w.shape[1:]:
                    fan_in *= d
            v = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(max(fan_in, 1))
        else:
            v = torch.randn(shape, generator=g) * 0.1
        out[key] = v.to(t.dtype).contiguous()
    return out


def wrap_checkpoint(state_dict, metric_key="best_test_acc"):
    """The container the reference writes (main_cls_dgcnn.py:208-214): saved from a DataParallel
    wrapper, hence the ``module.`` prefix."""
    return {
        "epoch": 0,
        "state_dict": {"module." + k: v for k, v in state_dict.items()},
        "optimizer": {},
        "scheduler": {},
        metric_key: 0.0,
    }


def strip_module_prefix(state_dict):
    return {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}


def synthetic_clouds(batch, n_points, seed, rotate=False):
    """p ~ N(0,1)^{BxNx3}; centre, scale to unit max radius (data.py:15-20);
    optional random SO(3) rotation (main_cls_dgcnn.py:229-234); returned as (B,3,N) float32
    (main_cls_dgcnn.py:235)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    p = torch.randn(batch, n_points, 3, generator=g)
    p = p - p.mean(dim=1, keepdim=True)
    r = p.norm(dim=2).max(dim=1)[0].view(batch, 1, 1)
    p = p / r
    if rotate:
        a = torch.randn(batch, 3, 3, generator=g)
        q, _ = torch.linalg.qr(a)
        det = torch.linalg.det(q)
        q[:, :, 0] = q[:, :, 0] * det.sign().view(batch, 1)
        p = torch.bmm(p, q)
    return p.permute(0, 2, 1).contiguous().float()


def one_hot_labels(batch, n_cat=16):
    """Part-seg category one-hot (B,16) float32, class = i mod 16."""
    l = torch.zeros(batch, n_cat)
    l[torch.arange(batch), torch.arange(batch) % n_cat] = 1.0
    return l


def make_args(k=20, binary=False, dropout=0.5):
    """The attribute bag the reference constructors read (sv_dgcnn_cls.py:25-26,
    sv_dgcnn_partseg.py:46)."""
    return SimpleNamespace(k=k, binary=binary, dropout=dropout)


def state_dict_digest(state_dict):
    h = hashlib.sha256()
    for k in sorted(state_dict.keys()):
        h.update(k.encode())
        h.update(state_dict[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()
