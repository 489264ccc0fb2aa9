#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported read-only) on CPU.  The reference has no tests or golden vectors of its
own, and it cannot travel to the GPU box, so its outputs on small seeded
inputs are committed here as .npz files together with this script.

Run (in the build container only):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Every fixture stores inputs, the seed of the synthetic state_dict (svnet_b200.synthetic), a digest
of that state_dict, and the reference's outputs.  Fixtures are float32/int64 exactly as torch
produced them.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from svnet_b200.synthetic import (make_args, one_hot_labels, state_dict_digest,  # noqa: E402
                                  synthetic_clouds, synthetic_state_dict)

with contextlib.redirect_stdout(io.StringIO()):
    import models as ref  # noqa: E402
    from models import sv_layers as ref_layers  # noqa: E402
    from models.utils import sv_util as ref_util  # noqa: E402

torch.set_num_threads(1)  # fixed summation order for the fixtures
torch.manual_seed(0)


def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def np_(t):
    return t.detach().cpu().numpy()


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %-40s %7.1f KiB" % (name + ".npz", os.path.getsize(path) / 1024))


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


# ------------------------------------------------------------------------------------------------
# 1. kNN + graph features (sv_util.py:19-116)
# ------------------------------------------------------------------------------------------------
def golden_knn():
    out = {}
    cases = [("xyz", 2, 3, 96, 12, 11), ("c62", 2, 62, 80, 20, 12), ("c127", 1, 127, 64, 8, 13)]
    for name, b, c, n, k, seed in cases:
        x = rnd((b, c, n), seed)
        idx = ref_util.knn(x, k)
        inner = -2 * torch.matmul(x.transpose(2, 1), x)
        xx = torch.sum(x ** 2, dim=1, keepdim=True)
        pd = -xx - inner - xx.transpose(2, 1)
        out[name + "_x"] = np_(x)
        out[name + "_k"] = np.int64(k)
        out[name + "_idx"] = np_(idx)
        out[name + "_pd"] = np_(pd)
    # tie cases: duplicated points and an integer lattice (exact equal distances)
    x = rnd((1, 3, 32), 14)
    x[:, :, 1] = x[:, :, 0]
    x[:, :, 9] = x[:, :, 8]
    out["dup_x"] = np_(x)
    out["dup_k"] = np.int64(6)
    out["dup_idx"] = np_(ref_util.knn(x, 6))
    gx, gy, gz = torch.meshgrid(torch.arange(4.), torch.arange(4.), torch.arange(2.), indexing="ij")
    lat = torch.stack([gx.flatten(), gy.flatten(), gz.flatten()], 0).unsqueeze(0)
    out["lat_x"] = np_(lat)
    out["lat_k"] = np.int64(7)
    out["lat_idx"] = np_(ref_util.knn(lat, 7))
    save("knn", **out)


def golden_graph_features():
    b, n, k = 2, 40, 6
    x = rnd((b, 1, 3, n), 21)
    idx = ref_util.knn(x.view(b, 3, n), k)
    out = dict(x=np_(x), idx=np_(idx), k=np.int64(k))
    out["gf"] = np_(ref_util.get_graph_feature(x, k=k))
    out["gf_cross"] = np_(ref_util.get_graph_feature_cross(x, k=k))
    s = rnd((b, n, 8), 22)
    v = rnd((b, n, 3, 5), 23)
    sf, vf = ref_util.get_graph_feature_sv((s, v), k=k)
    feat = torch.cat([s, v.view(b, n, -1)], dim=-1)
    out.update(s=np_(s), v=np_(v), sf=np_(sf), vf=np_(vf),
               idx_sv=np_(ref_util.knn(feat.transpose(-1, -2), k)))
    sp, vp = ref_util.svpool((sf, vf))
    out.update(pool_s=np_(sp), pool_v=np_(vp))
    sp, vp = ref_util.svpool((sf, vf), spool="mean")
    out.update(pool_mean_s=np_(sp))
    sp, vp = ref_util.svpool((sp, vp), dim=1, keepdim=True)
    out.update(pool1_s=np_(sp), pool1_v=np_(vp))
    save("graph_features", **out)


# ------------------------------------------------------------------------------------------------
# 2. layers (sv_layers.py)
# ------------------------------------------------------------------------------------------------
def load_synth(module, seed, beta_zero=False):
    sd = synthetic_state_dict(module.state_dict(), seed=seed, beta_zero=beta_zero)
    module.load_state_dict(sd)
    module.eval()
    return sd


def golden_layers():
    out = {}
    with torch.no_grad():
        # Linear, binary weights+activations, K not a multiple of 32, beta != 0 and beta == 0
        for tag, bz in (("lin_bin", False), ("lin_bin_b0", True)):
            lin = ref_layers.Linear(70, 24, bias=False, bw=True, ba=True)
            sd = load_synth(lin, 31, beta_zero=bz)
            x = rnd((3, 17, 70), 32)
            x[0, 0, :5] = 0.0                      # exact zeros: sign(0 + beta)
            if not bz:
                x[1, 2, 7] = -sd["beta"][0, 7]     # x + beta == 0 exactly
            y = lin(x)
            out[tag + "_x"] = np_(x)
            out[tag + "_y"] = np_(y)
            out[tag + "_sign"] = np_(torch.sign(x + sd["beta"])).astype(np.int8)
            out[tag + "_digest"] = np.array(state_dict_digest(sd))
        lin = ref_layers.Linear(42, 21, bias=False, bw=True)       # binary weights, fp32 activations
        sd = load_synth(lin, 33)
        x = rnd((2, 9, 3, 42), 34)
        out["lin_bw_x"], out["lin_bw_y"] = np_(x), np_(lin(x))
        conv = quiet(ref_layers.Conv1d, 45, 16, binary=True)
        sd = load_synth(conv, 35)
        x = rnd((2, 45, 19), 36)
        out["conv_bin_x"], out["conv_bin_y"] = np_(x), np_(conv(x))
        # VectorBN
        vbn = ref_layers.VectorBN(11)
        load_synth(vbn, 37)
        v = rnd((2, 7, 5, 3, 11), 38)
        out["vbn_x"], out["vbn_y"] = np_(v), np_(vbn(v))
        # Vector2Scalar fp / binary, 5-d and 3-d inputs, trans_back
        for tag, binary in (("v2s_fp", False), ("v2s_bin", True)):
            m = ref_layers.Vector2Scalar(20, 3, binary=binary, trans_back=True)
            load_synth(m, 39)
            v = rnd((2, 6, 4, 3, 20), 40)
            s, z = m(v)
            out[tag + "_x"], out[tag + "_s"], out[tag + "_z"] = np_(v), np_(s), np_(z)
        # SVBlock fp / binary on edge-shaped and point-shaped inputs
        for tag, binary in (("svb_fp", False), ("svb_bin", True)):
            blk = quiet(ref_layers.SVBlock, (64, 20), (32, 10), binary)
            sd = load_synth(blk, 41)
            s = rnd((2, 12, 5, 64), 42)
            v = rnd((2, 12, 5, 3, 20), 43)
            so, vo = blk((s, v))
            out[tag + "_s"], out[tag + "_v"] = np_(s), np_(v)
            out[tag + "_so"], out[tag + "_vo"] = np_(so), np_(vo)
            so, vo = blk((s[:, :, 0].contiguous(), v[:, :, 0].contiguous()))
            out[tag + "_pt_so"], out[tag + "_pt_vo"] = np_(so), np_(vo)
            if binary:
                sv = blk.v2s(v)
                u = torch.cat([s, sv], dim=-1)
                out[tag + "_sign"] = np_(torch.sign(u + sd["linear1.beta"])).astype(np.int8)
        # SVFuse
        fuse = quiet(ref_layers.SVFuse, 10, 3, True)
        load_synth(fuse, 44)
        s = rnd((2, 9, 16), 45)
        v = rnd((2, 9, 3, 10), 46)
        out["fuse_s"], out["fuse_v"], out["fuse_y"] = np_(s), np_(v), np_(fuse((s, v)))
        # SV_STNkd (small dims are fixed by the class; use dim=(32,10))
        stn = quiet(ref_layers.SV_STNkd, (32, 10), True)
        load_synth(stn, 47)
        s = rnd((2, 24, 32), 48)
        v = rnd((2, 24, 3, 10), 49)
        so, vo = stn((s, v))
        out["stn_s"], out["stn_v"], out["stn_so"], out["stn_vo"] = np_(s), np_(v), np_(so), np_(vo)
    save("layers", **out)


# ------------------------------------------------------------------------------------------------
# 3. whole models with recorded kNN indices and pooled per-point features
# ------------------------------------------------------------------------------------------------
class Recorder:
    """Wraps sv_util.knn and the model module's svpool to record intermediates."""

    def __init__(self, model_module):
        self.mod = model_module
        self.idx, self.pools = [], []

    def __enter__(self):
        self._knn = ref_util.knn
        self._pool = self.mod.svpool

        def knn(x, k):
            r = self._knn(x, k)
            self.idx.append(np_(r))
            return r

        def svpool(x, dim=2, keepdim=False, spool="max"):
            r = self._pool(x, dim=dim, keepdim=keepdim, spool=spool)
            if dim == 2:
                self.pools.append((np_(r[0]), np_(r[1])))
            return r

        ref_util.knn = knn
        self.mod.svpool = svpool
        return self

    def __exit__(self, *a):
        ref_util.knn = self._knn
        self.mod.svpool = self._pool


def golden_model(name, ctor, model_module, args, ncls, b, n, seed, with_label=False, beta_zero=False):
    model = quiet(ctor, args, ncls)
    sd = load_synth(model, seed, beta_zero=beta_zero)
    x = synthetic_clouds(b, n, seed + 1, rotate=True)
    out = dict(x=np_(x), seed=np.int64(seed), k=np.int64(args.k), binary=np.bool_(args.binary),
               ncls=np.int64(ncls), digest=np.array(state_dict_digest(sd)),
               beta_zero=np.bool_(beta_zero),
               sd_shapes=np.array(json.dumps({k_: [list(v_.shape), str(v_.dtype)] for k_, v_ in sd.items()})))
    with torch.no_grad(), Recorder(model_module) as rec:
        if with_label:
            l = one_hot_labels(b)
            out["label"] = np_(l)
            y = model(x, l.view(b, 1, 16) if "pointnet" in name else l)
        else:
            y = model(x)
    out["logits"] = np_(y)
    for i, a in enumerate(rec.idx):
        out["idx%d" % i] = a
    for i, (s, v) in enumerate(rec.pools):
        out["pool%d_s" % i] = s
        out["pool%d_v" % i] = v
    save(name, **out)


def golden_models():
    from models import sv_dgcnn_cls, sv_dgcnn_partseg, sv_pointnet_cls, sv_pointnet_partseg
    A = make_args
    golden_model("dgcnn_cls_bin", ref.SV_DGCNN_CLS, sv_dgcnn_cls, A(k=12, binary=True), 40, 2, 96, 101)
    golden_model("dgcnn_cls_bin_b0", ref.SV_DGCNN_CLS, sv_dgcnn_cls, A(k=12, binary=True), 40, 1, 64,
                 102, beta_zero=True)
    golden_model("dgcnn_cls_fp", ref.SV_DGCNN_CLS, sv_dgcnn_cls, A(k=12, binary=False), 15, 2, 96, 103)
    golden_model("dgcnn_pseg_bin", ref.SV_DGCNN_PSEG, sv_dgcnn_partseg, A(k=10, binary=True), 50, 2, 80,
                 104, with_label=True)
    golden_model("dgcnn_pseg_fp", ref.SV_DGCNN_PSEG, sv_dgcnn_partseg, A(k=10, binary=False), 50, 2, 80,
                 105, with_label=True)
    golden_model("pointnet_cls_fp", ref.SV_PointNet_CLS, sv_pointnet_cls, A(k=12, binary=False), 40, 2, 96,
                 106)
    golden_model("pointnet_cls_bin", ref.SV_PointNet_CLS, sv_pointnet_cls, A(k=12, binary=True), 40, 2, 96,
                 107)
    golden_model("pointnet_pseg_bin", ref.SV_PointNet_PSEG, sv_pointnet_partseg, A(k=10, binary=True), 50,
                 2, 64, 108, with_label=True)
    golden_model("pointnet_pseg_fp", ref.SV_PointNet_PSEG, sv_pointnet_partseg, A(k=10, binary=False), 50,
                 2, 64, 109, with_label=True)


def golden_metrics():
    """utils.calculate_shape_IoU (utils.py:68-91) on random predictions."""
    import utils as ref_utils
    rng = np.random.RandomState(5)
    label = rng.randint(0, 16, size=(6, 1))
    seg = np.zeros((6, 50), dtype=np.int64)
    pred = np.zeros((6, 50), dtype=np.int64)
    for i in range(6):
        lo, n = ref_utils_index_start[label[i, 0]], ref_utils_seg_num[label[i, 0]]
        seg[i] = rng.randint(lo, lo + n, size=50)
        pred[i] = np.where(rng.rand(50) < 0.6, seg[i], rng.randint(0, 50, size=50))
    seg[0] = seg[0][0]          # a shape with a single part present: empty unions count as IoU 1
    pred[0] = seg[0]
    ious = np.array(ref_utils.calculate_shape_IoU(pred, seg, label))
    save("metrics", pred=pred, seg=seg, label=label, ious=ious)


ref_utils_seg_num = [4, 2, 2, 4, 4, 3, 3, 2, 4, 2, 6, 2, 3, 3, 3, 3]
ref_utils_index_start = [0, 4, 6, 8, 12, 16, 19, 22, 24, 28, 30, 36, 38, 41, 44, 47]


if __name__ == "__main__":
    if "--metrics-only" in sys.argv:
        golden_metrics()
        sys.exit(0)
    golden_knn()
    golden_graph_features()
    golden_layers()
    golden_models()
    golden_metrics()
