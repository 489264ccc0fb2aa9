"""SV-PointNet classifier -- drop-in for models/sv_pointnet_cls.py:12-81 (same constructors and
state_dict keys).  The first (kNN edge) layer is the fused xyz kernel with the cross-product
channel; everything after it is per-point / per-cloud rows through the row kernels.
"""
import torch
import torch.nn as nn

from . import _native as nv
from .fused import chunked, first_edge_layer
from .sv_layers import (Linear, SV_STNkd, SVBlock, SVFuse, Vector2Scalar, _Cached, _inference_only, dense_rows,
                        folded_bn, head_layer)


def _bcast_rows(dst, src, N):
    """dst (B*N, ...) <- src (B, ...) repeated over the N points of each cloud (expand_as, :45,52)."""
    B = src.shape[0]
    dst.view((B, N) + tuple(dst.shape[1:])).copy_(src.unsqueeze(1).expand((B, N) + tuple(src.shape[1:])))


def stn_rows(stn, s, v, B, N):
    """SV_STNkd.forward on rows (sv_layers.py:234-244): 3 per-point blocks, svpool over points,
    3 per-cloud blocks."""
    s, v = stn.conv1.forward_rows(s, v, B, N)
    s, v = stn.conv2.forward_rows(s, v, B, N)
    s, v = stn.conv3.forward_rows(s, v, B, N)
    Cs, Cv = stn.conv3.out_dims
    sp, _ = nv.pool_rows(s, Cs, Cs, B, N, want_max=True)
    _, vp = nv.pool_rows(v, 3 * Cv, 3 * Cv, B, N, want_max=False, want_mean=True)
    vp = vp.view(B, 3, Cv)
    s, v = stn.fc1.forward_rows(sp, vp, B, 1)
    s, v = stn.fc2.forward_rows(s, v, B, 1)
    return stn.fc3.forward_rows(s, v, B, 1)


class SVPointNetEncoder(nn.Module):
    def __init__(self, k, binary):
        super(SVPointNetEncoder, self).__init__()
        self.k = k
        self.binary=binary

        self.init_scalar = Vector2Scalar(3, 3)
        self.conv_pos = SVBlock((9, 3), (64//2, 64//6))
        self.conv1 = SVBlock((64//2, 64//6), (64//2, 64//6), binary=binary)

        self.fstn = SV_STNkd((64//2, 64//6), binary=binary)

        self.conv2 = SVBlock((64//2*2, 64//6*2), (128//2, 128//6), binary=binary)
        self.conv3 = SVBlock((128//2, 128//6), (1024//2, 1024//6), binary=binary)

        self.conv_fuse = SVBlock((1024//2*2, 1024//6*2), (1024//2, 1024//6), binary=binary)

        self.svfuse = SVFuse(1024//6, 3, binary=binary)

    def forward(self, x, forced_idx=None, record=None):
        """x (B,3,N) -> global feature (B, 1022)   (sv_pointnet_cls.py:31-58)"""
        _inference_only(self)
        B, D, N = x.size()
        R = B * N
        dev = x.device
        xyz = x.transpose(1, 2).contiguous().view(R, 3)
        Cs0, Cv0 = self.conv_pos.out_dims
        s0 = torch.empty((R, Cs0), dtype=torch.float32, device=dev)
        v0 = torch.empty((R, 3, Cv0), dtype=torch.float32, device=dev)
        idx = first_edge_layer(xyz, B, N, self.k, 3, self.init_scalar, self.conv_pos, s0, v0,
                               idx32=forced_idx[0] if forced_idx else None)
        if record is not None:
            record["idx"] = [idx]
            record["pool0"] = (s0, v0)
        # conv1 writes the first half of the svcat([x, x_global]) table; the STN output fills the rest
        Cs1, Cv1 = self.conv1.out_dims
        s_cat = torch.empty((R, 2 * Cs1), dtype=torch.float32, device=dev)
        v_cat = torch.empty((R, 3, 2 * Cv1), dtype=torch.float32, device=dev)
        s1, v1 = s_cat[:, :Cs1], v_cat[:, :, :Cv1]
        self.conv1.forward_rows(s0, v0, B, N, s_out=s1, lds_out=s_cat.stride(0), v_out=v1)
        sg, vg = stn_rows(self.fstn, s1, v1, B, N)
        _bcast_rows(s_cat[:, Cs1:], sg, N)
        _bcast_rows(v_cat[:, :, Cv1:], vg, N)
        s2, v2 = self.conv2.forward_rows(s_cat, v_cat, B, N)
        # conv3 writes the first half of svcat([x, x_mean])
        Cs3, Cv3 = self.conv3.out_dims
        s_cat2 = torch.empty((R, 2 * Cs3), dtype=torch.float32, device=dev)
        v_cat2 = torch.empty((R, 3, 2 * Cv3), dtype=torch.float32, device=dev)
        s3, v3 = s_cat2[:, :Cs3], v_cat2[:, :, :Cv3]
        self.conv3.forward_rows(s2, v2, B, N, s_out=s3, lds_out=s_cat2.stride(0), v_out=v3)
        sm, _ = nv.pool_rows(s3, s_cat2.stride(0), Cs3, B, N, want_max=True)
        vm = torch.empty((B, 3, Cv3), dtype=torch.float32, device=dev)
        for a in range(3):  # v3 is a strided slice: pool each xyz row of the table separately
            nv.pool_rows(v3[:, a, :], v_cat2.stride(0), Cv3, B, N, want_max=False, want_mean=True,
                         mean_out=vm[:, a, :], ldo=3 * Cv3)
        _bcast_rows(s_cat2[:, Cs3:], sm, N)
        _bcast_rows(v_cat2[:, :, Cv3:], vm, N)
        sf, vf = self.conv_fuse.forward_rows(s_cat2, v_cat2, B, N)
        Csf, Cvf = self.conv_fuse.out_dims
        sp, _ = nv.pool_rows(sf, Csf, Csf, B, N, want_max=True)
        _, vp = nv.pool_rows(vf, 3 * Cvf, 3 * Cvf, B, N, want_max=False, want_mean=True)
        out, _ = self.svfuse.forward_rows(sp, vp.view(B, 3, Cvf))
        return out


class SV_PointNet_CLS(_Cached, nn.Module):
    def __init__(self, args, num_class=40):
        super(SV_PointNet_CLS, self).__init__()
        self.binary = args.binary
        self.k = args.k
        p = 0 if self.binary else 0.4

        self.feat = SVPointNetEncoder(k=self.k, binary=self.binary)
        self.fc1 = Linear(1024//2+1024//6*3, 512, bias=False, bw=self.binary, ba=self.binary)
        self.fc2 = Linear(512, 256, bias=False, bw=self.binary, ba=self.binary)
        self.fc3 = nn.Linear(256, num_class)
        self.dropout = nn.Dropout(p=p)
        self.bn1 = nn.BatchNorm1d(512)
        self.bn2 = nn.BatchNorm1d(256)
        self.relu = nn.ReLU()

    def forward(self, x, forced_idx=None, record=None):
        hooks = forced_idx is not None or record is not None
        return chunked(lambda xc: self._forward(xc, forced_idx, record), x, hooks=hooks)

    def _forward(self, x, forced_idx=None, record=None):
        _inference_only(self)
        f = self.feat(x, forced_idx=forced_idx, record=record)
        # dropout is the identity in eval; ReLU (not LeakyReLU) in this head (:78-79)
        return nv.head_fwd(f, [head_layer(self.fc1, folded_bn(self, "bn1"), nv.ACT_RELU),
                               head_layer(self.fc2, folded_bn(self, "bn2"), nv.ACT_RELU),
                               head_layer(self.fc3)])
