"""SV-PointNet part segmentation -- drop-in for models/sv_pointnet_partseg.py:12-97 (same
constructor and state_dict keys).  Fused xyz edge layer, then per-point SVBlocks on rows; the
equivariant->invariant frame projection einsum('bimj,bijk->bimk') (:93) is rows_prep with the frames
supplied (z_in); per-cloud-constant head channels are pre-reduced per cloud.
"""
import torch
import torch.nn as nn

from . import _native as nv
from .fused import chunked, first_edge_layer
from .sv_dgcnn_partseg import _Seq
from .sv_layers import Conv1d, SV_STNkd, SVBlock, SVFuse, Vector2Scalar, _inference_only, dense_rows
from .sv_pointnet_cls import _bcast_rows, stn_rows


class SV_PointNet_PSEG(nn.Module):
    def __init__(self, args, num_part=50):
        super(SV_PointNet_PSEG, self).__init__()
        self.k = args.k
        self.binary = args.binary

        self.init_scalar = Vector2Scalar(3, 3)
        self.conv_pos = SVBlock((9, 3), (64//2, 64//6))
        self.conv1 = SVBlock((64//2, 64//6), (64//2, 64//6), binary=self.binary)
        self.conv2 = SVBlock((64//2, 64//6), (128//2, 128//6), binary=self.binary)
        self.conv3 = SVBlock((128//2, 128//6), (128//2, 128//6), binary=self.binary)
        self.fstn = SV_STNkd((128//2, 128//6), binary=self.binary)
        self.conv4 = SVBlock((128//2*2, 128//6*2), (512//2, 512//6), binary=self.binary)
        self.conv5 = SVBlock((512//2, 512//6), (2048//2, 2048//6), binary=self.binary)

        self.svfuse = SVFuse(2048//6*2, 3, binary=self.binary, trans_back=True)
        self.channels = 2048//2*2+2048//6*2*3
        self.conv_fuse1 = _Seq(
                Conv1d(self.channels, self.channels//8, binary=self.binary),
                nn.BatchNorm1d(self.channels//8),
                nn.ReLU(inplace=True)
                )
        self.conv_fuse2 = _Seq(
                Conv1d(self.channels//8, self.channels, binary=self.binary),
                nn.BatchNorm1d(self.channels),
                nn.ReLU(inplace=True)
                )
        self.convs1 = _Seq(
                Conv1d(self.channels+16+64//2+128//2*2+512//2+2048//2+(64//6+128//6*2+512//6+2048//6)*3, 256, binary=self.binary),
                nn.BatchNorm1d(256),
                nn.ReLU(inplace=True))
        self.convs2 = _Seq(
                Conv1d(256, 256, binary=self.binary),
                nn.BatchNorm1d(256),
                nn.ReLU(inplace=True))
        self.convs3 = _Seq(
                Conv1d(256, 128, binary=self.binary),
                nn.BatchNorm1d(128),
                nn.ReLU(inplace=True))
        self.convs4 = nn.Conv1d(128, num_part, 1)

    def forward(self, x, l, forced_idx=None, record=None):
        hooks = forced_idx is not None or record is not None
        return chunked(lambda xc, lc: self._forward(xc, lc, forced_idx, record), x, (l,), hooks=hooks)

    def _forward(self, x, l, forced_idx=None, record=None):
        """x (B,3,N), l (B,1,16) or (B,16) -> (B, num_part, N)   (sv_pointnet_partseg.py:55-97)"""
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
        # the head concatenates out1..out5 (:92): write every block's output straight into that table
        dims = [self.conv1.out_dims, self.conv2.out_dims, self.conv3.out_dims, self.conv4.out_dims, self.conv5.out_dims]
        cs_tot, cv_tot = sum(d[0] for d in dims), sum(d[1] for d in dims)
        cat_s = torch.empty((R, cs_tot), dtype=torch.float32, device=dev)
        cat_v = torch.empty((R, 3, cv_tot), dtype=torch.float32, device=dev)
        offs, so, vo = [], 0, 0
        for d in dims:
            offs.append((so, vo))
            so += d[0]
            vo += d[1]

        def sl(i):
            (a, b), (cs, cv) = offs[i], dims[i]
            return cat_s[:, a:a + cs], cat_v[:, :, b:b + cv]

        s1, v1 = sl(0)
        self.conv1.forward_rows(s0, v0, B, N, s_out=s1, lds_out=cat_s.stride(0), v_out=v1)
        s2, v2 = sl(1)
        self.conv2.forward_rows(s1, v1, B, N, s_out=s2, lds_out=cat_s.stride(0), v_out=v2)
        s3, v3 = sl(2)
        self.conv3.forward_rows(s2, v2, B, N, s_out=s3, lds_out=cat_s.stride(0), v_out=v3)
        sg, vg = stn_rows(self.fstn, s3, v3, B, N)
        Cs3, Cv3 = dims[2]
        st = torch.empty((R, 2 * Cs3), dtype=torch.float32, device=dev)
        vt = torch.empty((R, 3, 2 * Cv3), dtype=torch.float32, device=dev)
        st[:, :Cs3].copy_(s3)
        vt[:, :, :Cv3].copy_(v3)
        _bcast_rows(st[:, Cs3:], sg, N)
        _bcast_rows(vt[:, :, Cv3:], vg, N)
        s4, v4 = sl(3)
        self.conv4.forward_rows(st, vt, B, N, s_out=s4, lds_out=cat_s.stride(0), v_out=v4)
        s5, v5 = sl(4)
        self.conv5.forward_rows(s4, v4, B, N, s_out=s5, lds_out=cat_s.stride(0), v_out=v5)
        # svcat([out5, mean over points]) -> svfuse with trans_back (:77-80)
        Cs5, Cv5 = dims[4]
        sc = torch.empty((R, 2 * Cs5), dtype=torch.float32, device=dev)
        vc = torch.empty((R, 3, 2 * Cv5), dtype=torch.float32, device=dev)
        sc[:, :Cs5].copy_(s5)
        vc[:, :, :Cv5].copy_(v5)
        _, smean = nv.pool_rows(s5, cat_s.stride(0), Cs5, B, N, want_max=False, want_mean=True)
        vmean = torch.empty((B, 3, Cv5), dtype=torch.float32, device=dev)
        for a in range(3):
            nv.pool_rows(v5[:, a, :], cat_v.stride(0), Cv5, B, N, want_max=False, want_mean=True,
                         mean_out=vmean[:, a, :], ldo=3 * Cv5)
        _bcast_rows(sc[:, Cs5:], smean, N)
        _bcast_rows(vc[:, :, Cv5:], vmean, N)
        fused, trans = self.svfuse.forward_rows(sc, vc, want_z=True)              # (R, channels), (R,3,3)
        h = self.conv_fuse1[0].forward_rows(fused, bn=self.conv_fuse1.bn_folded(), act=nv.ACT_RELU)
        h = self.conv_fuse2[0].forward_rows(h, bn=self.conv_fuse2.bn_folded(), act=nv.ACT_RELU)
        C = h.shape[1]
        x_l = torch.empty((B, C + 16), dtype=torch.float32, device=dev)
        if self.binary:
            nv.pool_rows(h, C, C, B, N, want_max=False, want_mean=True, mean_out=x_l, ldo=C + 16)   # :84-85
        else:
            nv.pool_rows(h, C, C, B, N, want_max=True, max_out=x_l, ldo=C + 16)                     # :86-87
        x_l[:, C:].copy_(l.reshape(B, -1).float())
        # invariant per-point features: [cat_s | v^T . trans] (:92-94), frames supplied to rows_prep
        concat = torch.empty((R, cs_tot + 3 * cv_tot), dtype=torch.float32, device=dev)
        nv.rows_prep(nv.view_of(cat_s, cat_v), R, z_in=trans, u_out=concat, ldu=concat.stride(0))
        net = self.convs1[0].forward_rows(concat, bn=self.convs1.bn_folded(), act=nv.ACT_RELU, cloud=x_l, rows_per_cloud=N)
        net = self.convs2[0].forward_rows(net, bn=self.convs2.bn_folded(), act=nv.ACT_RELU)
        net = self.convs3[0].forward_rows(net, bn=self.convs3.bn_folded(), act=nv.ACT_RELU)
        out = dense_rows(self.convs4.weight, net, bias=self.convs4.bias)
        if record is not None:
            record.update(concat=concat, x_l=x_l)
        return out.view(B, N, -1).transpose(1, 2).contiguous()
