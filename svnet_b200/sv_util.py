"""Host-side mirror of models/utils/sv_util.py (kNN graph, edge features, pooling, concat).

Same function names and signatures as the reference; all arithmetic runs in the CUDA library.
These are the materialising, module-level entry points (they return the (B,N,k,...) edge tensors
the reference returns).  The model classes do not call them on the hot path: they use the fused
kernels, which never write an edge tensor or the BxNxN distance matrix.
"""
import torch

from . import _native as nv


def knn(x, k):
    """sv_util.py:19-25.  x (B, C, N) (any strides) -> idx (B, N, k) int64, nearest first, self
    included; ties broken by lowest index."""
    B, C, N = x.shape
    feat = x.transpose(2, 1).contiguous()  # (B, N, C) point-major
    _, idx = nv.knn(nv.view_of(feat.view(B * N, C), None), B, N, k, want64=True, want32=False)
    return idx


def _xyz_rows(x):
    batch_size, num_points = x.size(0), x.size(3)
    x = x.view(batch_size, -1, num_points)
    if x.size(1) != 3:
        raise NotImplementedError("svnet_b200 graph features expect one xyz triple per point")
    return x, x.transpose(2, 1).contiguous()  # (B,3,N), (B,N,3)


def get_graph_feature(x, k=20, idx=None, x_coord=None, first=False):
    """sv_util.py:28-62.  x (B,1,3,N) -> (B,N,k,3,2) = [x_j - x_i | x_i]."""
    if first:
        raise NotImplementedError("first=True is not used by any SV model (sv_util.py:55-58)")
    xc, xyz = _xyz_rows(x)
    if idx is None:
        if x_coord is None:
            idx = knn(xc, k=k)
        else:
            idx = knn(x_coord.view(x.size(0), -1, x.size(3)), k=k)
    return nv.graph_feature_xyz(xyz, idx.contiguous(), 2)


def get_graph_feature_cross(x, k=20, idx=None):
    """sv_util.py:64-88.  -> (B,N,k,3,3) = [x_j - x_i | x_i | x_j x x_i]."""
    xc, xyz = _xyz_rows(x)
    if idx is None:
        idx = knn(xc, k=k)
    return nv.graph_feature_xyz(xyz, idx.contiguous(), 3)


def get_graph_feature_sv(x, k=20, idx=None):
    '''
    sv_util.py:90-116
    shape of s: B, N_points, s_dim
    shape of v: B, N_points, 3, v_dim
    '''
    s, v = x
    s, v = s.contiguous(), v.contiguous()
    B, N, Cs = s.shape
    Cv = v.size(-1)
    if idx is None:
        view = nv.view_of(s.view(B * N, Cs), v.view(B * N, 3, Cv))
        _, idx = nv.knn(view, B, N, k, want64=True, want32=False)
    else:
        # the reference accepts the already flattened, batch-offset index (sv_util.py:102-104)
        if idx.dim() == 1:
            base = torch.arange(0, B, device=idx.device).view(-1, 1, 1) * N
            idx = idx.view(B, N, -1) - base
    return nv.graph_feature_sv(s, v, idx.contiguous())


def svpool(x, dim=2, keepdim=False, spool='max'):
    '''
    sv_util.py:118-132
    shape of s: B, N_points, k, s_dim
    shape of v: B, N_points, k, 3, v_dim
    '''
    s, v = x
    if spool not in ('max', 'mean'):
        raise ValueError('not recognized pooling mean {}'.format(spool))
    s, v = s.contiguous(), v.contiguous()
    dim = dim % s.dim()
    outer = 1
    for d in s.shape[:dim]:
        outer *= d
    rows = s.shape[dim]
    cs = s.numel() // (outer * rows)
    cv = v.numel() // (outer * rows)
    smax, smean = nv.pool_rows(s, cs, cs, outer, rows, want_max=spool == 'max', want_mean=spool == 'mean')
    s2 = (smax if spool == 'max' else smean).view(s.shape[:dim] + s.shape[dim + 1:])
    _, vmean = nv.pool_rows(v, cv, cv, outer, rows, want_max=False, want_mean=True)
    v2 = vmean.view(v.shape[:dim] + v.shape[dim + 1:])
    if keepdim:
        s2, v2 = s2.unsqueeze(dim), v2.unsqueeze(dim)
    return (s2, v2)


def svcat(xlist):
    '''
    sv_util.py:134-144
    shape of s: B, N_points, [k,] s_dim
    shape of v: B, N_points, [k,] 3, v_dim
    '''
    s = torch.cat([x[0] for x in xlist], dim=-1)
    v = torch.cat([x[1] for x in xlist], dim=-1)
    return (s, v)
