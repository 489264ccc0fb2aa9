"""Python face of the CPU oracle (oracle/svnet_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package svnet_b200 never does.

The arithmetic lives in C (svnet_oracle.c, one function per reference function); this file only
sequences those functions the way the reference's ``forward`` methods do and moves numpy arrays
around (concatenate / broadcast / transpose are exact).  Each function cites the reference lines it
follows (paths relative to /root/reference).

Parity pin: tests/test_oracle_golden.py checks every function here against tests/golden/*.npz,
which were produced by the unmodified reference (tests/golden/make_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsvnet_oracle.so")
_LIB = None

FLAG_BW, FLAG_BA = 1, 2
BN_EPS = 1e-5
ACT_NONE, ACT_LEAKY, ACT_RELU = 0, 1, 2


def build(force=False):
    """gcc the C restatement (recipe also in oracle/Makefile)."""
    src = os.path.join(_HERE, "svnet_oracle.c")
    if (not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src)):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-mfma", "-mavx2", "-fopenmp", "-shared", "-fPIC",
           "-o", _SO, src, "-lm"]
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            build()
        _LIB = ctypes.CDLL(_SO)
    return _LIB


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(n):
    return ctypes.c_int(int(n))


def _l(n):
    return ctypes.c_long(int(n))


# ------------------------------------------------------------------------------------------------
# leaf functions
# ------------------------------------------------------------------------------------------------
def knn(feat, k, return_pd=False):
    """sv_util.py:19-25.  feat (B,N,C) point-major -> idx (B,N,k) int64."""
    feat = _f(feat)
    B, N, C = feat.shape
    idx = np.empty((B, N, k), dtype=np.int64)
    pd = np.empty((B, N, N), dtype=np.float32) if return_pd else None
    lib().orc_knn(_p(feat), _c(B), _c(N), _c(C), _c(k), _p(idx), _p(pd))
    return (idx, pd) if return_pd else idx


def graph_feature_xyz(xyz, idx, nv):
    """sv_util.py:28-62 (nv=2) / 64-88 (nv=3).  xyz (B,N,3) -> (B,N,k,3,nv)."""
    xyz = _f(xyz)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    B, N, k = idx.shape
    out = np.empty((B, N, k, 3, nv), dtype=np.float32)
    lib().orc_graph_feature_xyz(_p(xyz), _p(idx), _c(B), _c(N), _c(k), _c(nv), _p(out))
    return out


def graph_feature_sv(s, v, idx):
    """sv_util.py:106-114."""
    s, v = _f(s), _f(v)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    B, N, k = idx.shape
    Cs, Cv = s.shape[-1], v.shape[-1]
    sf = np.empty((B, N, k, 2 * Cs), dtype=np.float32)
    vf = np.empty((B, N, k, 3, 2 * Cv), dtype=np.float32)
    lib().orc_graph_feature_sv(_p(s), _p(v), _p(idx), _c(B), _c(N), _c(k), _c(Cs), _c(Cv), _p(sf), _p(vf))
    return sf, vf


def linear(x, W, beta=None, scale=None, bias=None, bw=False, ba=False):
    """sv_layers.py:29-53.  x (...,K) -> (...,Cout)."""
    x, W = _f(x), _f(W)
    K = x.shape[-1]
    Cout = W.shape[0]
    R = x.size // K
    y = np.empty(x.shape[:-1] + (Cout,), dtype=np.float32)
    beta = _f(beta).reshape(-1) if beta is not None else None
    scale = _f(scale).reshape(-1) if scale is not None else None
    bias = _f(bias).reshape(-1) if bias is not None else None
    flags = (FLAG_BW if bw else 0) | (FLAG_BA if ba else 0)
    lib().orc_linear(_p(x), _p(W), _p(beta), _p(scale), _p(bias), _l(R), _c(K), _c(Cout), _c(flags), _p(y))
    return y


def sign_plane(x, beta):
    x = _f(x)
    beta = _f(beta).reshape(-1)
    K = x.shape[-1]
    out = np.empty(x.shape, dtype=np.int8)
    lib().orc_sign_plane(_p(x), _p(beta), _l(x.size // K), _c(K), _p(out))
    return out


def v2s(v, W, scale=None, binary=False, return_z=False):
    """sv_layers.py:111-129.  v (...,3,C) -> s (...,3C) [, z (...,3,3)]."""
    v, W = _f(v), _f(W)
    C = v.shape[-1]
    R = v.size // (3 * C)
    s = np.empty(v.shape[:-2] + (3 * C,), dtype=np.float32)
    z = np.empty(v.shape[:-2] + (3, 3), dtype=np.float32) if return_z else None
    scale = _f(scale).reshape(-1) if scale is not None else None
    lib().orc_v2s(_p(v), _p(W), _p(scale), _l(R), _c(C), _c(1 if binary else 0), _p(s), _p(z))
    return (s, z) if return_z else s


def bn_act(x, bn, act=ACT_NONE):
    """nn.BatchNorm1d eval on the last axis + activation."""
    x = _f(x)
    C = x.shape[-1]
    y = np.empty_like(x)
    w, b, m, var = (_f(t) for t in bn)
    lib().orc_bn_act(_p(x), _p(w), _p(b), _p(m), _p(var), ctypes.c_float(BN_EPS), _l(x.size // C), _c(C),
                     _c(act), _p(y))
    return y


def vector_bn(v, bn):
    """sv_layers.py:86-102."""
    v = _f(v)
    C = v.shape[-1]
    out = np.empty_like(v)
    w, b, m, var = (_f(t) for t in bn)
    lib().orc_vector_bn(_p(v), _p(w), _p(b), _p(m), _p(var), ctypes.c_float(BN_EPS), _l(v.size // (3 * C)),
                        _c(C), _p(out))
    return out


def gate(s, G1, G2):
    """sv_layers.py:179-183.  s (B, ..., Cs) -> (B, Co)."""
    s, G1, G2 = _f(s), _f(G1), _f(G2)
    B, Cs = s.shape[0], s.shape[-1]
    rows = s.size // (B * Cs)
    H, Co = G1.shape[0], G2.shape[0]
    g = np.empty((B, Co), dtype=np.float32)
    lib().orc_gate(_p(s), _p(G1), _p(G2), _c(B), _l(rows), _c(Cs), _c(H), _c(Co), _p(g))
    return g


def scale_v(v, g):
    v, g = _f(v), _f(g)
    B, C = g.shape
    rows = v.size // (B * 3 * C)
    out = np.empty_like(v)
    lib().orc_scale_v(_p(v), _p(g), _c(B), _l(rows), _c(C), _p(out))
    return out


def _pool(fn, x, axis):
    x = _f(x)
    axis = axis % x.ndim
    A = int(np.prod(x.shape[:axis], dtype=np.int64))
    M = x.shape[axis]
    C = int(np.prod(x.shape[axis + 1:], dtype=np.int64))
    out = np.empty(x.shape[:axis] + x.shape[axis + 1:], dtype=np.float32)
    fn(_p(x), _l(A), _l(M), _l(C), _p(out))
    return out


def pool_max(x, axis):
    return _pool(lib().orc_pool_max, x, axis)


def pool_mean(x, axis):
    return _pool(lib().orc_pool_mean, x, axis)


def svpool(sv, dim=2, keepdim=False, spool="max"):
    """sv_util.py:118-132."""
    s, v = sv
    if spool == "max":
        s2 = pool_max(s, dim)
    elif spool == "mean":
        s2 = pool_mean(s, dim)
    else:
        raise ValueError("not recognized pooling mean {}".format(spool))
    v2 = pool_mean(v, dim)
    if keepdim:
        s2, v2 = np.expand_dims(s2, dim), np.expand_dims(v2, dim)
    return s2, v2


def svcat(xs):
    """sv_util.py:134-144."""
    return (np.concatenate([x[0] for x in xs], axis=-1), np.concatenate([x[1] for x in xs], axis=-1))


# ------------------------------------------------------------------------------------------------
# parameter access
# ------------------------------------------------------------------------------------------------
class Params:
    """state_dict view (torch tensors or numpy arrays; 'module.' prefix tolerated)."""

    def __init__(self, state_dict, prefix=""):
        self.sd = state_dict
        self.prefix = prefix

    def sub(self, name):
        return Params(self.sd, self.prefix + name + ".")

    def has(self, name):
        return (self.prefix + name) in self.sd or ("module." + self.prefix + name) in self.sd

    def get(self, name):
        key = self.prefix + name
        t = self.sd[key] if key in self.sd else self.sd["module." + key]
        if hasattr(t, "detach"):
            t = t.detach().cpu().numpy()
        return np.ascontiguousarray(t, dtype=np.float32)

    def opt(self, name):
        return self.get(name) if self.has(name) else None

    def bn(self, name):
        q = self.sub(name)
        return (q.get("weight"), q.get("bias"), q.get("running_mean"), q.get("running_var"))


def p_linear(p, x, ba=None):
    """sv_layers.Linear with flags inferred from the keys present (scale <=> bw, beta <=> ba)."""
    W = p.get("weight")
    if W.ndim == 3:
        W = W[:, :, 0]
    scale, beta = p.opt("scale"), p.opt("beta")
    return linear(x, W, beta=beta, scale=scale, bias=p.opt("bias"), bw=scale is not None,
                  ba=beta is not None)


def p_v2s(p, v, return_z=False):
    q = p.sub("linear")
    scale = q.opt("scale")
    return v2s(v, q.get("weight"), scale=scale, binary=scale is not None, return_z=return_z)


def svblock(p, sv):
    """SVBlock.forward, sv_layers.py:172-196."""
    s, v = sv
    g = gate(s, p.get("gate.0.weight"), p.get("gate.2.weight"))            # :179-183
    s_v = p_v2s(p.sub("v2s"), v)                                             # :185
    u = np.concatenate([s, s_v], axis=-1)                                    # :186
    y = p_linear(p.sub("linear1"), u)                                        # :187
    s_out = bn_act(y, p.bn("bn1"), ACT_LEAKY)                                # :188-190
    w = p_linear(p.sub("linear2"), v)                                        # :192
    w = vector_bn(w, p.bn("bn2.bn"))                                         # :193
    v_out = scale_v(w, g)                                                    # :194
    return s_out, v_out


def svfuse(p, sv, trans_back=False):
    """SVFuse.forward, sv_layers.py:206-220."""
    s, v = sv
    if trans_back:
        s_v, z = p_v2s(p.sub("v2s"), v, return_z=True)
        return np.concatenate([s, s_v], axis=-1), z
    return np.concatenate([s, p_v2s(p.sub("v2s"), v)], axis=-1)


def sv_stnkd(p, sv):
    """SV_STNkd.forward, sv_layers.py:234-244."""
    x = svblock(p.sub("conv1"), sv)
    x = svblock(p.sub("conv2"), x)
    x = svblock(p.sub("conv3"), x)
    x = svpool(x, dim=1)
    x = svblock(p.sub("fc1"), x)
    x = svblock(p.sub("fc2"), x)
    return svblock(p.sub("fc3"), x)


def _expand(x, n):
    return np.ascontiguousarray(np.broadcast_to(np.expand_dims(x, 1), (x.shape[0], n) + x.shape[1:]))


def edge_features_sv(sv, k, idx=None):
    """get_graph_feature_sv, sv_util.py:90-116 (kNN on cat[s, v.flat], :100-101)."""
    s, v = sv
    B, N = s.shape[:2]
    if idx is None:
        idx = knn(np.concatenate([s, v.reshape(B, N, -1)], axis=-1), k)
    return graph_feature_sv(s, v, idx), idx


# ------------------------------------------------------------------------------------------------
# models
# ------------------------------------------------------------------------------------------------
def _dgcnn_trunk(P, x, k, forced_idx, rec):
    """Shared first half of SV_DGCNN_CLS / SV_DGCNN_PSEG forward (sv_dgcnn_cls.py:47-67,
    sv_dgcnn_partseg.py:81-103)."""
    xyz = np.ascontiguousarray(np.transpose(_f(x), (0, 2, 1)))             # (B,N,3)
    idx = forced_idx[0] if forced_idx else knn(xyz, k)
    rec["idx"].append(idx)
    v = graph_feature_xyz(xyz, idx, 2)
    s = p_v2s(P.sub("init_scalar"), v)
    cur = svpool(svblock(P.sub("conv1"), (s, v)))
    rec["pools"].append(cur)
    outs = [cur]
    for li, name in enumerate(("conv2", "conv3", "conv4")):
        fi = forced_idx[li + 1] if forced_idx else None
        if rec.get("teacher"):
            cur = rec["teacher"][li]
        e, idx = edge_features_sv(cur, k, fi)
        rec["idx"].append(idx)
        cur = svpool(svblock(P.sub(name), e))
        rec["pools"].append(cur)
        outs.append(cur)
    return outs


def sv_dgcnn_cls(state_dict, x, k, forced_idx=None, rec=None):
    """SV_DGCNN_CLS.forward, sv_dgcnn_cls.py:46-82.  x (B,3,N) -> logits (B,num_class)."""
    P = Params(state_dict)
    rec = rec if rec is not None else {}
    rec.setdefault("idx", []); rec.setdefault("pools", [])
    outs = _dgcnn_trunk(P, x, k, forced_idx, rec)
    y = svblock(P.sub("conv5"), svcat(outs))                                 # :67-68
    y = svfuse(P.sub("svfuse"), y)                                           # :69
    rec["fused"] = y
    g = np.concatenate([pool_max(y, 1), pool_mean(y, 1)], axis=-1)           # :70-74
    rec["global"] = g
    h = bn_act(p_linear(P.sub("linear1"), g), P.bn("bn1"), ACT_LEAKY)        # :76
    rec["h1"] = h
    h = bn_act(p_linear(P.sub("linear2"), h), P.bn("bn2"), ACT_LEAKY)        # :78
    rec["h2"] = h
    return p_linear(P.sub("linear3"), h)                                     # :80


def _conv_bn_act(P, name, x, act):
    """nn.Sequential(Conv1d(kernel 1), BatchNorm1d, act) on row-major (rows, C) data."""
    y = p_linear(P.sub(name + ".0"), x)
    return bn_act(y, P.bn(name + ".1"), act)


def sv_dgcnn_pseg(state_dict, x, label, k, forced_idx=None, rec=None):
    """SV_DGCNN_PSEG.forward, sv_dgcnn_partseg.py:80-128.  -> (B,num_part,N)."""
    P = Params(state_dict)
    rec = rec if rec is not None else {}
    rec.setdefault("idx", []); rec.setdefault("pools", [])
    outs = _dgcnn_trunk(P, x, k, forced_idx, rec)
    B, N = outs[0][0].shape[:2]
    cat = svcat(outs)
    x_fine = svfuse(P.sub("svfuse1"), cat)                                   # :104
    y = svblock(P.sub("conv5"), cat)                                         # :106
    x_pool = svpool(y, dim=1, keepdim=True)                                  # :107
    x_pool = svblock(P.sub("conv6"), x_pool)                                 # :108
    x_pool = svfuse(P.sub("svfuse2"), x_pool)                                # :109  (B,1,C)
    y = svfuse(P.sub("svfuse3"), y)                                          # :111
    y = pool_max(y, 1)                                                       # :112  (B,emb)
    l = _conv_bn_act(P, "conv7", _f(label).reshape(B, -1), ACT_LEAKY)        # :114-115
    glob = np.concatenate([y, x_pool[:, 0, :], l], axis=-1)                  # :117
    rec["glob"] = glob
    rec["x_fine"] = x_fine
    h = np.concatenate([_expand(glob, N), x_fine], axis=-1)                  # :118-120 (B,N,C)
    h = _conv_bn_act(P, "conv8", h, ACT_LEAKY)                               # :121
    h = _conv_bn_act(P, "conv9", h, ACT_LEAKY)                               # :123
    h = _conv_bn_act(P, "conv10", h, ACT_LEAKY)                              # :125
    out = p_linear(P.sub("conv11"), h)                                       # :126
    return np.ascontiguousarray(np.transpose(out, (0, 2, 1)))


def _pointnet_front(P, x, k, forced_idx, rec):
    """First edge layer of both SV-PointNets (sv_pointnet_cls.py:32-40)."""
    xyz = np.ascontiguousarray(np.transpose(_f(x), (0, 2, 1)))
    idx = forced_idx[0] if forced_idx else knn(xyz, k)
    rec["idx"].append(idx)
    v = graph_feature_xyz(xyz, idx, 3)
    s = p_v2s(P.sub("init_scalar"), v)
    cur = svpool(svblock(P.sub("conv_pos"), (s, v)))
    rec["pools"].append(cur)
    return cur


def sv_pointnet_encoder(P, x, k, forced_idx=None, rec=None):
    """SVPointNetEncoder.forward, sv_pointnet_cls.py:31-58."""
    cur = _pointnet_front(P, x, k, forced_idx, rec)
    N = cur[0].shape[1]
    cur = svblock(P.sub("conv1"), cur)                                       # :42
    g = sv_stnkd(P.sub("fstn"), cur)                                         # :44
    cur = svcat([cur, (_expand(g[0], N), _expand(g[1], N))])                 # :45-46
    cur = svblock(P.sub("conv2"), cur)
    cur = svblock(P.sub("conv3"), cur)                                       # :48-49
    m = svpool(cur, dim=1)                                                   # :51
    cur = svcat([cur, (_expand(m[0], N), _expand(m[1], N))])                 # :52-53
    cur = svblock(P.sub("conv_fuse"), cur)                                   # :54
    cur = svpool(cur, dim=1)                                                 # :56
    return svfuse(P.sub("svfuse"), cur)                                      # :57


def sv_pointnet_cls(state_dict, x, k, forced_idx=None, rec=None):
    """SV_PointNet_CLS.forward, sv_pointnet_cls.py:76-81 (ReLU; dropout is identity in eval)."""
    P = Params(state_dict)
    rec = rec if rec is not None else {}
    rec.setdefault("idx", []); rec.setdefault("pools", [])
    f = sv_pointnet_encoder(P.sub("feat"), x, k, forced_idx, rec)
    rec["feat"] = f
    h = bn_act(p_linear(P.sub("fc1"), f), P.bn("bn1"), ACT_RELU)
    h = bn_act(p_linear(P.sub("fc2"), h), P.bn("bn2"), ACT_RELU)
    return p_linear(P.sub("fc3"), h)


def sv_pointnet_pseg(state_dict, x, label, k, forced_idx=None, rec=None):
    """SV_PointNet_PSEG.forward, sv_pointnet_partseg.py:55-97."""
    P = Params(state_dict)
    rec = rec if rec is not None else {}
    rec.setdefault("idx", []); rec.setdefault("pools", [])
    binary = P.has("conv1.linear1.beta")
    cur = _pointnet_front(P, x, k, forced_idx, rec)
    B, N = cur[0].shape[:2]
    out1 = svblock(P.sub("conv1"), cur)
    out2 = svblock(P.sub("conv2"), out1)
    out3 = svblock(P.sub("conv3"), out2)                                     # :67-69
    g = sv_stnkd(P.sub("fstn"), out3)                                        # :71
    xt = svcat([out3, (_expand(g[0], N), _expand(g[1], N))])                 # :72-73
    out4 = svblock(P.sub("conv4"), xt)
    out5 = svblock(P.sub("conv5"), out4)                                     # :74-75
    m = svpool(out5, dim=1, spool="mean")                                    # :77
    y = svcat([out5, (_expand(m[0], N), _expand(m[1], N))])                  # :78-79
    y, trans = svfuse(P.sub("svfuse"), y, trans_back=True)                   # :80  y (B,N,C) trans (B,N,3,3)
    y = _conv_bn_act(P, "conv_fuse1", y, ACT_RELU)                           # :82
    y = _conv_bn_act(P, "conv_fuse2", y, ACT_RELU)                           # :83
    y = pool_mean(y, 1) if binary else pool_max(y, 1)                        # :84-87
    x_l = np.concatenate([y, _f(label).reshape(B, -1)], axis=-1)             # :89
    cs, cv = svcat([out1, out2, out3, out4, out5])                           # :92
    # einsum('bimj,bijk->bimk', v^T (B,N,C,3), trans (B,N,3,3)) -> (B,N,C,3): row c = v[:,c] . trans
    C = cv.shape[-1]
    vt = np.ascontiguousarray(np.transpose(cv, (0, 1, 3, 2)))                # (B,N,C,3)
    cvt = np.empty((B, N, C, 3), dtype=np.float32)
    for kk in range(3):
        acc = vt[..., 0] * trans[:, :, None, 0, kk]
        acc = acc + vt[..., 1] * trans[:, :, None, 1, kk]
        acc = acc + vt[..., 2] * trans[:, :, None, 2, kk]
        cvt[..., kk] = acc
    concat = np.concatenate([_expand(x_l, N), cs, cvt.reshape(B, N, -1)], axis=-1)   # :94-95
    rec["concat"] = concat
    h = _conv_bn_act(P, "convs1", concat, ACT_RELU)
    h = _conv_bn_act(P, "convs2", h, ACT_RELU)
    h = _conv_bn_act(P, "convs3", h, ACT_RELU)
    out = p_linear(P.sub("convs4"), h)
    return np.ascontiguousarray(np.transpose(out, (0, 2, 1)))
