"""GPU parity tests: the CUDA path (through the C ABI, via the nn.Module mirror) against the CPU
oracle and the reference-generated golden fixtures.

Bars (north_star): kNN indices and packed sign bits bit-exact; floats within 1e-3 rel / 1e-4 abs.
Binary models are chaotic end to end, so layers are checked teacher-forced: each
layer gets the oracle's/golden inputs and indices and must reproduce that layer exactly.
"""
import numpy as np
import pytest
import torch

from oracle import svnet_oracle as orc
from tests.util import assert_close, golden, golden_state_dict, knn_is_valid, quiet, t2n
from svnet_b200.synthetic import make_args, synthetic_clouds, synthetic_state_dict

pytestmark = pytest.mark.gpu

DEV = "cuda"


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).numpy()


# ------------------------------------------------------------------------------------------------
# kNN
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N,C,k", [(2, 96, 3, 12), (1, 200, 3, 1), (2, 130, 62, 20), (1, 257, 127, 40),
                                     (1, 300, 136, 64), (1, 150, 20, 100), (3, 64, 5, 64), (1, 33, 3, 33)])
def test_knn_bit_exact_vs_oracle(B, N, C, k):
    import svnet_b200 as sv
    feat = rnd((B, N, C), 100 + N + C)
    ref = orc.knn(feat, k)
    x = cu(feat).transpose(1, 2)  # (B,C,N) view, like the reference call site (sv_util.py:101)
    idx = sv.knn(x, k)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (B, N, k)
    assert (t2n(idx) == ref).all()


def test_knn_sv_view_and_errors():
    import svnet_b200 as sv
    from svnet_b200 import _native as nv
    B, N, Cs, Cv, k = 2, 120, 32, 10, 20
    s, v = rnd((B, N, Cs), 1), rnd((B, N, 3, Cv), 2)
    ref = orc.knn(np.concatenate([s, v.reshape(B, N, -1)], -1), k)
    # strided slices of a wider table, like the svcat layout of the fused path
    S = torch.zeros((B * N, 100), device=DEV)
    V = torch.zeros((B * N, 3, 37), device=DEV)
    S[:, 7:7 + Cs] = cu(s).view(B * N, Cs)
    V[:, :, 5:5 + Cv] = cu(v).view(B * N, 3, Cv)
    i32, i64 = nv.knn(nv.view_of(S[:, 7:7 + Cs], V[:, :, 5:5 + Cv]), B, N, k, want64=True)
    assert (t2n(i32) == ref).all() and (t2n(i64) == ref).all()
    with pytest.raises(RuntimeError):
        sv.knn(cu(rnd((1, 3, 10), 3)), 11)  # k > N: torch.topk raises as well


def test_knn_ties_lowest_index():
    import svnet_b200 as sv
    g = golden("knn")
    for name in ("dup", "lat"):
        x, k = g[name + "_x"], int(g[name + "_k"])
        ref, pd = orc.knn(np.transpose(x, (0, 2, 1)), k, return_pd=True)
        idx = t2n(sv.knn(cu(x), k))
        assert (idx == ref).all()
        assert knn_is_valid(idx, pd, k)


def test_knn_golden_reference_indices():
    import svnet_b200 as sv
    g = golden("knn")
    for name in ("xyz", "c62", "c127"):
        idx = t2n(sv.knn(cu(g[name + "_x"]), int(g[name + "_k"])))
        assert (idx == g[name + "_idx"]).all(), name


def test_knn_full_size_properties():
    """BASELINE sizes (N=1024,k=20 and N=2048,k=40): self is rank 0, scores non-increasing along k,
    every unselected score <= the k-th selected (checked with fp64 distances + a few-ulp slack),
    and exact equality with the oracle on the first clouds."""
    import svnet_b200 as sv
    for (B, N, C, k) in [(8, 1024, 62, 20), (2, 2048, 136, 40)]:
        feat = rnd((B, N, C), 7 + N, scale=0.5)
        idx = sv.knn(cu(feat).transpose(1, 2), k)
        ref = orc.knn(feat[:2], k)
        assert (t2n(idx[:2]) == ref).all()
        f = cu(feat).double()
        d2 = torch.cdist(f, f) ** 2
        sel = torch.gather(d2, 2, idx)
        assert (idx[:, :, 0] == torch.arange(N, device=DEV).view(1, N)).all()
        scale = float(d2.max())
        assert (sel[:, :, 1:] - sel[:, :, :-1] >= -1e-5 * scale).all()
        rest = d2.clone()
        rest.scatter_(2, idx, float("inf"))
        assert (rest.min(dim=2)[0] >= sel[:, :, -1] - 1e-5 * scale).all()


def _tc_stats(reset=True):
    import ctypes
    from svnet_b200 import _native as nv
    out = (ctypes.c_ulonglong * 12)()
    assert nv.lib().svnet_knn_tc_stats(out, ctypes.c_int(1 if reset else 0)) == 0
    return list(out)


@pytest.mark.parametrize("case", ["duplicates", "heavy_ties", "lattice", "offset", "all_equal", "ragged", "two_clusters",
                                  "k40", "k40_ties", "k32"])
def test_knn_tensor_core_path_adversarial(case, monkeypatch):
    """Shapes served by the tcgen05 filter (csrc/knn_tc.cu: N >= 64, k <= 48): exact ties (duplicated
    points, lattice), norms much larger than neighbour gaps, everything equal (queue overflow ->
    brute-force rows), N not a multiple of the 128/256 tiles.  Indices must equal the oracle's, and
    the CUDA-core kernel's, bit for bit; the counters prove that the tensor-core path ran."""
    import svnet_b200 as sv
    from svnet_b200 import _native as nv
    g = torch.Generator().manual_seed(17)
    B, N, C, k = 2, 384, 62, 20
    if case == "duplicates":
        base = torch.randn((B, N // 4, C), generator=g)
        feat = base.repeat(1, 4, 1)[:, torch.randperm(N, generator=g)]
    elif case == "heavy_ties":
        B, N, C, k = 1, 640, 62, 20          # 16 distinct points x 40 copies: > 32 exact ties around every query
        base = torch.randn((B, 16, C), generator=g)
        feat = base.repeat(1, 40, 1)[:, torch.randperm(N, generator=g)]
    elif case == "k40":                      # part-segmentation shape: 32 < k <= 48 -> wide finish kernel
        B, N, C, k = 2, 700, 80, 40
        feat = torch.randn((B, N, C), generator=g) * 0.2 + 0.5
    elif case == "k40_ties":
        B, N, C, k = 1, 512, 16, 40
        base = torch.randn((B, 64, C), generator=g)
        feat = base.repeat(1, 8, 1)[:, torch.randperm(N, generator=g)]
    elif case == "k32":
        B, N, C, k = 1, 400, 62, 32
        feat = torch.randn((B, N, C), generator=g)
    elif case == "lattice":
        B, N, C, k = 1, 343, 3, 20
        ax = torch.arange(7, dtype=torch.float32)
        feat = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(1, N, 3) * 0.25
    elif case == "offset":
        feat = torch.randn((B, N, C), generator=g) * 0.05 + 3.0
    elif case == "all_equal":
        B, N, C, k = 1, 200, 5, 8
        feat = torch.ones((B, N, C)) * 0.5
    elif case == "ragged":
        B, N, C, k = 3, 333, 127, 24
        feat = torch.randn((B, N, C), generator=g)
    else:
        B, N, C, k = 1, 1024, 30, 20
        feat = torch.cat([torch.randn((B, 1000, C), generator=g) * 0.01, torch.randn((B, 24, C), generator=g) + 5.0], 1)
    feat = feat.contiguous()
    ref = orc.knn(feat.numpy(), k)
    x = feat.to(DEV).view(B * N, C)
    monkeypatch.setenv("SVNET_KNN_TC_STATS", "1")
    _tc_stats()
    monkeypatch.setenv("SVNET_KNN_TC", "1")
    tc = t2n(nv.knn(nv.view_of(x, None), B, N, k)[0])
    st = _tc_stats()
    monkeypatch.setenv("SVNET_KNN_TC", "0")
    cc = t2n(nv.knn(nv.view_of(x, None), B, N, k)[0])
    assert st[0] == B * N, "tensor-core path did not run"
    assert (cc == ref).all()
    assert (tc == ref).all(), "%d rows differ" % int((tc != ref).any(-1).sum())
    if case == "all_equal":
        assert st[2] > 0      # every candidate reaches the threshold: the rows take the brute-force path
    if case == "heavy_ties":
        assert st[10] > 0 and st[2] == 0      # more than 32 survivors: every survivor re-scored exactly, no brute force


def test_knn_tensor_core_error_model_on_model_features():
    """The filter's error bound eps(C) (knn_tc_eps) must dominate the measured |tensor-core score -
    oracle chain score| / (xx_i + xx_j) with margin on real layer features (binary SV-DGCNN)."""
    import struct
    import svnet_b200 as sv
    import os
    net = quiet(sv.SV_DGCNN_CLS, make_args(k=20, binary=True), 40)
    net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=1002))
    net = net.to(DEV).eval()
    x = synthetic_clouds(8, 1024, 1002).to(DEV)
    os.environ["SVNET_KNN_TC_STATS"] = "1"
    try:
        _tc_stats()
        with torch.no_grad():
            net(x)
        torch.cuda.synchronize()
        st = _tc_stats()
    finally:
        os.environ["SVNET_KNN_TC_STATS"] = "0"
    assert st[0] == 4 * 8 * 1024 and st[2] == 0
    err = struct.unpack("f", struct.pack("I", st[9] & 0xffffffff))[0]
    eps_c127 = 1.5 * (8 * 7.2e-7 + 127 * 3.0e-8 + 8.0e-7)
    assert 0.0 < err < eps_c127 / 1.5, err


def test_rows_prep_fast_bits_equal_generic_and_oracle():
    """csrc/rows_fast.cu (sign words only, conv5 shape) against the generic kernel of csrc/rows.cu (forced by
    also requesting u) and the oracle's v2s + sign planes: words must be bit-identical."""
    from svnet_b200 import _native as nv
    rows, Cs, Cv = 3001, 256, 83
    s, v = rnd((rows, Cs), 5), rnd((rows, 3, Cv), 6)
    Wz, beta = rnd((3, Cv), 7, 0.3), rnd((Cs + 3 * Cv,), 8, 0.05)
    beta[::7] = 0.0
    s[::5, ::3] = 0.0                                   # exact zeros: sign(0) = 0 -> mask plane
    S, V = torch.zeros((rows, Cs + 9), device=DEV), torch.zeros((rows, 3, Cv + 4), device=DEV)
    S[:, 4:4 + Cs], V[:, :, 2:2 + Cv] = cu(s), cu(v)
    view = nv.view_of(S[:, 4:4 + Cs], V[:, :, 2:2 + Cv])
    fast = nv.rows_prep(view, rows, Wz=cu(Wz), beta=cu(beta), want_bits=True)
    u = torch.empty((rows, Cs + 3 * Cv), device=DEV)
    slow = nv.rows_prep(view, rows, Wz=cu(Wz), beta=cu(beta), want_bits=True, u_out=u, ldu=u.stride(0))
    for a, b in zip(fast, slow):
        assert (a == b).all()
    q = orc.v2s(v, Wz)
    t = orc.sign_plane(np.concatenate([s, q], -1), beta)
    pos, nz = t > 0, t != 0
    K = Cs + 3 * Cv
    bits = np.unpackbits(t2n(fast[0]).view(np.uint8).reshape(rows, -1), axis=1, bitorder="little")[:, :K]
    msk = np.unpackbits(t2n(fast[1]).view(np.uint8).reshape(rows, -1), axis=1, bitorder="little")[:, :K]
    assert (bits == pos.astype(np.uint8)).all() and (msk == nz.astype(np.uint8)).all()


@pytest.mark.parametrize("rows,K,Cout", [(3000, 505, 512), (2048, 256, 128), (4100, 160, 40), (2500, 640, 300)])
def test_binlinear_tensor_core_equals_popcount(rows, K, Cout, monkeypatch):
    """csrc/binlinear_tc.cu (bf16 UMMAs, exact integers) against the XNOR/popcount kernel: integer dots and the
    float epilogue (scale, bias, BN, LeakyReLU, per-cloud constant part) must be bit-identical."""
    from svnet_b200 import _native as nv
    x, W = rnd((rows, K), 21), rnd((Cout, K), 22)
    beta = rnd((K,), 23, 0.1)
    x[::3, ::5] = 0.0
    beta[::5] = 0.0                                    # exact zeros -> mask plane
    bits, mask, nvalid = nv.rows_prep(nv.view_of(cu(x), None), rows, beta=cu(beta), want_bits=True)
    W1b = nv.pack_sign(cu(W))
    scale, bias = cu(np.abs(rnd((Cout,), 24)) + 0.1), cu(rnd((Cout,), 25))
    bn = (cu(rnd((Cout,), 26)), cu(rnd((Cout,), 27)))
    rpc = 500
    cloud = torch.randint(-50, 50, ((rows + rpc - 1) // rpc, Cout), dtype=torch.int32, device=DEV)
    outs = {}
    for tc in ("2", "0"):            # "2": every covered call on the tensor cores (default "1": lean epilogue only)
        monkeypatch.setenv("SVNET_BINLINEAR_TC", tc)
        assert (int(nv.lib().svnet_binlinear_workspace_bytes(rows, K, Cout)) > 0) == (tc == "2")
        outs[tc] = (nv.binlinear_rows(bits, mask, nvalid, K, W1b, Cout, out_i32=True),
                    nv.binlinear_rows(bits, mask, nvalid, K, W1b, Cout, scale=scale, bias=bias, bn=bn, act=nv.ACT_LEAKY),
                    nv.binlinear_rows(bits, mask, nvalid, K, W1b, Cout, scale=scale, cloud_dot=cloud, rows_per_cloud=rpc),
                    # scale + BN + LeakyReLU without bias: the lean epilogue when rows % 128 == 0
                    nv.binlinear_rows(bits, mask, nvalid, K, W1b, Cout, scale=scale, bn=bn, act=nv.ACT_LEAKY))
    for a, b in zip(outs["2"], outs["0"]):
        assert torch.equal(a, b)
    t = orc.sign_plane(x, beta).astype(np.int32)
    ref = t @ np.sign(W).astype(np.int32).T
    assert (t2n(outs["2"][0]) == ref).all()


def test_binlinear_pool_equals_linear_then_pool():
    """svnet_binlinear_pool_ws (binarised Linear -> BN -> LeakyReLU -> per-cloud max | mean in one tensor-core
    kernel) against svnet_binlinear_rows + svnet_pool_rows: the max is exact, the mean differs only by the
    summation order."""
    from svnet_b200 import _native as nv
    B, N, K, Cout = 3, 1024, 505, 512
    rows = B * N
    x, W, beta = rnd((rows, K), 31), rnd((Cout, K), 32), rnd((K,), 33, 0.1)
    bits, mask, nvalid = nv.rows_prep(nv.view_of(cu(x), None), rows, beta=cu(beta), want_bits=True)
    W1b = nv.pack_sign(cu(W))
    scale = cu(np.abs(rnd((Cout,), 34)) * 0.05 + 0.01)
    bn = (cu(rnd((Cout,), 35)), cu(rnd((Cout,), 36)))
    y = nv.binlinear_rows(bits, mask, nvalid, K, W1b, Cout, scale=scale, bn=bn, act=nv.ACT_LEAKY)
    mx, mean = nv.pool_rows(y, Cout, Cout, B, N, want_max=True, want_mean=True)
    assert nv.binlinear_pool_workspace(rows, K, Cout, N) > 0
    g = torch.zeros((B, 2 * Cout + 8), device=DEV)
    nv.binlinear_pool(bits, mask, K, W1b, Cout, scale, bn, N, g[:, 4:4 + Cout], g[:, 8 + Cout:], g.stride(0))
    assert torch.equal(g[:, 4:4 + Cout], mx)
    assert_close(t2n(g[:, 8 + Cout:8 + 2 * Cout]), t2n(mean), rtol=1e-5, atol=1e-6, what="binlinear_pool mean")
    assert float(g[:, :4].abs().max()) == 0.0 and float(g[:, 4 + Cout:8 + Cout].abs().max()) == 0.0
    assert nv.binlinear_pool_workspace(rows, K, Cout, 1000) == 0        # rows_per_cloud % 128 != 0: not covered


@pytest.mark.parametrize("rows,K,N", [(3000, 505, 512), (2048, 62, 64), (5000, 300, 100), (4096, 32, 40), (2100, 640, 130),
                                      (32, 2044, 512), (200, 1022, 300)])
def test_linear_three_plane_tensor_core_matches_cuda_core(rows, K, N, monkeypatch):
    """csrc/gemm_tc3.cu (fp32 linear as six bf16 plane products on tcgen05) against the CUDA-core GEMM with the
    oracle's sequential chain: fp32-level agreement (the summation order differs), all epilogue options."""
    from svnet_b200 import _native as nv
    A, W = cu(rnd((rows, K + 3), 41)), cu(rnd((N, K), 42, 0.2))
    bias, scale = cu(rnd((N,), 43)), cu(np.abs(rnd((N,), 44)) + 0.5)
    bn = (cu(rnd((N,), 45)), cu(rnd((N,), 46)))
    outs = {}
    for tc in ("1", "0"):
        monkeypatch.setenv("SVNET_LINEAR_TC", tc)
        res = []
        for kw in (dict(), dict(bias=bias, bn=bn, act=nv.ACT_RELU), dict(colscale=scale, bn=bn, act=nv.ACT_LEAKY)):
            C = torch.full((rows, N + 5), 7.0, device=DEV)
            nv.linear_rows(A, A.stride(0), 0, 1, rows, K, W, N, C, C.stride(0), 0, **kw)
            res.append(C)
        outs[tc] = res
    ref = t2n(A[:, :K]).astype(np.float64) @ t2n(W).astype(np.float64).T
    for a, b in zip(outs["1"], outs["0"]):
        assert torch.equal(a[:, N:], b[:, N:])                      # columns beyond N untouched
        assert_close(t2n(a[:, :N]), t2n(b[:, :N]), rtol=1e-4, atol=1e-4, what="tc3 vs cuda-core")
    # accuracy against fp64: the tensor cores truncate when they accumulate, so the error grows with the number
    # of accumulation steps (6 per 16 channels) -- about 1e-5 of the row/column norms, 10x the fp32 chain
    scale = np.linalg.norm(t2n(A[:, :K]), axis=1, keepdims=True) * np.linalg.norm(t2n(W), axis=1)[None, :]
    err_tc = (np.abs(t2n(outs["1"][0][:, :N]) - ref) / scale).max()
    err_cc = (np.abs(t2n(outs["0"][0][:, :N]) - ref) / scale).max()
    print("relative error vs fp64: tensor cores %.2e, fp32 chain %.2e" % (err_tc, err_cc))
    assert err_tc < 2e-6, (err_tc, err_cc)


def test_svfuse_pool_equals_materialised_path():
    """svnet_svfuse_pool (v2s reduced on the fly) against rows_prep(u_out) + pool_rows: the max is exact,
    the mean differs only by the summation order."""
    from svnet_b200 import _native as nv
    B, N, Cv = 3, 1000, 170
    v = cu(rnd((B * N, 3, Cv), 9))
    Wz, zs = cu(rnd((3, Cv), 10, 0.2)), cu(np.abs(rnd((3,), 11)) + 0.5)
    K = 3 * Cv
    u = torch.empty((B * N, K), device=DEV)
    nv.rows_prep(nv.view_of(None, v), B * N, Wz=Wz, zscale=zs, u_out=u, ldu=K)
    mx, mean = nv.pool_rows(u, K, K, B, N, want_max=True, want_mean=True)
    g = torch.zeros((B, 2 * K + 6), device=DEV)
    nv.svfuse_pool(v, B, N, Wz, zs, g[:, 3:3 + K], g[:, 3 + K + 3:], g.stride(0))
    assert (g[:, 3:3 + K] == mx).all()
    assert_close(t2n(g[:, 6 + K:6 + 2 * K]), t2n(mean), rtol=1e-5, atol=1e-6, what="svfuse_pool mean")
    assert float(g[:, :3].abs().max()) == 0.0 and float(g[:, 3 + K:6 + K].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------
# module-level API vs golden (reference outputs) and oracle
# ------------------------------------------------------------------------------------------------
def _load(module, tmpl_seed, beta_zero=False):
    sd = synthetic_state_dict(module.state_dict(), seed=tmpl_seed, beta_zero=beta_zero)
    module.load_state_dict(sd)
    return module.to(DEV).eval(), sd


def test_graph_features_and_pool():
    import svnet_b200 as sv
    g = golden("graph_features")
    x = cu(g["x"])
    k = int(g["k"])
    assert (t2n(sv.get_graph_feature(x, k=k)) == g["gf"]).all()
    assert (t2n(sv.get_graph_feature(x, k=k, idx=cu(g["idx"]))) == g["gf"]).all()
    gfc = t2n(sv.get_graph_feature_cross(x, k=k))
    assert (gfc == orc.graph_feature_xyz(np.transpose(g["x"][:, 0], (0, 2, 1)), g["idx"], 3)).all()
    assert_close(gfc, g["gf_cross"], rtol=1e-6, atol=1e-6)
    sf, vf = sv.get_graph_feature_sv((cu(g["s"]), cu(g["v"])), k=k)
    assert (t2n(sf) == g["sf"]).all() and (t2n(vf) == g["vf"]).all()
    ps, pv = sv.svpool((sf, vf))
    assert (t2n(ps) == g["pool_s"]).all()
    assert_close(t2n(pv), g["pool_v"], rtol=1e-5, atol=1e-6)
    ps, pv = sv.svpool((sf, vf), spool="mean")
    assert_close(t2n(ps), g["pool_mean_s"], rtol=1e-5, atol=1e-6)
    ps1, pv1 = sv.svpool((ps, pv), dim=1, keepdim=True)
    assert_close(t2n(ps1), g["pool1_s"], rtol=1e-5, atol=1e-6)
    assert_close(t2n(pv1), g["pool1_v"], rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        sv.svpool((sf, vf), spool="median")
    s2, v2 = sv.svcat([(ps, pv), (ps, pv)])
    assert s2.shape[-1] == 2 * ps.shape[-1] and v2.shape[-1] == 2 * pv.shape[-1]


def test_layers_vs_golden():
    import svnet_b200 as sv
    g = golden("layers")
    for tag, bz in (("lin_bin", False), ("lin_bin_b0", True)):
        lin, sd = _load(sv.Linear(70, 24, bias=False, bw=True, ba=True), 31, bz)
        y = t2n(lin(cu(g[tag + "_x"])))
        # integer-exact dot products: bit-identical to the oracle, tolerance-identical to torch
        yo = orc.linear(g[tag + "_x"], sd["weight"].numpy(), beta=sd["beta"].numpy(), scale=sd["scale"].numpy(),
                        bw=True, ba=True)
        assert (y == yo).all(), tag
        assert_close(y, g[tag + "_y"], rtol=1e-6, atol=1e-6, what=tag)
    lin, sd = _load(sv.Linear(42, 21, bias=False, bw=True), 33)
    assert_close(t2n(lin(cu(g["lin_bw_x"]))), g["lin_bw_y"], rtol=1e-5, atol=1e-6)
    conv, sd = _load(quiet(sv.Conv1d, 45, 16, binary=True), 35)
    assert_close(t2n(conv(cu(g["conv_bin_x"]))), g["conv_bin_y"], rtol=1e-6, atol=1e-6)
    vbn, sd = _load(sv.VectorBN(11), 37)
    assert_close(t2n(vbn(cu(g["vbn_x"]))), g["vbn_y"], rtol=1e-5, atol=1e-6)
    for tag, binary in (("v2s_fp", False), ("v2s_bin", True)):
        m, sd = _load(sv.Vector2Scalar(20, 3, binary=binary, trans_back=True), 39)
        s, z = m(cu(g[tag + "_x"]))
        so, zo = orc.v2s(g[tag + "_x"], sd["linear.weight"].numpy(),
                         scale=sd["linear.scale"].numpy() if binary else None, binary=binary, return_z=True)
        assert (t2n(s) == so).all() and (t2n(z) == zo).all(), tag  # sequential chains: bit-exact vs oracle
        assert_close(t2n(s), g[tag + "_s"], rtol=1e-5, atol=1e-5)
        assert_close(t2n(z), g[tag + "_z"], rtol=1e-5, atol=1e-6)
    for tag, binary in (("svb_fp", False), ("svb_bin", True)):
        blk, sd = _load(quiet(sv.SVBlock, (64, 20), (32, 10), binary), 41)
        so, vo = blk((cu(g[tag + "_s"]), cu(g[tag + "_v"])))
        assert_close(t2n(so), g[tag + "_so"], rtol=1e-4, atol=1e-5, what=tag + " s")
        assert_close(t2n(vo), g[tag + "_vo"], rtol=1e-4, atol=1e-5, what=tag + " v")
        so, vo = blk((cu(g[tag + "_s"][:, :, 0]), cu(g[tag + "_v"][:, :, 0])))
        assert_close(t2n(so), g[tag + "_pt_so"], rtol=1e-4, atol=1e-5)
        assert_close(t2n(vo), g[tag + "_pt_vo"], rtol=1e-4, atol=1e-5)
    fuse, sd = _load(quiet(sv.SVFuse, 10, 3, True), 44)
    assert_close(t2n(fuse((cu(g["fuse_s"]), cu(g["fuse_v"])))), g["fuse_y"], rtol=1e-5, atol=1e-5)
    stn, sd = _load(quiet(sv.SV_STNkd, (32, 10), True), 47)
    so, vo = stn((cu(g["stn_s"]), cu(g["stn_v"])))
    assert_close(t2n(so), g["stn_so"], rtol=1e-3, atol=1e-4)
    assert_close(t2n(vo), g["stn_vo"], rtol=1e-3, atol=1e-4)


def test_training_mode_raises():
    import svnet_b200 as sv
    lin = sv.Linear(8, 4, bias=False, bw=True, ba=True).to(DEV)
    lin.train()
    with pytest.raises(RuntimeError):
        lin(torch.zeros(2, 8, device=DEV))


def test_cpu_tensor_raises():
    import svnet_b200 as sv
    lin = sv.Linear(8, 4, bias=False, bw=True, ba=True).eval()
    with pytest.raises(RuntimeError):
        lin(torch.zeros(2, 8))


# ------------------------------------------------------------------------------------------------
# fused SV-DGCNN layers, teacher-forced against the oracle
# ------------------------------------------------------------------------------------------------
def _unpack(words, K):
    """(rows, Kw) int32 -> (rows, K) uint8 bits"""
    w = words.view(np.uint32)
    bits = ((w[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).astype(np.uint8)
    return bits.reshape(w.shape[0], -1)[:, :K]


@pytest.mark.parametrize("name", ["dgcnn_cls_bin", "dgcnn_cls_bin_b0", "dgcnn_cls_fp"])
def test_dgcnn_cls_layers_teacher_forced(name):
    import svnet_b200 as sv
    g = golden(name)
    sd = golden_state_dict(g)
    k, binary = int(g["k"]), bool(g["binary"])
    net = quiet(sv.SV_DGCNN_CLS, make_args(k=k, binary=binary), int(g["ncls"]))
    net.load_state_dict(sd)
    net = net.to(DEV).eval()
    x = cu(g["x"])
    B, _, N = x.shape
    gidx = [g["idx%d" % i] for i in range(4)]
    teacher = [(cu(g["pool%d_s" % i]).view(B * N, -1), cu(g["pool%d_v" % i]).view(B * N, 3, -1)) for i in range(3)]
    rec = {"teacher": teacher, "want_taps": True}
    with torch.no_grad():
        net(x, forced_idx=[cu(i, torch.int32) for i in gidx], record=rec)
    P = orc.Params(sd)
    so, vo = 0, 0
    xyz = np.ascontiguousarray(np.transpose(g["x"], (0, 2, 1)))
    for li, cname in enumerate(["conv1", "conv2", "conv3", "conv4"]):
        # oracle for this layer on the golden inputs
        if li == 0:
            v_e = orc.graph_feature_xyz(xyz, gidx[0], 2)
            s_e = orc.p_v2s(P.sub("init_scalar"), v_e)
        else:
            (s_e, v_e) = orc.graph_feature_sv(g["pool%d_s" % (li - 1)], g["pool%d_v" % (li - 1)], gidx[li])
        o_s, o_v = orc.svpool(orc.svblock(P.sub(cname), (s_e, v_e)))
        cs, cv = o_s.shape[-1], o_v.shape[-1]
        got_s = t2n(rec["s_cat"][:, so:so + cs]).reshape(B, N, cs)
        got_v = t2n(rec["v_cat"][:, :, vo:vo + cv]).reshape(B, N, 3, cv)
        if li == 0 or binary:
            # scalar branch: sequential chains (layer 1) / integer popcounts (binary) -> bit-exact
            assert (got_s == o_s).all(), "%s %s pooled scalars differ from the oracle" % (name, cname)
        else:
            assert_close(got_s, o_s, rtol=1e-4, atol=1e-5, what=cname + " s")
        assert_close(got_v, o_v, rtol=1e-4, atol=1e-5, what=cname + " v")
        # and against the reference's recorded outputs
        assert_close(got_s, g["pool%d_s" % li], rtol=1e-4, atol=1e-5, what=cname + " s vs reference")
        assert_close(got_v, g["pool%d_v" % li], rtol=1e-4, atol=1e-5, what=cname + " v vs reference")
        if li > 0 and binary:
            u = np.concatenate([s_e, orc.p_v2s(P.sub(cname + ".v2s"), v_e)], axis=-1)
            sign = orc.sign_plane(u, P.get(cname + ".linear1.beta")).reshape(B * N * k, -1)
            K = sign.shape[1]
            taps = rec["taps%d" % li]
            assert (_unpack(t2n(taps["bits"]), K) == (sign > 0)).all(), cname + " sign bits"
            assert (_unpack(t2n(taps["mask"]), K) == (sign != 0)).all(), cname + " zero mask"
        so += cs
        vo += cv


@pytest.mark.parametrize("name", ["dgcnn_cls_bin", "dgcnn_cls_bin_b0", "dgcnn_cls_fp"])
def test_dgcnn_cls_logits(name):
    """Whole forward.  With the kNN graphs teacher-forced to the reference's, logits must meet the
    north_star tolerance and argmax must be identical; free-running, the kNN graphs must agree with
    the reference's (these small fixtures have no near-ties)."""
    import svnet_b200 as sv
    g = golden(name)
    sd = golden_state_dict(g)
    k, binary = int(g["k"]), bool(g["binary"])
    net = quiet(sv.SV_DGCNN_CLS, make_args(k=k, binary=binary), int(g["ncls"]))
    net.load_state_dict({"module." + kk: vv for kk, vv in sd.items()} if False else sd)
    net = net.to(DEV).eval()
    x = cu(g["x"])
    gidx = [cu(g["idx%d" % i], torch.int32) for i in range(4)]
    with torch.no_grad():
        y = t2n(net(x, forced_idx=gidx))
        rec = {}
        y_free = t2n(net(x, record=rec))
    assert_close(y, g["logits"], what=name + " logits (forced kNN)")
    assert (y.argmax(1) == g["logits"].argmax(1)).all()
    agree = [float((t2n(rec["idx"][i]) == g["idx%d" % i]).all(-1).mean()) for i in range(4)]
    assert min(agree) >= 0.98, agree
    if min(agree) == 1.0:
        assert_close(y_free, g["logits"], what=name + " logits (free-running)")
        assert (y_free.argmax(1) == g["logits"].argmax(1)).all()


def test_dgcnn_cls_batch_independence_and_full_size():
    """cfg2 shape (B=32 -> 4 here to keep the oracle fast, N=1024, k=20): each cloud's logits do not
    depend on its batch-mates (eval-mode property) and match the oracle run on
    the same cloud."""
    import svnet_b200 as sv
    net = quiet(sv.SV_DGCNN_CLS, make_args(k=20, binary=True), 40)
    sd = synthetic_state_dict(net.state_dict(), seed=1002)
    net.load_state_dict(sd)
    net = net.to(DEV).eval()
    x = synthetic_clouds(4, 1024, 1002).to(DEV)
    with torch.no_grad():
        y = net(x)
        y1 = net(x[1:2].contiguous())
        y_perm = net(x[[2, 0, 3, 1]].contiguous())
    assert torch.equal(y[1:2], y1)
    assert torch.equal(y[[2, 0, 3, 1]], y_perm)
    rec = {}
    yo = orc.sv_dgcnn_cls(sd, t2n(x[:1]), 20, rec=rec)
    rec_g = {}
    with torch.no_grad():
        net(x[:1].contiguous(), record=rec_g)
    # kNN graph of layer 1 (xyz): bit-exact; deeper layers are chaotic -> report agreement only
    assert (t2n(rec_g["idx"][0]) == rec["idx"][0]).all()
    # teacher-forced with the oracle's graphs, logits must agree
    with torch.no_grad():
        yf = net(x[:1].contiguous(), forced_idx=[cu(i, torch.int32) for i in rec["idx"]])
    # N=1024 runs the cluster/DSMEM gate kernels and the specialised edge kernels
    assert_close(t2n(yf), yo, what="cfg2-shaped logits vs oracle (kNN graphs forced)")
    assert (t2n(yf).argmax(1) == yo.argmax(1)).all()


# ------------------------------------------------------------------------------------------------
# the other three model families
# ------------------------------------------------------------------------------------------------
OTHER_MODELS = [
    ("dgcnn_pseg_bin", "SV_DGCNN_PSEG", "sv_dgcnn_pseg", True, 4),
    ("dgcnn_pseg_fp", "SV_DGCNN_PSEG", "sv_dgcnn_pseg", True, 4),
    ("pointnet_cls_fp", "SV_PointNet_CLS", "sv_pointnet_cls", False, 1),
    ("pointnet_cls_bin", "SV_PointNet_CLS", "sv_pointnet_cls", False, 1),
    ("pointnet_pseg_bin", "SV_PointNet_PSEG", "sv_pointnet_pseg", True, 1),
    ("pointnet_pseg_fp", "SV_PointNet_PSEG", "sv_pointnet_pseg", True, 1),
]


@pytest.mark.parametrize("name,cls,ofn,with_label,nidx", OTHER_MODELS)
def test_other_models_vs_reference_and_oracle(name, cls, ofn, with_label, nidx):
    import svnet_b200 as sv
    g = golden(name)
    sd = golden_state_dict(g)
    k, binary = int(g["k"]), bool(g["binary"])
    net = quiet(getattr(sv, cls), make_args(k=k, binary=binary), int(g["ncls"]))
    net.load_state_dict(sd)
    net = net.to(DEV).eval()
    x = cu(g["x"])
    gidx = [cu(g["idx%d" % i], torch.int32) for i in range(nidx)]
    args = (x, cu(g["label"])) if with_label else (x,)
    with torch.no_grad():
        y = t2n(net(*args, forced_idx=gidx))
        rec = {}
        y_free = t2n(net(*args, record=rec))
    assert y.shape == g["logits"].shape
    assert_close(y, g["logits"], what=name + " logits (forced kNN) vs reference")
    assert (y.argmax(1) == g["logits"].argmax(1)).all()
    oargs = (sd, g["x"], g["label"], k) if with_label else (sd, g["x"], k)
    yo = getattr(orc, ofn)(*oargs, forced_idx=[g["idx%d" % i] for i in range(nidx)])
    assert_close(y, yo, what=name + " logits vs oracle")
    agree = [float((t2n(rec["idx"][i]) == g["idx%d" % i]).all(-1).mean()) for i in range(nidx)]
    assert min(agree) >= 0.98, agree
    if min(agree) == 1.0:
        assert_close(y_free, g["logits"], what=name + " logits (free-running)")


def test_checkpoint_container_roundtrip(tmp_path):
    """The reference saves {'epoch','state_dict' (module.-prefixed),...} (main_cls_dgcnn.py:208-214);
    the drop-in must load it the way main_cls_dgcnn.py:125,142-144 does (through DataParallel)."""
    import svnet_b200 as sv
    from svnet_b200.synthetic import wrap_checkpoint
    net = quiet(sv.SV_DGCNN_CLS, make_args(k=8, binary=True), 40)
    sd = synthetic_state_dict(net.state_dict(), seed=5)
    path = str(tmp_path / "sv_dgcnn_binary_modelnet40.pth")
    torch.save(wrap_checkpoint(sd), path)
    ckpt = torch.load(path, map_location="cpu")
    model = torch.nn.DataParallel(net.to(DEV), device_ids=[0])
    model.load_state_dict(ckpt["state_dict"])
    model.eval()
    x = synthetic_clouds(2, 64, 5).to(DEV)
    with torch.no_grad():
        y = model(x)
    ref = orc.sv_dgcnn_cls(sd, t2n(x), 8)
    assert_close(t2n(y), ref, what="DataParallel-wrapped forward vs oracle")


def test_full_size_configs_run_and_are_batch_independent():
    """BASELINE.json shapes beyond cfg2: part segmentation at N=2048, k=40 (cfg4, 8 of its 16 clouds
    per GPU) and a large-batch classifier pass that exercises the sub-batch chunking; every cloud's
    output must equal the output of that cloud run alone (bit for bit: clouds are independent)."""
    import svnet_b200 as sv
    from svnet_b200 import fused
    from svnet_b200.synthetic import one_hot_labels
    net = quiet(sv.SV_DGCNN_PSEG, make_args(k=40, binary=True), 50)
    net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=1004))
    net = net.to(DEV).eval()
    x = synthetic_clouds(8, 2048, 1004).to(DEV)
    l = one_hot_labels(8).to(DEV)
    with torch.no_grad():
        y = net(x, l)
        y3 = net(x[3:4].contiguous(), l[3:4].contiguous())
    assert tuple(y.shape) == (8, 50, 2048) and torch.isfinite(y).all()
    assert torch.equal(y[3:4], y3)
    cls = quiet(sv.SV_DGCNN_CLS, make_args(k=20, binary=True), 40)
    cls.load_state_dict(synthetic_state_dict(cls.state_dict(), seed=1005))
    cls = cls.to(DEV).eval()
    xb = synthetic_clouds(24, 1024, 1005).to(DEV)
    old = fused.MAX_POINTS_PER_PASS
    try:
        with torch.no_grad():
            y_full = cls(xb)
            fused.MAX_POINTS_PER_PASS = 5 * 1024      # force 5-cloud sub-batches (24 = 4*5 + 4)
            y_chunk = cls(xb)
    finally:
        fused.MAX_POINTS_PER_PASS = old
    assert torch.equal(y_full, y_chunk)


@pytest.mark.parametrize("B,N,k,binary", [(4, 1024, 20, True), (3, 200, 20, True), (2, 512, 40, True), (4, 1024, 20, False), (3, 200, 20, False),
                                               (1, 64, 20, True), (2, 4096, 20, True), (5, 1000, 20, False), (33, 1024, 20, True)])
def test_model_c_entry_equals_module(B, N, k, binary):
    """svnet_model_create / _forward / _destroy (the whole SV-DGCNN classifier, binary or fp, behind one C call,
    csrc/model.cu) against the nn.Module path on the same checkpoint tensors: bit-identical logits, also when the
    C forward is captured into a CUDA graph; uncovered shapes and kinds fail loudly."""
    import svnet_b200 as sv
    net = quiet(sv.SV_DGCNN_CLS, make_args(k=k, binary=binary), 40)
    sd = synthetic_state_dict(net.state_dict(), seed=1002)
    net.load_state_dict(sd)
    net = net.to(DEV).eval()
    x = synthetic_clouds(B, N, 1002).to(DEV)
    native = sv.NativeModel("SV_DGCNN_CLS", {"module." + n: t for n, t in sd.items()}, k=k, binary=binary, num_class=40, device=DEV)
    from svnet_b200 import fused
    keep = fused.NATIVE_FORWARD
    try:
        with torch.no_grad():
            fused.NATIVE_FORWARD = False            # the nn.Module path sequenced in Python
            y_mod = net(x)
            fused.NATIVE_FORWARD = True             # model(x) delegating to the C entry (the default)
            y_del = net(x)
            y_c = native(x)
    finally:
        fused.NATIVE_FORWARD = keep
    assert torch.equal(y_del, y_mod)
    assert tuple(y_c.shape) == (B, 40)
    assert torch.equal(y_c, y_mod)
    # graph capture of the C forward (no allocation / synchronisation inside)
    g = torch.cuda.CUDAGraph()
    xs = x.clone()
    native(xs)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        y_g = native(xs)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(y_g, y_mod)
    with pytest.raises(ValueError):
        native(synthetic_clouds(1, 48, 1).to(DEV))           # N < 64: not covered
    with pytest.raises(RuntimeError):
        sv.NativeModel("SV_DGCNN_PSEG", sd, k=k, binary=binary, num_class=40, device=DEV)
    native.close()


@pytest.mark.parametrize("B,N,k", [(2, 2048, 40), (3, 320, 20)])
def test_model_c_entry_partseg_equals_module(B, N, k):
    """svnet_model_forward_seg: the whole binary SV-DGCNN part-segmentation model behind one C call (csrc/model.cu) against
    the nn.Module path on the same checkpoint tensors: bit-identical per-point logits (B, parts, N)."""
    import svnet_b200 as sv
    from svnet_b200.synthetic import one_hot_labels
    net = quiet(sv.SV_DGCNN_PSEG, make_args(k=k, binary=True), 50)
    sd = synthetic_state_dict(net.state_dict(), seed=1004)
    net.load_state_dict(sd)
    net = net.to(DEV).eval()
    x = synthetic_clouds(B, N, 1004).to(DEV)
    l = one_hot_labels(B).to(DEV)
    native = sv.NativeModel("SV_DGCNN_PSEG", sd, k=k, binary=True, num_class=50, device=DEV)
    from svnet_b200 import fused
    keep = fused.NATIVE_FORWARD
    try:
        with torch.no_grad():
            fused.NATIVE_FORWARD = False
            y_mod = net(x, l)
            fused.NATIVE_FORWARD = True
            y_del = net(x, l)
            y_c = native(x, l)
    finally:
        fused.NATIVE_FORWARD = keep
    assert torch.equal(y_del, y_mod)
    assert tuple(y_c.shape) == (B, 50, N)
    assert torch.equal(y_c, y_mod)
    with pytest.raises(RuntimeError):
        nv_native = sv.NativeModel("SV_DGCNN_CLS", {n: t for n, t in sd.items()}, k=k, binary=True, num_class=50, device=DEV)   # wrong kind for these tensors
        del nv_native
    native.close()


@pytest.mark.parametrize("B,N,k", [(2, 2048, 40), (3, 200, 12)])
def test_seg_head_call_equals_layerwise(B, N, k):
    """svnet_seg_head_fwd (conv8 .. conv11 + the (B, parts, N) layout as one C-ABI call, csrc/seg_head.cu) against the
    same layers called one by one (sv_dgcnn_partseg.py:112-126): bit-identical logits."""
    import svnet_b200 as sv
    from svnet_b200 import sv_dgcnn_partseg as ps
    from svnet_b200.synthetic import one_hot_labels
    net = quiet(sv.SV_DGCNN_PSEG, make_args(k=k, binary=True), 50)
    net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=1004))
    net = net.to(DEV).eval()
    x = synthetic_clouds(B, N, 1004).to(DEV)
    l = one_hot_labels(B).to(DEV)
    from svnet_b200 import fused
    old, old_n = ps.SEG_HEAD_CALL, fused.NATIVE_FORWARD
    try:
        with torch.no_grad():
            fused.NATIVE_FORWARD = False            # Python-sequenced forward: the head as one call / layer by layer
            ps.SEG_HEAD_CALL = True
            y_call = net(x, l)
            ps.SEG_HEAD_CALL = False
            y_layers = net(x, l)
    finally:
        ps.SEG_HEAD_CALL, fused.NATIVE_FORWARD = old, old_n
    assert tuple(y_call.shape) == (B, 50, N) and y_call.is_contiguous()
    assert torch.equal(y_call, y_layers)


@pytest.mark.parametrize("cls_name,k,N", [("SV_DGCNN_PSEG", 40, 192), ("SV_DGCNN_CLS", 20, 320)])
def test_fast_edge_kernels_teacher_forced_k_full(cls_name, k, N):
    """The shape-specialised edge kernels at the BASELINE neighbourhood sizes (k=20 cls, k=40 pseg:
    two popcount passes, 4-warp CTAs), teacher-forced with the oracle's per-layer inputs and graphs:
    pooled scalars bit-identical, pooled vectors within tolerance."""
    import svnet_b200 as sv
    net = quiet(getattr(sv, cls_name), make_args(k=k, binary=True), 50 if "PSEG" in cls_name else 40)
    sd = synthetic_state_dict(net.state_dict(), seed=2000 + k)
    net.load_state_dict(sd)
    net = net.to(DEV).eval()
    x = synthetic_clouds(1, N, 77 + k, rotate=True)
    rec_o = {}
    P = orc.Params(sd)
    outs = orc._dgcnn_trunk(P, x.numpy(), k, None, rec_o.setdefault("r", {"idx": [], "pools": []}))
    ro = rec_o["r"]
    teacher = [(cu(s).view(N, -1), cu(v).view(N, 3, -1)) for (s, v) in ro["pools"][:3]]
    rec = {"teacher": teacher}
    fi = [cu(i, torch.int32) for i in ro["idx"]]
    with torch.no_grad():
        if "PSEG" in cls_name:
            from svnet_b200.synthetic import one_hot_labels
            net(x.to(DEV), one_hot_labels(1).to(DEV), forced_idx=fi, record=rec)
        else:
            net(x.to(DEV), forced_idx=fi, record=rec)
    so = vo = 0
    for li, (o_s, o_v) in enumerate(ro["pools"]):
        cs, cv = o_s.shape[-1], o_v.shape[-1]
        got_s = t2n(rec["s_cat"][:, so:so + cs]).reshape(1, N, cs)
        got_v = t2n(rec["v_cat"][:, :, vo:vo + cv]).reshape(1, N, 3, cv)
        assert (got_s == o_s).all(), "layer %d pooled scalars differ from the oracle" % (li + 1)
        assert_close(got_v, o_v, rtol=1e-4, atol=1e-5, what="layer %d v" % (li + 1))
        so += cs
        vo += cv


def test_rotate_permute_input_transform():
    """Input side of the eval loop (main_cls_dgcnn.py:229-235): (B,N,3) @ R then permute -> (B,3,N)."""
    from svnet_b200.evalutil import random_rotations, rotate_points
    pts = synthetic_clouds(3, 200, 4).transpose(1, 2).contiguous()          # (B,N,3)
    R = random_rotations(3, generator=torch.Generator().manual_seed(2))
    out = rotate_points(pts.to(DEV), R.to(DEV))
    ref = torch.bmm(pts.double(), R.double()).permute(0, 2, 1)
    assert tuple(out.shape) == (3, 3, 200)
    assert_close(t2n(out), ref.numpy(), rtol=1e-6, atol=1e-6)
    assert torch.equal(rotate_points(pts.to(DEV)).cpu(), pts.permute(0, 2, 1))
    # SO(3) invariance of the fp classifier: rotated input, same logits
    import svnet_b200 as sv
    net = quiet(sv.SV_DGCNN_CLS, make_args(k=12, binary=False), 40)
    net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=33))
    net = net.to(DEV).eval()
    with torch.no_grad():
        y0 = net(rotate_points(pts.to(DEV)))
        y1 = net(out)
    assert_close(t2n(y1), t2n(y0), rtol=1e-3, atol=1e-4, what="SO(3) invariance of fp logits")


def test_vector_linear_tensor_core_paths_agree(monkeypatch):
    """The binary-weight vector linear + VectorBN + gate has three implementations: CUDA cores
    (sequential fp32 chains), mma.sync bf16 and tcgen05/TMEM bf16, the latter two on an exact 3-way
    bf16 split of the activations.  All must agree to fp32 rounding, and with the oracle."""
    from svnet_b200 import _native as nv
    R, K, N, rpc = 1000, 83, 170, 250
    v, W = rnd((R, 3, K), 1), rnd((N, K), 2)
    sc = np.abs(rnd((N,), 3)) + 0.5
    bn = tuple(np.ascontiguousarray(t, dtype=np.float32) for t in (np.abs(rnd((N,), 4)) + 0.5, rnd((N,), 5, 0.1),
                                                                   rnd((N,), 6, 0.1), np.abs(rnd((N,), 7)) + 0.5))
    gate = 1.0 / (1.0 + np.exp(-rnd((R // rpc, N), 8)))
    ref = orc.scale_v(orc.vector_bn(orc.linear(v, W, scale=sc, bw=True), bn).reshape(R // rpc, rpc, 3, N), gate)
    inv = 1.0 / np.sqrt(bn[3] + np.float32(1e-5))
    a_np = (bn[0] * inv).astype(np.float32)
    c_np = (bn[1] - bn[2] * a_np).astype(np.float32)
    outs = {}
    for name, env in (("cuda_cores", {"SVNET_TCGEN05": "0", "SVNET_NO_TC": "1"}),
                      ("mma_sync", {"SVNET_TCGEN05": "0", "SVNET_NO_TC": "0"}),
                      ("tcgen05", {"SVNET_TCGEN05": "1", "SVNET_NO_TC": "0"})):
        for k_, v_ in env.items():
            monkeypatch.setenv(k_, v_)
        out = torch.zeros(R, 3, N, device=DEV)
        vt = cu(v)
        nv.linear_rows(vt, vt.stride(0), vt.stride(1), 3, 3 * R, K, cu(W), N, out, out.stride(0), out.stride(1),
                       sign_w=True, colscale=cu(sc), bn=(cu(a_np), cu(c_np)), vbn=True, gate=cu(gate.astype(np.float32)),
                       groups_per_cloud=rpc)
        outs[name] = t2n(out).reshape(R // rpc, rpc, 3, N)
        assert_close(outs[name], ref, what=name + " vs oracle")          # north_star tolerance
    # the tensor-core paths only reorder the fp32 summation: typical agreement is ~1e-6 relative
    for name in ("mma_sync", "tcgen05"):
        assert_close(outs[name], outs["cuda_cores"], what=name + " vs CUDA cores")
        rel = np.abs(outs[name] - outs["cuda_cores"]).mean() / np.abs(outs["cuda_cores"]).mean()
        assert rel < 1e-5, (name, rel)


def test_point_permutation_invariance_fp():
    """SURVEY.md 4.3 property test: the fp classifier's logits do not depend on the order of the
    points of a cloud (kNN, pooling and the gates are symmetric functions of the point set)."""
    import svnet_b200 as sv
    net = quiet(sv.SV_DGCNN_CLS, make_args(k=16, binary=False), 40)
    net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=44))
    net = net.to(DEV).eval()
    x = synthetic_clouds(2, 300, 9).to(DEV)
    perm = torch.randperm(300, generator=torch.Generator().manual_seed(3)).to(DEV)
    with torch.no_grad():
        y0 = net(x)
        y1 = net(x[:, :, perm].contiguous())
    assert_close(t2n(y1), t2n(y0), rtol=1e-3, atol=1e-4, what="permutation invariance")


def test_graphed_forward_equals_eager():
    """svnet_b200.GraphedForward (CUDA-graph replay, batch halves on two streams) returns exactly what the
    eager single-stream forward returns, for new inputs of the captured shape, and rejects other shapes."""
    import svnet_b200 as sv
    from svnet_b200 import fused
    net = quiet(sv.SV_DGCNN_CLS, make_args(k=20, binary=True), 40)
    net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=3))
    net = net.to(DEV).eval()
    x0 = synthetic_clouds(16, 256, 11).to(DEV)
    fast = sv.GraphedForward(net, x0)
    for seed in (12, 13):
        x = synthetic_clouds(16, 256, seed).to(DEV)
        y_fast = fast(x).clone()
        keep = fused.CONCURRENT_HALVES
        fused.CONCURRENT_HALVES = False
        try:
            with torch.no_grad():
                y = net(x)
        finally:
            fused.CONCURRENT_HALVES = keep
        assert torch.equal(y_fast, y)
    with pytest.raises(ValueError):
        fast(synthetic_clouds(8, 256, 1).to(DEV))


@pytest.mark.parametrize("model", ["pseg", "pointnet"])
def test_two_stream_halves_equal_single_stream(model):
    """fused.chunked runs a batch of >= 16 clouds as two halves on two CUDA streams; clouds are independent in
    eval mode, so the result must equal the single-stream forward exactly (also with an extra per-cloud input)."""
    import svnet_b200 as sv
    from svnet_b200 import fused
    from svnet_b200.synthetic import one_hot_labels
    B, N = 16, 256
    x = synthetic_clouds(B, N, 21).to(DEV)
    if model == "pseg":
        net = quiet(sv.SV_DGCNN_PSEG, make_args(k=12, binary=True), 50)
        extra = (one_hot_labels(B).to(DEV),)
    else:
        net = quiet(sv.SV_PointNet_CLS, make_args(k=12, binary=False), 40)
        extra = ()
    net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=5))
    net = net.to(DEV).eval()
    keep = fused.CONCURRENT_HALVES
    try:
        with torch.no_grad():
            fused.CONCURRENT_HALVES = True
            y2 = net(x, *extra)
            fused.CONCURRENT_HALVES = False
            y1 = net(x, *extra)
    finally:
        fused.CONCURRENT_HALVES = keep
    torch.cuda.synchronize()
    assert y2.shape == y1.shape and torch.equal(y2, y1)


@pytest.mark.parametrize("model,binary", [("cls", True), ("cls", False), ("pseg", True)])
def test_auxiliary_streams_do_not_change_bits(model, binary):
    """Round 2 runs independent chains of a forward on auxiliary streams (per-point tables / Ya|Yb next to the kNN kernels,
    conv5's scalar branch next to gate -> vector linear, the part-seg head's early sign words and label branch): with the
    side streams switched off -- everything on one stream, the seg head's sign words computed inside svnet_seg_head_fwd --
    the outputs must be the same bits, repeatedly."""
    import svnet_b200 as sv
    from svnet_b200 import fused
    from svnet_b200.synthetic import one_hot_labels
    B, N, k = (4, 1024, 20) if model == "cls" else (2, 2048, 40)
    x = synthetic_clouds(B, N, 33).to(DEV)
    if model == "pseg":
        net = quiet(sv.SV_DGCNN_PSEG, make_args(k=k, binary=True), 50)
        extra = (one_hot_labels(B).to(DEV),)
    else:
        net = quiet(sv.SV_DGCNN_CLS, make_args(k=k, binary=binary), 40)
        extra = ()
    net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=9))
    net = net.to(DEV).eval()
    keep, keep_n = fused.SIDE_STREAM, fused.NATIVE_FORWARD
    try:
        with torch.no_grad():
            fused.NATIVE_FORWARD = False            # the Python-sequenced path is the one under test
            fused.SIDE_STREAM = True
            ys = [net(x, *extra).clone() for _ in range(3)]
            fused.SIDE_STREAM = False
            y1 = net(x, *extra)
    finally:
        fused.SIDE_STREAM, fused.NATIVE_FORWARD = keep, keep_n
    torch.cuda.synchronize()
    for y in ys:
        assert torch.equal(y, y1)


# ------------------------------------------------------------------------------------------------
# tensor-core edge kernel (csrc/edge_tc.cu) against the XNOR/popcount kernel and against itself
# ------------------------------------------------------------------------------------------------
def _run_edge_layer(blk, s_in, v_in, idx32, B, N, k, tc):
    import os
    from svnet_b200 import fused
    Cout, Cvo = blk.out_dims
    s_out = torch.empty((B * N, Cout), device=DEV)
    v_out = torch.empty((B * N, 3, Cvo), device=DEV)
    taps = {}
    old = os.environ.get("SVNET_EDGE_TC")
    os.environ["SVNET_EDGE_TC"] = "1" if tc else "0"
    try:
        with torch.no_grad():
            fused.sv_edge_layer(s_in, v_in, B, N, k, blk, s_out, v_out, idx32=idx32, taps=taps)
        torch.cuda.synchronize()
    finally:
        if old is None:
            del os.environ["SVNET_EDGE_TC"]
        else:
            os.environ["SVNET_EDGE_TC"] = old
    return s_out, v_out, taps


@pytest.mark.parametrize("Cs,Cv,Cout,Cvo,k,B,N", [
    (32, 10, 32, 10, 20, 2, 256), (32, 10, 64, 21, 20, 1, 100), (64, 21, 128, 42, 20, 2, 256),
    (32, 16, 32, 16, 40, 1, 130), (32, 16, 64, 24, 40, 2, 128), (64, 24, 128, 40, 40, 1, 256),
    (64, 21, 128, 42, 40, 1, 96), (32, 16, 64, 24, 20, 1, 64)])
def test_edge_tensor_core_kernel(Cs, Cv, Cout, Cvo, k, B, N):
    """linear1 of the binary edge layers on tcgen05 (fp8 ternary operands, exact integer accumulators):
    * the pooled scalars equal, bit for bit, what the kernel's own sign bytes (read back through the taps) give
      under integer arithmetic and the float epilogue chain -- operand layout, K permutation, descriptors,
      accumulators and epilogue are all covered by this;
    * the s-channel signs equal the popcount kernel's exactly; q-channel signs (frames from the per-point
      table) may differ only where |q + beta| is at rounding level;
    * the vector branch is the same code: identical at k = 20, summation-order tolerance at k = 40."""
    import svnet_b200 as sv
    from svnet_b200 import _native as nv
    if nv.edge_tc_weight_bytes(Cs, Cv, Cout, Cvo, k) == 0:
        pytest.skip("tensor-core edge path disabled")
    blk = quiet(sv.SVBlock, (2 * Cs, 2 * Cv), (Cout, Cvo), True)
    blk.load_state_dict(synthetic_state_dict(blk.state_dict(), seed=77 + Cs + Cv))
    blk = blk.to(DEV).eval()
    R = B * N
    s_in = cu(rnd((R, Cs), 3))
    v_in = cu(rnd((R, 3, Cv), 4))
    g = torch.Generator().manual_seed(5)
    idx = torch.randint(0, N, (B, N, k), generator=g, dtype=torch.int32)
    idx[:, :, 0] = torch.arange(N, dtype=torch.int32)            # the self edge: exact zeros in the difference channels
    idx32 = idx.to(DEV)
    s_tc, v_tc, taps_tc = _run_edge_layer(blk, s_in, v_in, idx32, B, N, k, tc=True)
    s_pc, v_pc, taps_pc = _run_edge_layer(blk, s_in, v_in, idx32, B, N, k, tc=False)
    K = 2 * Cs + 6 * Cv
    bt, mt = _unpack(t2n(taps_tc["bits"]), K), _unpack(t2n(taps_tc["mask"]), K)
    bp, mp = _unpack(t2n(taps_pc["bits"]), K), _unpack(t2n(taps_pc["mask"]), K)
    t_tc = np.where(mt == 1, np.where(bt == 1, 1, -1), 0).astype(np.int32)           # (R*k, K)
    t_pc = np.where(mp == 1, np.where(bp == 1, 1, -1), 0).astype(np.int32)
    assert (t_tc[:, :2 * Cs] == t_pc[:, :2 * Cs]).all(), "s-channel signs differ"
    assert (t_tc[:, :Cs].reshape(R, k, Cs)[:, 0] == np.sign(t2n(blk.linear1.beta)[0, :Cs])).all()   # self edge: sign(0 + beta)
    nd = int((t_tc != t_pc).sum())
    assert nd <= 1e-4 * t_tc.size, "%d of %d signs differ" % (nd, t_tc.size)
    # pooled scalars from the kernel's own signs: integer dot, then the kernel's float chain
    W = np.sign(t2n(blk.linear1.weight)).astype(np.int32)                             # (Cout, K)
    dot = (t_tc @ W.T).reshape(R, k, Cout)
    sc = t2n(blk.linear1.scale).reshape(-1).astype(np.float32)
    a1, c1 = (t2n(t).astype(np.float32) for t in blk.bn1_folded())
    y = (dot.astype(np.float32) * sc).astype(np.float32)
    y = ((y * a1).astype(np.float32) + c1).astype(np.float32)
    y = np.where(y > 0, y, (np.float32(0.2) * y).astype(np.float32))
    assert (t2n(s_tc) == y.max(axis=1)).all(), "pooled scalars differ from the integer recomputation"
    if nd == 0:
        assert torch.equal(s_tc, s_pc)
    if k == 20:
        assert torch.equal(v_tc, v_pc)
    else:
        assert_close(t2n(v_tc), t2n(v_pc), rtol=1e-5, atol=1e-6, what="vector branch")
