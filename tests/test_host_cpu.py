"""CPU tests (no GPU): the C-ABI library loads and exports every symbol the header declares, the
nn.Module mirror has the reference's state_dict layout, host-side logic (sharding, synthetic
checkpoints, error behaviour without a GPU)."""
import json
import os
import re

import numpy as np
import pytest
import torch

from tests.util import GOLDEN, golden, quiet
from svnet_b200.synthetic import (make_args, state_dict_digest, strip_module_prefix, synthetic_clouds,
                                  synthetic_state_dict, wrap_checkpoint)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_header_symbols():
    from svnet_b200 import _native as nv
    lib = nv.lib()
    header = open(os.path.join(ROOT, "include", "svnet_b200.h")).read()
    declared = set(re.findall(r"\b(svnet_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), "libsvnet_b200.so does not export %s" % name
    assert set(nv.EXPORTS) == declared
    assert lib.svnet_version() == 3


def test_argument_errors_without_gpu():
    """Shape validation happens before any launch, so it is testable without a device."""
    import ctypes
    from svnet_b200 import _native as nv
    lib = nv.lib()
    v = nv.View()
    rc = lib.svnet_knn(ctypes.byref(v), 1, 8, 9, None, None, None)
    assert rc == -1 and b"out of range" in lib.svnet_last_error()
    rc = lib.svnet_pool_rows(None, 4, 4, 1, ctypes.c_long(4), None, None, 4, None)
    assert rc == -1
    with pytest.raises(RuntimeError):
        import svnet_b200 as sv
        sv.knn(torch.zeros(1, 3, 16), 4)  # CPU tensor: no fallback path exists
    # workspace-based entry points: shape queries need no device; bad workspaces are rejected before any launch
    lib.svnet_knn_workspace_bytes.restype = ctypes.c_size_t
    lib.svnet_svfuse_pool_workspace.restype = ctypes.c_size_t
    v.Cs, v.lds = 62, 62
    assert lib.svnet_knn_workspace_bytes(ctypes.byref(v), 4, 1024, 20) > 0          # covered by the tensor-core path
    assert lib.svnet_knn_workspace_bytes(ctypes.byref(v), 4, 1024, 40) > 0          # 32 < k <= 48: wide finish kernel
    assert lib.svnet_knn_workspace_bytes(ctypes.byref(v), 4, 1024, 100) == 0        # k > 48: CUDA-core kernel
    assert lib.svnet_knn_workspace_bytes(ctypes.byref(v), 4, 32, 8) == 0            # N < 64
    assert lib.svnet_svfuse_pool_workspace(2, 170, ctypes.c_long(1024)) > 0
    rc = lib.svnet_svfuse_pool(None, 1, ctypes.c_long(8), None, None, None, None, 0, None, ctypes.c_size_t(0), None)
    assert rc == -1 and b"svnet_svfuse_pool" in lib.svnet_last_error()
    with pytest.raises(RuntimeError):
        net = sv.SV_DGCNN_CLS(__import__("svnet_b200.synthetic", fromlist=["make_args"]).make_args(k=4, binary=True), 40).eval()
        sv.GraphedForward(net, torch.zeros(1, 3, 16))     # CPU example input: no CPU path


MODEL_FIXTURES = [("dgcnn_cls_bin", "SV_DGCNN_CLS"), ("dgcnn_cls_fp", "SV_DGCNN_CLS"), ("dgcnn_pseg_bin", "SV_DGCNN_PSEG"),
                  ("dgcnn_pseg_fp", "SV_DGCNN_PSEG"), ("pointnet_cls_bin", "SV_PointNet_CLS"),
                  ("pointnet_cls_fp", "SV_PointNet_CLS"), ("pointnet_pseg_bin", "SV_PointNet_PSEG"),
                  ("pointnet_pseg_fp", "SV_PointNet_PSEG")]


@pytest.mark.parametrize("fixture,cls", MODEL_FIXTURES)
def test_state_dict_layout_matches_reference(fixture, cls):
    import svnet_b200 as sv
    g = golden(fixture)
    net = quiet(getattr(sv, cls), make_args(k=int(g["k"]), binary=bool(g["binary"])), int(g["ncls"]))
    shapes = json.loads(str(g["sd_shapes"]))
    sd = net.state_dict()
    assert list(sd.keys()) == list(shapes.keys())  # same keys in the same order as the reference
    for key, (shape, dtype) in shapes.items():
        assert list(sd[key].shape) == shape and str(sd[key].dtype) == dtype, key
    # reference checkpoints are saved from DataParallel ('module.' prefix)
    ck = wrap_checkpoint(synthetic_state_dict(sd, seed=3))
    assert all(k.startswith("module.") for k in ck["state_dict"])
    net.load_state_dict(strip_module_prefix(ck["state_dict"]))


def test_synthetic_checkpoint_is_deterministic_and_representative():
    import svnet_b200 as sv
    net = quiet(sv.SV_DGCNN_CLS, make_args(k=20, binary=True), 40)
    a = synthetic_state_dict(net.state_dict(), seed=11)
    b = synthetic_state_dict(net.state_dict(), seed=11)
    assert state_dict_digest(a) == state_dict_digest(b)
    assert state_dict_digest(a) != state_dict_digest(synthetic_state_dict(net.state_dict(), seed=12))
    assert (a["conv2.linear1.beta"] != 0).all() and (a["conv2.bn1.running_var"] > 0).all()
    z = synthetic_state_dict(net.state_dict(), seed=11, beta_zero=True)
    assert (z["conv2.linear1.beta"] == 0).all()
    x = synthetic_clouds(3, 64, 1)
    assert x.shape == (3, 3, 64) and abs(float(x.transpose(1, 2).norm(dim=2).max()) - 1.0) < 1e-5


def test_training_mode_is_rejected():
    import svnet_b200 as sv
    blk = quiet(sv.SVBlock, (8, 2), (4, 2), True)
    with pytest.raises(RuntimeError):
        blk((torch.zeros(1, 4, 8), torch.zeros(1, 4, 3, 2)))


def test_shard_bounds_cover_batch():
    from svnet_b200.parallel import shard_bounds
    for batch in (1, 7, 32, 33, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, batch, q):
    import torch.distributed as dist
    from svnet_b200.parallel import ShardedInference
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    torch.manual_seed(0)
    x = torch.randn(batch, 3, 16)
    lab = torch.randn(batch, 4)
    f = lambda a, b: torch.cat([a.sum(dim=2) * 2.0, b], dim=1)  # per-cloud function, like an eval-mode forward
    y = ShardedInference(f)(x, lab)
    q.put((rank, torch.equal(y, f(x, lab))))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 5, 1])
def test_sharded_inference_gloo_world2(batch):
    """N>1 path on CPU: batch sharding + all-gather reproduce the unsharded result on every rank."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + batch) % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == [0, 1] and all(ok for _, ok in res)


def test_shape_iou_matches_reference():
    from svnet_b200.evalutil import calculate_shape_IoU, classification_metrics
    g = golden("metrics")
    ious = calculate_shape_IoU(g["pred"], g["seg"], g["label"])
    assert np.allclose(ious, g["ious"], rtol=0, atol=1e-12)
    acc, bal = classification_metrics([0, 0, 1, 1, 2], [0, 1, 1, 1, 2])
    assert abs(acc - 0.8) < 1e-12 and abs(bal - (0.5 + 1 + 1) / 3) < 1e-12


def test_checkpoint_helpers(tmp_path):
    import svnet_b200 as sv
    from svnet_b200.evalutil import load_checkpoint, load_weights, random_rotations
    net = quiet(sv.SV_PointNet_CLS, make_args(k=8, binary=True), 40)
    sd = synthetic_state_dict(net.state_dict(), seed=9)
    path = str(tmp_path / "sv_pointnet_binary_modelnet40.pth")
    torch.save(wrap_checkpoint(sd), path)
    state = load_checkpoint(path)
    assert set(state) >= {"epoch", "state_dict"} and all(k.startswith("module.") for k in state["state_dict"])
    load_weights(net, path)
    assert state_dict_digest(net.state_dict()) == state_dict_digest(sd)
    R = random_rotations(5, generator=torch.Generator().manual_seed(1))
    eye = torch.eye(3).expand(5, 3, 3)
    assert torch.allclose(R @ R.transpose(1, 2), eye, atol=1e-5) and torch.allclose(torch.linalg.det(R), torch.ones(5), atol=1e-5)


def test_dataset_readers_without_h5py(tmp_path):
    """the npz readers return what the reference's HDF5 readers return (data.py:71-115,
    186-201, 260-300, 302-340) -- same globbing of shards, dtypes, item tuples and partition behaviour."""
    from svnet_b200 import data as D
    rng = np.random.default_rng(7)
    mn = tmp_path / "modelnet40_ply_hdf5_2048"
    mn.mkdir()
    shards = []
    for i in range(2):
        pts, lab = rng.standard_normal((5, 2048, 3)).astype("float64"), rng.integers(0, 40, (5, 1)).astype("uint8")
        np.savez(mn / f"ply_data_test{i}.npz", data=pts, label=lab)
        shards.append((pts, lab))
    np.savez(mn / "ply_data_train0.npz", data=shards[0][0][:2], label=shards[0][1][:2])
    ds = D.ModelNet40(num_points=1024, data_dir=str(tmp_path), partition="test")
    assert len(ds) == 10 and ds.data.dtype == np.float32 and ds.label.dtype == np.int64
    cloud, label = ds[7]
    assert cloud.shape == (1024, 3) and np.array_equal(cloud, shards[1][0][2, :1024].astype("float32")) and label[0] == shards[1][1][2, 0]
    # train items: the reference's draw order (scale, shift, then an in-place shuffle of the points)
    tr = D.ModelNet40(num_points=64, data_dir=str(tmp_path), partition="train")
    np.random.seed(3)
    got, _ = tr[1]
    np.random.seed(3)
    want = D.translate_pointcloud(tr.data[1][:64])
    np.random.shuffle(want)
    assert np.array_equal(got, want)

    sp = tmp_path / "shapenet_part_seg_hdf5_data"
    sp.mkdir()
    for name, n in (("ply_data_train0", 4), ("ply_data_val0", 3), ("ply_data_test0", 6)):
        np.savez(sp / (name + ".npz"), data=rng.standard_normal((n, 2048, 3)).astype("float32"),
                 label=(np.arange(n) % 3).reshape(n, 1).astype("uint8"), pid=rng.integers(0, 50, (n, 2048)).astype("uint8"))
    te = D.ShapeNetPart(num_points=2048, data_dir=str(tmp_path), partition="test")
    assert len(te) == 6 and te.seg_num_all == 50 and te.seg_start_index == 0
    pts, lab, seg = te[0]
    assert pts.shape == (2048, 3) and seg.shape == (2048,) and seg.dtype == np.int64 and lab.shape == (1,)
    tv = D.ShapeNetPart(num_points=256, data_dir=str(tmp_path), partition="trainval")
    assert len(tv) == 7
    pts, _, seg = tv[2]                                # shuffled, but every point keeps its part id
    base = {tuple(p): s for p, s in zip(tv.data[2][:256], tv.seg[2][:256])}
    assert all(base[tuple(p)] == s for p, s in zip(pts, seg))
    bag = D.ShapeNetPart(num_points=2048, data_dir=str(tmp_path), partition="test", class_choice="bag")
    assert len(bag) == 2 and bag.seg_num_all == 2 and bag.seg_start_index == 4 and (bag.label == 1).all()

    so = tmp_path / "h5_files" / "main_split"
    so.mkdir(parents=True)
    np.savez(so / "test_objectdataset.npz", data=rng.standard_normal((3, 2048, 3)).astype("float32"), label=np.arange(3))
    sc = D.ScanObjectNNCls(num_points=1024, data_dir=str(tmp_path), partition="test")
    pts, lab = sc[1]
    assert pts.shape == (1024, 3) and lab == 1 and {tuple(p) for p in pts} <= {tuple(p) for p in sc.points[1]}
    with pytest.raises(ValueError):
        D.ScanObjectNNCls(num_points=1024, data_dir=str(tmp_path), partition="val")
    with pytest.raises(FileNotFoundError):
        D.ModelNet40(num_points=1024, data_dir=str(tmp_path / "nowhere"), partition="test")
    unit = D.pc_normalize(shards[0][0][0])
    assert abs(np.linalg.norm(unit, axis=1).max() - 1.0) < 1e-12 and np.abs(unit.mean(0)).max() < 1e-12


def test_dataset_readers_equal_reference_readers(tmp_path):
    """SURVEY 8(f) f4, pinned: tests/golden/data_readers.npz holds what the UNMODIFIED reference readers
    (data.py:186-201, 260-340, run over a stand-in h5py by tests/golden/make_golden_data.py) returned for seeded
    items of every partition; the npz readers must return the same arrays, dtypes and draw order."""
    import importlib.util
    from svnet_b200 import data as D
    spec = importlib.util.spec_from_file_location("make_golden_data", os.path.join(GOLDEN, "make_golden_data.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gold = np.load(os.path.join(GOLDEN, "data_readers.npz"))
    shards = {}
    for key in gold.files:
        if key.startswith("shard|"):
            _, rel, name = key.split("|")
            shards.setdefault(rel, {})[name] = gold[key]
    assert len(shards) == 9
    gen.write_shards(str(tmp_path), shards)
    got = gen.run_cases(D, str(tmp_path))
    assert len(got) > 40
    for key, val in got.items():
        want = gold[key]
        if key.endswith("_dtype"):
            assert str(val) == str(want), key
        else:
            assert np.asarray(val).shape == want.shape and np.array_equal(val, want), key
    assert np.array_equal(D.pc_normalize(gold["pc_normalize_in"]), gold["pc_normalize_out"])
    np.random.seed(5)
    assert np.array_equal(D.translate_pointcloud(shards["modelnet40_ply_hdf5_2048/ply_data_test0"]["data"][0]), gold["translate_out"])


def test_dataset_readers_take_h5_files_where_h5py_exists(tmp_path, monkeypatch):
    """With h5py importable and no .npz shard present, the readers open the reference's .h5 files themselves
    (data.py:79-82): same golden items.  h5py is not installed here: a stand-in serves npz content under .h5 names."""
    import importlib.util
    import types
    from svnet_b200 import data as D
    spec = importlib.util.spec_from_file_location("make_golden_data", os.path.join(GOLDEN, "make_golden_data.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gold = np.load(os.path.join(GOLDEN, "data_readers.npz"))

    class File:
        def __init__(self, name, mode="r"):
            assert name.endswith(".h5")
            self._z = np.load(name)

        def __getitem__(self, key):
            return self._z[key]

        def close(self):
            self._z.close()

    for key in gold.files:
        if key.startswith("shard|"):
            _, rel, name = key.split("|")
            os.makedirs(os.path.dirname(os.path.join(str(tmp_path), rel)), exist_ok=True)
    shards = {}
    for key in gold.files:
        if key.startswith("shard|"):
            _, rel, name = key.split("|")
            shards.setdefault(rel, {})[name] = gold[key]
    for rel, arrays in shards.items():
        with open(os.path.join(str(tmp_path), rel + ".h5"), "wb") as f:
            np.savez(f, **arrays)
    with pytest.raises(FileNotFoundError):                       # no h5py, no npz: loud
        monkeypatch.setattr(D, "_h5py", lambda: None)
        D.ModelNet40(num_points=64, data_dir=str(tmp_path), partition="test")
    monkeypatch.setattr(D, "_h5py", lambda: types.SimpleNamespace(File=File))
    got = gen.run_cases(D, str(tmp_path))
    for key, val in got.items():
        if not key.endswith("_dtype"):
            assert np.array_equal(val, gold[key]), key


def test_sub_batch_bounds_cover_the_batch():
    """fused.chunked splits a batch into contiguous sub-batches that run concurrently (clouds are
    independent in eval mode): every cloud exactly once, none smaller than the minimum."""
    from svnet_b200 import fused
    for B in (16, 17, 31, 32, 33, 100, 256, 4096):
        b = fused.split_bounds(B)
        assert b[0] == 0 and b[-1] == B and all(lo < hi for lo, hi in zip(b[:-1], b[1:]))
        assert len(b) - 1 <= fused.N_SPLIT and min(hi - lo for lo, hi in zip(b[:-1], b[1:])) >= fused.MIN_CLOUDS
    assert fused.split_bounds(fused.MIN_CLOUDS - 1 or 1) == [0, fused.MIN_CLOUDS - 1 or 1]


def test_packed_cache_keys_replicas_on_their_source_and_invalidates():
    """The packed-weight cache: nn.DataParallel replicas (fresh parameter copies every forward) resolve their
    dependencies on the source module, so nothing is re-packed per forward; load_state_dict / .to() and
    invalidate_packed() drop stale entries (in-place writes through .data do not bump _version)."""
    import svnet_b200 as sv
    from svnet_b200.sv_layers import _Cached
    lin = sv.Linear(8, 4, bias=False, bw=True, ba=True).eval()
    n0 = _Cached.BUILDS
    a = lin.scale_vec()
    assert _Cached.BUILDS == n0 + 1 and lin.scale_vec() is a and _Cached.BUILDS == n0 + 1
    # what torch.nn.parallel.replicate does for every forward
    rep = lin._replicate_for_data_parallel()
    rep._parameters = {k: torch.nn.Parameter(v.detach().clone()) for k, v in lin._parameters.items() if v is not None}
    assert rep.__dict__["_sv_src"] is lin
    assert rep.scale_vec() is a and _Cached.BUILDS == n0 + 1          # same device: the source's entry, no build
    rep2 = rep._replicate_for_data_parallel()
    assert rep2.__dict__["_sv_src"] is lin
    # stale-data protection
    lin.scale.data.fill_(3.0)                                          # no version bump
    assert lin.scale_vec() is a
    lin.invalidate_packed()
    assert float(lin.scale_vec()[0]) == 3.0
    b = lin.scale_vec()
    lin.load_state_dict(lin.state_dict())
    assert lin.scale_vec() is not b
    blk = quiet(sv.SVBlock, (8, 2), (4, 2), True).eval()
    w1, _ = blk.pq_weight()
    blk.double().float()                                               # _apply invalidates children too
    assert blk.pq_weight()[0] is not w1
    seq = quiet(sv.SV_DGCNN_PSEG, make_args(k=4, binary=True), 50)     # '1.weight' paths of the nn.Sequential blocks
    assert seq.conv8._packed("probe", ("1.weight",), lambda: 1) == 1


def test_committed_bench_line_keeps_the_contract():
    """The newest bench line under profiles/ (written by `python bench.py` on the GPU box) carries every key of the
    bench contract and is internally consistent: value = clouds per step / time, roofline.frac = achieved / peak,
    achieved = algorithmic bytes per launch / kernel time, e2e measured with real copies, launches counted."""
    prof = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
    names = sorted(n for n in os.listdir(prof) if re.fullmatch(r"r2[a-z]_bench\.json", n))
    assert names, "no bench line committed"
    d = json.loads(open(os.path.join(prof, names[-1])).read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["unit"] == "clouds/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    per_step = d["config"]["global_batch"]
    assert abs(d["value"] - per_step * 1000.0 / d["ms_per_step"]) <= 1e-6 * d["value"]
    assert d["warmup"] >= 3 and d["gpu_launches"] >= d["steps"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and 0 < e["value"] < d["value"]
    assert e["h2d_bytes_per_step"] == per_step * 3 * d["config"]["n_points"] * 4 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["kernel_ms"] * 1e-3) / 1e9) <= 1e-6 * r["achieved"]
    assert r["kernel_ms"] < d["ms_per_step"] and r["traffic"] >= r["algorithmic_bytes_per_launch"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["unit"] == d["unit"] and c["sample"]
    k = d["clocks"]
    assert not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert k["sm_mhz"] >= 0.9 * k["sm_max_mhz"]
