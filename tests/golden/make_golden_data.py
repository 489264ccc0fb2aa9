#!/usr/bin/env python
"""Pin svnet_b200/data.py to the reference's own dataset readers.

h5py is not installed in this image, so the UNMODIFIED /root/reference/data.py is imported with a
stand-in ``h5py`` module whose ``File`` serves the arrays of the ``.npz`` shard that sits next to the
``.h5`` name it was asked for (only ``f[key][:]``, ``np.array(f[key])`` and ``close()`` are used by
data.py:79-82, 104-108, 318-321).  The reference's classes then read small seeded shards written
here, and the items they return under ``np.random.seed(s)`` are stored in
``tests/golden/data_readers.npz`` together with the shard arrays, so the CPU test can rebuild the
same directory and compare ``svnet_b200.data`` item by item (tests/test_host_cpu.py).

Run (in the build container only):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_data.py
"""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True


class _File:
    def __init__(self, name, mode="r"):
        self._z = np.load(os.path.splitext(name)[0] + ".npz")

    def __getitem__(self, key):
        return self._z[key]

    def keys(self):
        return self._z.files

    def close(self):
        self._z.close()


def shard_arrays():
    """name -> dict of arrays; dtypes as the public HDF5 files hold them (float32 points, uint8 labels / part ids)."""
    rng = np.random.default_rng(20260)
    P = 96
    out = {}
    for part, n in (("train0", 3), ("test0", 4), ("test1", 2)):
        out["modelnet40_ply_hdf5_2048/ply_data_" + part] = dict(
            data=rng.standard_normal((n, P, 3)).astype("float32"), label=rng.integers(0, 40, (n, 1)).astype("uint8"))
    for part, n in (("train0", 3), ("val0", 2), ("test0", 5)):
        out["shapenet_part_seg_hdf5_data/ply_data_" + part] = dict(
            data=rng.standard_normal((n, P, 3)).astype("float32"), label=(np.arange(n) % 3 + 1).reshape(n, 1).astype("uint8"),
            pid=rng.integers(0, 50, (n, P)).astype("uint8"))
    for name, n in (("training_objectdataset", 3), ("test_objectdataset", 3), ("test_objectdataset_augmentedrot_scale75", 2)):
        out["h5_files/main_split/" + name] = dict(
            data=rng.standard_normal((n, P, 3)).astype("float32"), label=rng.integers(0, 15, (n,)).astype("int32"))
    return out


def write_shards(root, shards, also_h5_names=False):
    for rel, arrays in shards.items():
        path = os.path.join(root, rel + ".npz")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez(path, **arrays)
        if also_h5_names:                      # the reference globs for '*.h5'
            open(os.path.join(root, rel + ".h5"), "wb").close()


# (class name, constructor kwargs, item indices, seed)
CASES = [
    ("ModelNet40", dict(num_points=64, partition="test"), [0, 3, 5], 11),
    ("ModelNet40", dict(num_points=48, partition="train"), [0, 2], 12),
    ("ShapeNetPart", dict(num_points=96, partition="test"), [0, 4], 13),
    ("ShapeNetPart", dict(num_points=80, partition="trainval"), [1, 3, 4], 14),
    ("ShapeNetPart", dict(num_points=96, partition="test", class_choice="cap"), [0], 15),
    ("ScanObjectNNCls", dict(num_points=64, partition="test"), [0, 2], 16),
    ("ScanObjectNNCls", dict(num_points=64, partition="train"), [1], 17),
    ("ScanObjectNNCls", dict(num_points=32, partition="test", subset="hard"), [1], 18),
]


def run_cases(mod, root):
    """Items of every case as a flat dict of arrays; multi-shard partitions are reported in the sorted-shard order."""
    res = {}
    for ci, (cls, kw, items, seed) in enumerate(CASES):
        with contextlib.redirect_stdout(io.StringIO()):
            ds = getattr(mod, cls)(data_dir=root, **kw)
        res["c%d_len" % ci] = np.int64(len(ds))
        if cls == "ShapeNetPart":
            res["c%d_seg" % ci] = np.array([ds.seg_num_all, ds.seg_start_index], dtype=np.int64)
        np.random.seed(seed)
        for it in items:
            for j, a in enumerate(ds[it]):
                a = np.asarray(a)
                res["c%d_i%d_%d" % (ci, it, j)] = a
                res["c%d_i%d_%d_dtype" % (ci, it, j)] = np.array(str(a.dtype))
    return res


def main():
    sys.modules["h5py"] = types.SimpleNamespace(File=_File)
    sys.path.insert(0, "/root/reference")
    import data as ref_data                      # /root/reference/data.py, unmodified
    assert ref_data.__file__.startswith("/root/reference/")
    # the reference concatenates shards in glob order (file-system order); svnet_b200.data sorts them.  Pin the order here.
    real_glob = ref_data.glob.glob
    ref_data.glob = types.SimpleNamespace(glob=lambda p: sorted(real_glob(p)))
    shards = shard_arrays()
    with tempfile.TemporaryDirectory() as root:
        write_shards(root, shards, also_h5_names=True)
        res = run_cases(ref_data, root)
    for rel, arrays in shards.items():           # the inputs travel with the outputs
        for k, a in arrays.items():
            res["shard|%s|%s" % (rel, k)] = a
    res["pc_normalize_in"] = shards["modelnet40_ply_hdf5_2048/ply_data_test0"]["data"][1].astype("float64")
    res["pc_normalize_out"] = ref_data.pc_normalize(res["pc_normalize_in"])
    np.random.seed(5)
    res["translate_out"] = ref_data.translate_pointcloud(shards["modelnet40_ply_hdf5_2048/ply_data_test0"]["data"][0])
    path = os.path.join(HERE, "data_readers.npz")
    np.savez_compressed(path, **res)
    print("wrote data_readers.npz %.1f KiB, %d arrays" % (os.path.getsize(path) / 1024, len(res)))


if __name__ == "__main__":
    main()
