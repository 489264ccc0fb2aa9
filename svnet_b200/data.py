"""Caller-side dataset readers of the eval loops, without h5py.

The reference reads ModelNet40 / ShapeNetPart / ScanObjectNN from HDF5 (data.py:71-115, 186-201,
260-340); h5py is not installable here, so the same arrays are read from ``.npz`` files that sit where
the ``.h5`` files sit and hold the same keys (``data``, ``label``, ``pid``): ``convert_h5_to_npz`` makes
them once on a machine that has h5py.  Where h5py is installed and no ``.npz`` shard matches, the ``.h5``
files themselves are read.  Class names, constructor arguments, item tuples and partition
behaviour follow the reference, so ``DataLoader(ModelNet40(partition='test', num_points=1024,
data_dir=...))`` feeds ``model(data.permute(0, 2, 1))`` exactly as main_cls_dgcnn.py:218-238 does.
"""
import glob
import os

import numpy as np
from torch.utils.data import Dataset


def pc_normalize(pc):
    """Centre on the centroid, scale to the unit sphere (data.py:15-20)."""
    pc = pc - np.mean(pc, axis=0)
    return pc / np.max(np.sqrt(np.sum(pc ** 2, axis=1)))


def translate_pointcloud(pointcloud):
    """Train-time anisotropic scale + shift (data.py:163-168)."""
    xyz1 = np.random.uniform(low=2. / 3., high=3. / 2., size=[3])
    xyz2 = np.random.uniform(low=-0.2, high=0.2, size=[3])
    return np.add(np.multiply(pointcloud, xyz1), xyz2).astype('float32')


def convert_h5_to_npz(h5_path, npz_path=None):
    """One-time conversion next to the source file: every top-level HDF5 dataset becomes an npz array."""
    try:
        import h5py
    except ImportError as e:          # loud: there is no silent fallback format
        raise RuntimeError("convert_h5_to_npz needs h5py; run it where h5py is installed") from e
    npz_path = npz_path or os.path.splitext(h5_path)[0] + ".npz"
    with h5py.File(h5_path, "r") as f:
        np.savez(npz_path, **{k: np.asarray(f[k]) for k in f.keys()})
    return npz_path


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError:
        return None


def _shards(pattern):
    """Shards matching ``pattern`` (written with a ``.npz`` suffix): the npz files if there are any, else -- where
    h5py is installed -- the reference's own ``.h5`` files under the same names (data.py:78, 99-102, 318)."""
    files = glob.glob(pattern)
    if not files and _h5py() is not None:
        files = glob.glob(pattern[:-4] + ".h5")
    return files


def _load(files, keys):
    if not files:
        raise FileNotFoundError("no .npz shards found (convert the .h5 files with svnet_b200.data.convert_h5_to_npz, "
                                "or install h5py to read them directly)")
    cols = [[] for _ in keys]
    for name in sorted(files):
        if name.endswith(".h5"):
            f = _h5py().File(name, "r")
            try:
                for c, (k, dt) in zip(cols, keys):
                    c.append(np.asarray(f[k]).astype(dt))
            finally:
                f.close()
        else:
            with np.load(name) as z:
                for c, (k, dt) in zip(cols, keys):
                    c.append(z[k].astype(dt))
    return tuple(np.concatenate(c, axis=0) for c in cols)


def load_data_cls(data_dir, partition):
    """data.py:71-88: every ``modelnet40*hdf5_2048/*<partition>*`` shard, concatenated."""
    return _load(_shards(os.path.join(data_dir, 'modelnet40*hdf5_2048', '*%s*.npz' % partition)),
                 [("data", "float32"), ("label", "int64")])


def load_data_partseg(data_dir, partition):
    """data.py:91-115: 'trainval' joins the train and val shards."""
    root = os.path.join(data_dir, 'shapenet*hdf5*')
    if partition == 'trainval':
        files = _shards(os.path.join(root, '*train*.npz')) + _shards(os.path.join(root, '*val*.npz'))
    else:
        files = _shards(os.path.join(root, '*%s*.npz' % partition))
    return _load(files, [("data", "float32"), ("label", "int64"), ("pid", "int64")])


class ModelNet40(Dataset):
    """data.py:186-201: item = (first ``num_points`` points, label); 'train' items are rescaled and
    their points shuffled, every other partition is returned as stored."""

    def __init__(self, num_points, data_dir, partition='train', **kwargs):
        self.num_points, self.partition = num_points, partition
        self.data, self.label = load_data_cls(data_dir, partition)

    def __len__(self):
        return len(self.data)

    def __getitem__(self, item):
        cloud = self.data[item, :self.num_points]
        if self.partition == 'train':
            cloud = translate_pointcloud(cloud)[np.random.permutation(len(cloud))]
        return cloud, self.label[item]


class ShapeNetPart(Dataset):
    """data.py:260-300: (points, category label, per-point part id); optional single category."""
    cat2id = {'airplane': 0, 'bag': 1, 'cap': 2, 'car': 3, 'chair': 4, 'earphone': 5, 'guitar': 6, 'knife': 7,
              'lamp': 8, 'laptop': 9, 'motor': 10, 'mug': 11, 'pistol': 12, 'rocket': 13, 'skateboard': 14, 'table': 15}
    seg_num = [4, 2, 2, 4, 4, 3, 3, 2, 4, 2, 6, 2, 3, 3, 3, 3]
    index_start = [0, 4, 6, 8, 12, 16, 19, 22, 24, 28, 30, 36, 38, 41, 44, 47]

    def __init__(self, num_points, data_dir, partition='train', class_choice=None):
        self.data, self.label, self.seg = load_data_partseg(data_dir, partition)
        self.num_points = num_points
        self.partition = partition
        self.class_choice = class_choice
        if class_choice is not None:
            cid = self.cat2id[class_choice]
            keep = (self.label == cid).reshape(-1)
            self.data, self.label, self.seg = self.data[keep], self.label[keep], self.seg[keep]
            self.seg_num_all = self.seg_num[cid]
            self.seg_start_index = self.index_start[cid]
        else:
            self.seg_num_all = 50
            self.seg_start_index = 0

    def __getitem__(self, item):
        pointcloud = self.data[item][:self.num_points]
        label = self.label[item]
        seg = self.seg[item][:self.num_points]
        if self.partition == 'trainval':
            order = np.random.permutation(pointcloud.shape[0])
            pointcloud, seg = pointcloud[order], seg[order]
        return pointcloud, label, seg

    def __len__(self):
        return self.data.shape[0]


class ScanObjectNNCls(Dataset):
    """data.py:302-340: ``num_points`` of the 2048 points drawn in random order on every access
    (test items too, as in the reference); train items are rescaled."""
    _files = {('train', 'easy'): 'training_objectdataset', ('train', 'hard'): 'training_objectdataset_augmentedrot_scale75',
              ('test', 'easy'): 'test_objectdataset', ('test', 'hard'): 'test_objectdataset_augmentedrot_scale75'}

    def __init__(self, num_points, data_dir, partition='train', subset='easy'):
        super().__init__()
        if partition not in ('train', 'test'):
            raise ValueError('not recognized partition {}'.format(partition))
        name = self._files[(partition, 'easy' if subset == 'easy' else 'hard')]
        self.points, self.labels = _load(_shards(os.path.join(data_dir, 'h5_files', 'main_split', name + '.npz')),
                                         [("data", "float32"), ("label", "int64")])
        self.num_points = num_points
        self.partition = partition

    def __getitem__(self, idx):
        pt_idxs = np.arange(0, self.points.shape[1])
        np.random.shuffle(pt_idxs)
        pointcloud = self.points[idx, pt_idxs[:self.num_points]].copy()
        if self.partition == 'train':
            pointcloud = translate_pointcloud(pointcloud)
        return pointcloud, self.labels[idx]

    def __len__(self):
        return self.points.shape[0]
