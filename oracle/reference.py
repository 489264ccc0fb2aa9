"""Loader for the UNMODIFIED reference model code staged by oracle/make_ref.sh under oracle/_ref/
(git-ignored; travels to the GPU box with the snapshot).  TEST INFRASTRUCTURE ONLY -- imported by
tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs, never by svnet_b200/.

    ref = oracle.reference.load()          # the reference's `models` package (models/__init__.py:1-16)
    net = ref.SV_DGCNN_CLS(args, 40)       # models/sv_dgcnn_cls.py:22-82, stock PyTorch eager

The package is imported under the private name ``svnet_reference_models`` so that it cannot collide
with anything called ``models`` on sys.path.
"""
import contextlib
import importlib.util
import io
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
NAME = "svnet_reference_models"


def available():
    return os.path.exists(os.path.join(REF_DIR, "models", "__init__.py"))


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    if not available():
        raise RuntimeError("oracle/_ref is empty: run `sh oracle/make_ref.sh` where /root/reference exists "
                           "(__graft_entry__.build() does)")
    pkg_dir = os.path.join(REF_DIR, "models")
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    dont = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(mod)
    except Exception:
        del sys.modules[NAME]
        raise
    finally:
        sys.dont_write_bytecode = dont
    return mod


def submodule(name):
    """e.g. submodule('utils.sv_util'), submodule('sv_layers')"""
    load()
    return importlib.import_module(NAME + "." + name)


class Recorder:
    """Records, during one reference forward: the kNN indices of every `knn` call
    (models/utils/sv_util.py:19-25), the per-point pooled (s, v) of every svpool(dim=2) in the model
    file (sv_util.py:118-132), and the input of every binarised-activation Linear named in ``sign_of``
    (models/sv_layers.py:36-39) as its sign plane sign(x + beta) together with min |x + beta| stats."""

    def __init__(self, model_module, model=None, sign_of=()):
        self.mod, self.model, self.sign_of = model_module, model, tuple(sign_of)
        self.idx, self.pools, self.knn_in, self.signs = [], [], [], {}

    def __enter__(self):
        import torch
        util = submodule("utils.sv_util")
        self._util, self._knn, self._pool = util, util.knn, self.mod.svpool

        def knn(x, k):
            r = self._knn(x, k)
            self.idx.append(r)
            self.knn_in.append(x)
            return r

        def svpool(x, dim=2, keepdim=False, spool="max"):
            r = self._pool(x, dim=dim, keepdim=keepdim, spool=spool)
            if dim == 2:
                self.pools.append(r)
            return r

        util.knn = knn
        self.mod.svpool = svpool
        self._hooks = []
        if self.model is not None:
            mods = dict(self.model.named_modules())
            for name in self.sign_of:
                lin = mods[name]

                def hook(m, inp, name=name):
                    t = inp[0] + m.beta
                    self.signs[name] = (torch.sign(t).to(torch.int8), t.abs())
                self._hooks.append(lin.register_forward_pre_hook(hook))
        return self

    def __exit__(self, *a):
        self._util.knn = self._knn
        self.mod.svpool = self._pool
        for h in self._hooks:
            h.remove()
