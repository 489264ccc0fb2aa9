"""Shared helpers for the test-suite: golden fixtures, synthetic state dicts, comparisons."""
import contextlib
import io
import json
import os

import numpy as np
import torch

from svnet_b200.synthetic import make_args, state_dict_digest, synthetic_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_state_dict(g):
    """Rebuild the synthetic state_dict a model fixture was generated with; verify its digest."""
    shapes = json.loads(str(g["sd_shapes"]))
    tmpl = {k: torch.empty(s, dtype=getattr(torch, d.split(".")[1])) for k, (s, d) in shapes.items()}
    sd = synthetic_state_dict(tmpl, seed=int(g["seed"]), beta_zero=bool(g["beta_zero"]))
    assert state_dict_digest(sd) == str(g["digest"]), "synthetic state_dict generator drifted from the fixtures"
    return sd


def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def max_abs(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


def assert_close(a, b, rtol=1e-3, atol=1e-4, what=""):
    """north_star tolerance: 1e-3 relative / 1e-4 absolute at fp32 accumulation."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    bad = err > tol
    assert not bad.any(), "%s: %d/%d elements out of tolerance, max abs err %.3g" % (what, bad.sum(), bad.size, err.max())


def knn_is_valid(idx, pd, k):
    """idx (B,N,k) is a correct top-k of the score matrix pd (B,N,N) up to the order of exact ties:
    scores along idx are non-increasing and every selected score >= every unselected one."""
    B, N, _ = pd.shape
    sel = np.take_along_axis(pd, idx, axis=2)
    if not (np.diff(sel, axis=2) <= 0).all():
        return False
    kth = sel[:, :, -1]
    mask = np.ones_like(pd, dtype=bool)
    np.put_along_axis(mask, idx, False, axis=2)
    rest_max = np.where(mask, pd, -np.inf).max(axis=2)
    uniq = all(len(set(r)) == k for r in idx.reshape(-1, k))
    return bool((rest_max <= kth).all() and uniq)


def t2n(t):
    return t.detach().cpu().numpy()
