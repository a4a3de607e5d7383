"""TEST INFRASTRUCTURE ONLY -- live oracle: the UNMODIFIED reference, imported from /root/reference.

This module never copies reference code.  It puts ``/root/reference`` on ``sys.path`` and applies the
monkey-patch shims listed in SURVEY.md section 8c so that the reference's own ``SPUIGACF``
(graphattention/SPUIGACF.py:5-52), ``BPRLoss`` (graphattention/BPRLoss.py:4-9), ``train_bpr``
(train_eval_Gowalla.py:90-144) and ``eval_neg_all`` (train_eval_Gowalla.py:274-354) run on a CPU-only
host with torch 2.11 / numpy 2 / pandas 3 / Python 3.12.

It exists only in the build container (``/root/reference`` does not travel to the GPU box): it is used by
``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/`` and by the
``not gpu`` tests that pin ``oracle/port.py`` against the reference (they skip when the reference is
absent).  Nothing in ``ngacf_b200/`` imports it.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import random
import sys

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("NGACF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "graphattention", "SPUIGACF.py"))


_PATCHED = False
_ORIG = {}


def _install_shims():
    """Idempotent process-wide shims."""
    global _PATCHED
    if _PATCHED:
        return
    _PATCHED = True

    # numpy 2 removed asfarray (metrics.py:52,77)
    if not hasattr(np, "asfarray"):
        np.asfarray = lambda a, dtype=float: np.asarray(a, dtype=dtype)  # type: ignore[attr-defined]

    # device='cuda' hard-coded in the layer (SPUIGACF.py:342,367,370)
    if not torch.cuda.is_available():
        for name in ("ones", "zeros"):
            orig = getattr(torch, name)
            _ORIG[name] = orig

            def make(orig):
                def shim(*a, **kw):
                    if kw.get("device", None) == "cuda":
                        kw = dict(kw)
                        kw["device"] = "cpu"
                    return orig(*a, **kw)
                return shim
            setattr(torch, name, make(orig))

        # .cuda() everywhere in the loops (train_eval_Gowalla.py:106,126-128,323-324)
        torch.Tensor.cuda = lambda self, *a, **k: self  # type: ignore[assignment]
        torch.nn.Module.cuda = lambda self, *a, **k: self  # type: ignore[assignment]
        torch.cuda.device_count = lambda: 1  # type: ignore[assignment]
        torch.cuda.empty_cache = lambda: None  # type: ignore[assignment]
        torch.cuda.manual_seed_all = lambda s: None  # type: ignore[assignment]

    # loss.backward(torch.ones(ndev)) on a 0-dim loss (train_eval_Gowalla.py:137)
    orig_backward = torch.Tensor.backward
    _ORIG["backward"] = orig_backward

    def backward(self, gradient=None, *a, **k):
        if gradient is not None and gradient.shape != self.shape and gradient.numel() == self.numel():
            gradient = gradient.reshape(self.shape)
        return orig_backward(self, gradient, *a, **k)
    torch.Tensor.backward = backward  # type: ignore[assignment]

    # random.sample(set, k) raises TypeError on Python >= 3.11 (loadGowalla.py:75-76)
    orig_sample = random.sample
    _ORIG["sample"] = orig_sample

    def sample(population, k, **kw):
        if isinstance(population, (set, frozenset)):
            population = sorted(population)
        return orig_sample(population, k, **kw)
    random.sample = sample  # type: ignore[assignment]


class _SerialPool:
    def __init__(self, n=None):
        pass

    def map(self, fn, it):
        return [fn(x) for x in it]

    def close(self):
        pass


def load():
    """Import the reference modules (unmodified) and return them in a namespace dict."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    _install_shims()
    # The repo ships its own drop-in packages called `graphattention`, `data`, `train_eval_Gowalla`;
    # the reference ones must win inside this harness, so they are imported under the reference root
    # being FIRST on sys.path and then pinned under private aliases.
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k == "graphattention" or k.startswith("graphattention.")
             or k == "data" or k.startswith("data.")
             or k in ("train_eval_Gowalla", "parallel")}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        ns = {
            "SPUIGACF": importlib.import_module("graphattention.SPUIGACF"),
            "SPGA": importlib.import_module("graphattention.SPGA"),
            "BPRLoss": importlib.import_module("graphattention.BPRLoss"),
            "metrics": importlib.import_module("graphattention.metrics"),
            "loadGowalla": importlib.import_module("data.loadGowalla"),
            "train_eval": importlib.import_module("train_eval_Gowalla"),
        }
        for mod in ns.values():
            assert os.path.abspath(mod.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), mod.__file__
        # eval_neg_all pickles report_one_user to a process pool by module name
        # (train_eval_Gowalla.py:282,341); a forked worker would re-import the name and find the
        # repo's own train_eval_Gowalla, so the harness maps the pool to a serial one.
        import types
        real_mp = ns["train_eval"].multiprocessing
        ns["train_eval"].multiprocessing = types.SimpleNamespace(
            Pool=_SerialPool, cpu_count=real_mp.cpu_count)
    finally:
        sys.path.remove(REFERENCE_ROOT)
        # drop the reference copies from sys.modules again and restore whatever was there, so the
        # repo's own same-named packages stay importable by the caller
        for k in list(sys.modules):
            if (k == "graphattention" or k.startswith("graphattention.") or k == "data"
                    or k.startswith("data.") or k in ("train_eval_Gowalla", "parallel")):
                del sys.modules[k]
        sys.modules.update(saved)
    return ns


def make_model(ns, userNum, itemNum, droprate=0.0, seed=2019, dtype=torch.float32):
    """Reference SPUIGACF on CPU (useCuda=False), seeded like run_Gowalla.py:191."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        torch.manual_seed(seed)
        model = ns["SPUIGACF"].SPUIGACF(userNum, itemNum, 64, [64, 64], droprate, useCuda=False)
    finally:
        torch.set_default_dtype(old)
    return model


def coalesced_indices(ns, rt, userNum, itemNum):
    """ui_mat indices exactly as run_Gowalla.py:82,94 produces them, bypassing the npz cache write
    of get_adj_mat (loadGowalla.py:245 writes INTO the reference data dir)."""
    L = ns["loadGowalla"].buildLaplacianMat(rt, userNum, itemNum, "ui_mat")
    idx = torch.from_numpy(np.stack([np.asarray(L.row), np.asarray(L.col)]).astype(np.int64))
    val = torch.from_numpy(np.asarray(L.data, dtype=np.float32))
    sp = torch.sparse_coo_tensor(idx, val, (userNum, itemNum)).coalesce()
    return sp.indices()


@contextlib.contextmanager
def injected_dropout(masks, droprate):
    """Replace torch's dropout inside the reference with caller-supplied keep masks.

    ``masks`` is an iterator of boolean keep tensors consumed in the reference's call order
    (SPUIGACF.py:208 features, :375 x8 edge, :213 features, :375 edge).  Scaling is 1/(1-p) as
    torch does.  This is how dropout parity is pinned without sharing an RNG stream.
    """
    import torch.nn.functional as F
    it = iter(masks)
    scale = 1.0 / (1.0 - droprate) if droprate > 0 else 1.0
    orig_f = F.dropout
    orig_m = torch.nn.Dropout.forward

    def fdrop(x, p=0.5, training=True, inplace=False):
        if not training or p == 0:
            return x
        keep = next(it).to(x.device).reshape(x.shape)
        return x * keep.to(x.dtype) * scale

    def mdrop(self, x):
        if not self.training or self.p == 0:
            return x
        keep = next(it).to(x.device).reshape(x.shape)
        return x * keep.to(x.dtype) * scale

    F.dropout = fdrop
    torch.nn.Dropout.forward = mdrop
    try:
        yield
    finally:
        F.dropout = orig_f
        torch.nn.Dropout.forward = orig_m


@contextlib.contextmanager
def captured_dropout():
    """Record the keep masks torch's own RNG produced inside the reference, in call order."""
    import torch.nn.functional as F
    rec = []
    orig_f = F.dropout
    orig_m = torch.nn.Dropout.forward

    def fdrop(x, p=0.5, training=True, inplace=False):
        y = orig_f(x, p, training, inplace)
        if training and p > 0:
            rec.append((y != 0) | (x == 0))
        return y

    def mdrop(self, x):
        y = orig_m(self, x)
        if self.training and self.p > 0:
            rec.append((y != 0) | (x == 0))
        return y

    F.dropout = fdrop
    torch.nn.Dropout.forward = mdrop
    try:
        yield rec
    finally:
        F.dropout = orig_f
        torch.nn.Dropout.forward = orig_m


@contextlib.contextmanager
def default_dtype(dtype):
    """fp64 oracle runs: the layer's torch.ones(...) takes the default dtype (SPUIGACF.py:367,370)."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        yield
    finally:
        torch.set_default_dtype(old)
