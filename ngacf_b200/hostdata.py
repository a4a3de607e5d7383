"""Host data layer of run_Gowalla.py (reference: run_Gowalla.py:41-94 prepareData, data/loadGowalla.py:19-45):
loads a dataset into (userNum, itemNum, rt edges, train rows, test rows) WITHOUT building the per-user python
sets of loadGowalla.py:63-67 -- the CSR arrays of ngacf_b200.data.Interactions replace them."""
from __future__ import annotations

import os

import numpy as np

SYNTH_SHAPES = {"synth-gowalla": (29858, 40981, 1027370), "synth-yelp2018": (31668, 38048, 1561406),
                "synth-amazon-book": (52643, 91599, 2984108), "synth-ml100k": (943, 1682, 100000), "synth-tiny": (2000, 3000, 60000)}


def _find(data_root, rel):
    for root in (data_root, os.environ.get("NGACF_DATA_ROOT"), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")):
        if root and os.path.exists(os.path.join(root, rel)):
            return os.path.join(root, rel)
    raise FileNotFoundError("%s not found under --data_root / $NGACF_DATA_ROOT / ./data (the reference ships data/1K/u.data only; "
                            "use --dataset synth-* for synthetic shapes)" % rel)


def synth_bipartite(U, I, E, seed=0):
    """Power-law bipartite graph (same generator as the benchmark's)."""
    rng = np.random.default_rng(seed)
    cu = np.cumsum(np.arange(1, U + 1, dtype=np.float64) ** -0.6)
    ci = np.cumsum(np.arange(1, I + 1, dtype=np.float64) ** -0.8)
    cu /= cu[-1]
    ci /= ci[-1]
    pu, pi = rng.permutation(U).astype(np.int64), rng.permutation(I).astype(np.int64)
    draw_i = lambda m: pi[np.minimum(np.searchsorted(ci, rng.random(m)), I - 1)]
    draw_u = lambda m: pu[np.minimum(np.searchsorted(cu, rng.random(m)), U - 1)]
    key = np.unique(np.arange(U, dtype=np.int64) * I + draw_i(U))
    while key.shape[0] < E:
        m = int((E - key.shape[0]) * 1.3) + 1024
        new = np.setdiff1d(np.unique(draw_u(m) * I + draw_i(m)), key, assume_unique=True)
        if new.shape[0] > E - key.shape[0]:
            new = rng.choice(new, E - key.shape[0], replace=False)
        key = np.union1d(key, new)
    return key // I, key % I


def split_per_user(u, i, U, seed=1, test_frac=0.2):
    rng = np.random.default_rng(seed)
    order = np.lexsort((rng.random(u.shape[0]), u))
    u, i = u[order], i[order]
    ptr = np.searchsorted(u, np.arange(U + 1))
    deg = np.diff(ptr)
    pos = np.arange(u.shape[0]) - ptr[u]
    ntest = np.minimum(np.floor(deg * test_frac).astype(np.int64), deg - 1)
    is_test = pos < ntest[u]
    return (u[~is_test], i[~is_test]), (u[is_test], i[is_test])


def load_dataset(name, data_root=None, train_mode="PairSampling"):
    """-> dict(userNum, itemNum, rt_u, rt_i, train_u, train_i, test_u, test_i).  Mirrors prepareData's branches
    (run_Gowalla.py:44-80): Gowalla/Yelp come pre-split as CSVs; ml100k/ml1m are split 80/20 by sklearn's
    train_test_split, which draws from numpy's global RNG (seeded with --seed at run_Gowalla.py:193)."""
    import pandas as pd
    if name in SYNTH_SHAPES:
        U, I, E = SYNTH_SHAPES[name]
        u, i = synth_bipartite(U, I, E, 0)
        if train_mode == "NegSampling":     # leave-one-out like split_loo (loadGowalla.py:307-313): one held-out row per user
            order = np.lexsort((np.random.default_rng(1).random(u.shape[0]), u))
            u, i = u[order], i[order]
            first = np.r_[True, u[1:] != u[:-1]]
            (tu, ti), (su, si) = (u[~first], i[~first]), (u[first], i[first])
        else:
            (tu, ti), (su, si) = split_per_user(u, i, U, 1)
    elif name in ("Gowalla", "Yelp"):
        sub, pre = ("Gowalla", "g") if name == "Gowalla" else ("Yelp", "y")
        cols = dict(names=["userId", "itemId", "rating"], dtype={"userId": np.int64, "itemId": np.int64})
        tr = pd.read_csv(_find(data_root, "%s/%s_train.csv" % (sub, pre)), **cols)
        te = pd.read_csv(_find(data_root, "%s/%s_test.csv" % (sub, pre)), **cols)
        tu, ti, su, si = tr["userId"].values, tr["itemId"].values, te["userId"].values, te["itemId"].values
        U = int(max(tu.max(), su.max())) + 1
        I = int(max(ti.max(), si.max())) + 1
    elif name in ("ml100k", "ml1m"):
        from sklearn.model_selection import train_test_split
        if name == "ml100k":
            rt = pd.read_table(_find(data_root, "1K/u.data"), sep="\t", names=["userId", "itemId", "rating", "timestamp"])
        else:
            rt = pd.read_table(_find(data_root, "1M/ratings.dat"), sep="::", names=["userId", "itemId", "rating", "timestamp"], engine="python")
        U, I = int(rt["userId"].max()), int(rt["itemId"].max())     # ids start at 1 (run_Gowalla.py:60-63,72-76)
        rt["userId"] -= 1
        rt["itemId"] -= 1
        if train_mode == "NegSampling":     # split_loo (loadGowalla.py:307-313): the latest rating of every user is the test row
            rt["rank_latest"] = rt.groupby(["userId"])["timestamp"].rank(method="first", ascending=False)
            tr, te = rt[rt["rank_latest"] > 1], rt[rt["rank_latest"] == 1]
        else:
            tr, te = train_test_split(rt, test_size=0.2)
        tu, ti, su, si = tr["userId"].values, tr["itemId"].values, te["userId"].values, te["itemId"].values
    else:
        raise ValueError("unknown dataset %r" % name)
    return dict(userNum=U, itemNum=I, rt_u=np.concatenate([tu, su]), rt_i=np.concatenate([ti, si]),
                train_u=np.asarray(tu), train_i=np.asarray(ti), test_u=np.asarray(su), test_i=np.asarray(si))
