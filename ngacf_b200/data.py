"""Interaction data resident in HBM as CSR arrays -- replaces the reference's per-user Python sets
(data/loadGowalla.py:63-67,86-92: `positive_items`/`negative_items` sets, the 60-100 GB of host RAM of
README.md:19,25).  Negatives are never materialised: the sampler and the evaluator work against the
sorted train rows and the sorted item pool."""
from __future__ import annotations

import numpy as np
import torch


class Interactions:
    """train/test positives of every user + item pool, as device int32 tensors.

    train_rows_user : userId column of train_df in file order (the batching order of train_bpr,
                      train_eval_Gowalla.py:111-115)
    train_ptr/items : CSR of unique train items per user, sorted
    train_rank      : rank of each train item inside `pool`
    pool / in_pool  : sorted unique item ids of rt = train+test (item_pool, loadGowalla.py:64) / its bitmap
    test_ptr/items  : CSR of test items per user, sorted
    eval_users      : users present in train AND test (inner merge, loadGowalla.py:91)
    train_rows_item, test_rows_user/item : the (user, item) rows of train_df / test_df in file order (NegSampling, SampledNeg)
    all_ptr/all_rank: CSR of train+test items per user as pool ranks (positives_negtives, loadGowalla.py:56-60)
    n_train_users   : divisor of the metrics (train_eval_Gowalla.py:283)
    """

    def __init__(self, U, I, train_u, train_i, test_u, test_i, device, pool=None):
        self.U, self.I = int(U), int(I)
        train_u = np.asarray(train_u, np.int64)
        train_i = np.asarray(train_i, np.int64)
        test_u = np.asarray(test_u, np.int64)
        test_i = np.asarray(test_i, np.int64)
        for name, a, hi in (("train user", train_u, U), ("train item", train_i, I), ("test user", test_u, U), ("test item", test_i, I)):
            if a.size and (a.min() < 0 or a.max() >= hi):
                raise ValueError("%s id out of range" % name)
        pool = np.unique(np.concatenate([train_i, test_i])) if pool is None else np.unique(np.asarray(pool, np.int64))

        def csr(u, i):
            key = np.unique(u * np.int64(I) + i)
            uu, ii = key // I, key % I
            ptr = np.zeros(U + 1, np.int64)
            np.add.at(ptr, uu + 1, 1)
            return np.cumsum(ptr), ii
        tp, ti = csr(train_u, train_i)
        sp, si = csr(test_u, test_i)
        has_train = np.diff(tp) > 0
        has_test = np.diff(sp) > 0
        in_pool = np.zeros(I, np.uint8)
        in_pool[pool] = 1
        self.n_train_rows = int(train_u.shape[0])
        self.n_train_users = int(has_train.sum())
        self.host = dict(train_ptr=tp, train_items=ti, test_ptr=sp, test_items=si, pool=pool)
        t = lambda a, dt=torch.int32: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(device)
        self.train_rows_user = t(train_u)
        self.train_ptr, self.train_items = t(tp), t(ti)
        self.train_rank = t(np.searchsorted(pool, ti))
        self.pool, self.in_pool = t(pool), t(in_pool, torch.uint8)
        self.test_ptr, self.test_items = t(sp), t(si)
        self.eval_users = t(np.nonzero(has_train & has_test)[0])
        # NegSampling / SampledNeg: the rows' own items in file order, and the per-user union of train and test
        # items (positives_negtives, loadGowalla.py:56-60) as ranks in the pool -- the negatives are "pool minus this row"
        self.train_rows_item = t(train_i)
        self.test_rows_user, self.test_rows_item = t(test_u), t(test_i)
        self.n_test_rows = int(test_u.shape[0])
        ap, ai = csr(np.concatenate([train_u, test_u]), np.concatenate([train_i, test_i]))
        self.all_ptr, self.all_rank = t(ap), t(np.searchsorted(pool, ai))
        self.host.update(all_ptr=ap, all_items=ai)
        self.device = torch.device(device)

    def min_negatives(self) -> int:
        """Smallest number of negatives (item_pool minus a user's train and test items) over the users that have any row."""
        ap = self.host["all_ptr"]
        deg = np.diff(ap)
        deg = deg[deg > 0]
        return int(self.host["pool"].shape[0] - (deg.max() if deg.size else 0))

    def __len__(self):      # len(train_df) in train_bpr, len(test_pos_neg) in eval_neg_all
        return self.n_train_rows

    @classmethod
    def from_arrays(cls, U, I, train_u, train_i, test_u, test_i, device="cuda"):
        return cls(U, I, train_u, train_i, test_u, test_i, device)

    @classmethod
    def from_frames(cls, U, I, train_df, test_df, device="cuda"):
        """train_df/test_df: DataFrames with userId,itemId columns (loadGowalla.py:33-45)."""
        return cls(U, I, train_df["userId"].values, train_df["itemId"].values, test_df["userId"].values, test_df["itemId"].values, device)

    @classmethod
    def from_reference_frames(cls, U, I, train_df=None, train_pos_neg=None, test_pos=None, device="cuda"):
        """The reference's own structures: train_df rows (userId,itemId), train_pos_neg with python sets
        (loadGowalla.py:63-67) and test_df after test_positives (:86-88).  item_pool = positive | negative
        of any user (the reference builds negative_items = item_pool - positive_items, :66)."""
        first = train_pos_neg.iloc[0]
        pool = np.array(sorted(set(first["positive_items"]) | set(first["negative_items"])), np.int64)
        if train_df is not None:
            tu, ti = train_df["userId"].values, train_df["itemId"].values
        else:   # evaluation only needs the train SETS; row order is irrelevant there
            tu = np.concatenate([np.full(len(s), int(u), np.int64) for u, s in zip(train_pos_neg["userId"].values, train_pos_neg["positive_items"].values)])
            ti = np.concatenate([np.fromiter(s, np.int64, len(s)) for s in train_pos_neg["positive_items"].values])
        if test_pos is not None:
            su = np.concatenate([np.full(len(s), int(u), np.int64) for u, s in zip(test_pos["userId"].values, test_pos["positive_items"].values)])
            si = np.concatenate([np.fromiter(s, np.int64, len(s)) for s in test_pos["positive_items"].values])
        else:
            su = si = np.zeros(0, np.int64)
        return cls(U, I, tu, ti, su, si, device, pool=pool)

    @classmethod
    def from_negsampling_frames(cls, U, I, pos_neg, train_df=None, test_df=None, device="cuda"):
        """NegSampling / SampledNeg structures of the reference (run_Gowalla.py:89-92): pos_neg = positives_negtives(rt) with ALL
        positives of a user (train and test, loadGowalla.py:56-60), train_df / test_df = (userId, itemId) rows in file order.
        Whichever of the two frames is absent is reconstructed as "all positives minus the given rows" (its row order is not used)."""
        first = pos_neg.iloc[0]
        pool = np.array(sorted(set(first["positive_items"]) | set(first["negative_items"])), np.int64)
        au = np.concatenate([np.full(len(s), int(u), np.int64) for u, s in zip(pos_neg["userId"].values, pos_neg["positive_items"].values)])
        ai = np.concatenate([np.fromiter(s, np.int64, len(s)) for s in pos_neg["positive_items"].values])

        def rows(df):
            return np.asarray(df["userId"].values, np.int64), np.asarray(df["itemId"].values, np.int64)

        def minus(u, i):       # all positives not among the rows (u, i)
            key_all = au * np.int64(I) + ai
            rest = np.setdiff1d(key_all, u * np.int64(I) + i)
            return rest // I, rest % I
        if train_df is not None:
            tu, ti = rows(train_df)
            su, si = rows(test_df) if test_df is not None else minus(tu, ti)
        else:
            su, si = rows(test_df)
            tu, ti = minus(su, si)
        return cls(U, I, tu, ti, su, si, device, pool=pool)
