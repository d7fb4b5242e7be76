"""Minimal stand-in for the third-party ``hnswlib`` package (not installable here: no network), used ONLY by
make_golden.py so that the reference's own ``vectordb_optimized.Collection`` can be imported and its
``brute_force_search`` run unmodified.  It implements the eight ``Index`` methods that module
calls -- init_index / add_items / get_items / knn_query / mark_deleted / set_ef / save_index / load_index -- as a
plain label -> vector store.  ``brute_force_search`` only ever calls ``get_items`` (through ``_rebuild_cache``,
vectordb_optimized.py:246-269), which returns exactly the float32 rows that were inserted, as the real library does;
``knn_query`` here is an exact scan with hnswlib's own distance conventions and is never used for golden data."""
import pickle

import numpy as np


class Index:
    def __init__(self, space="cosine", dim=0):
        self.space, self.dim = space, int(dim)
        self._rows, self._deleted, self._ef = {}, set(), 10

    def init_index(self, max_elements=0, ef_construction=200, M=16, random_seed=100, allow_replace_deleted=False):
        self.max_elements = max_elements

    def set_ef(self, ef):
        self._ef = ef

    def set_num_threads(self, n):
        pass

    def add_items(self, data, ids=None, num_threads=-1, replace_deleted=False):
        data = np.asarray(data, dtype=np.float32).reshape(-1, self.dim)
        ids = list(range(len(self._rows), len(self._rows) + len(data))) if ids is None else list(np.asarray(ids).reshape(-1))
        for lab, row in zip(ids, data):
            self._rows[int(lab)] = row.copy()
            self._deleted.discard(int(lab))

    def get_items(self, ids=None, return_type="numpy"):
        return np.stack([self._rows[int(i)] for i in ids]) if len(ids) else np.zeros((0, self.dim), np.float32)

    def mark_deleted(self, label):
        self._deleted.add(int(label))

    def get_current_count(self):
        return len(self._rows)

    def knn_query(self, data, k=1, num_threads=-1, filter=None):
        data = np.asarray(data, dtype=np.float32).reshape(-1, self.dim)
        labs = np.array([l for l in self._rows if l not in self._deleted], dtype=np.int64)
        mat = np.stack([self._rows[int(l)] for l in labs])
        out_l, out_d = [], []
        for q in data:
            if self.space == "l2":
                d = ((mat - q) ** 2).sum(1)
            elif self.space == "ip":
                d = 1.0 - mat @ q
            else:
                d = 1.0 - (mat @ q) / (np.linalg.norm(mat, axis=1) * np.linalg.norm(q) + 1e-30)
            order = np.argsort(d, kind="stable")[:k]
            out_l.append(labs[order]); out_d.append(d[order])
        return np.stack(out_l), np.stack(out_d).astype(np.float32)

    def save_index(self, path):
        with open(path, "wb") as f:
            pickle.dump((self.space, self.dim, self._rows, self._deleted), f)

    def load_index(self, path, max_elements=0, allow_replace_deleted=False):
        with open(path, "rb") as f:
            self.space, self.dim, self._rows, self._deleted = pickle.load(f)
