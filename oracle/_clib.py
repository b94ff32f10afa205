"""ctypes binding of oracle/csrc/oracle.c (test infrastructure only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "csrc", "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


class _Lib:
    def __init__(self):
        self._l = None

    def _load(self):
        if self._l is None:
            build()
            l = ctypes.CDLL(_SO)
            p, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
            l.oracle_scores_fma.argtypes = [p, p, i64, i64, i64, p]
            l.oracle_pair_scores_fma.argtypes = [p, p, p, p, i64, i64, p]
            l.oracle_topk.argtypes = [p, i64, i64, p, p, i32, p, p]
            l.oracle_fullsort_topk.argtypes = [p, p, p, i64, i64, i64, p, p, i32, p, p]
            for f in (l.oracle_scores_fma, l.oracle_pair_scores_fma, l.oracle_topk, l.oracle_fullsort_topk):
                f.restype = None
            self._l = l
        return self._l

    @staticmethod
    def _f32(a):
        return np.ascontiguousarray(a, dtype=np.float32)

    @staticmethod
    def _i64(a):
        return np.ascontiguousarray(a, dtype=np.int64)

    def scores_fma(self, U, V):
        """[nu,d] x [ni,d] -> [nu,ni] canonical fp32 fma-chain scores."""
        l = self._load()
        U, V = self._f32(U), self._f32(V)
        out = np.empty((U.shape[0], V.shape[0]), dtype=np.float32)
        l.oracle_scores_fma(U.ctypes.data, V.ctypes.data, U.shape[0], V.shape[0], U.shape[1], out.ctypes.data)
        return out

    def pair_scores_fma(self, U, V, uid, iid):
        l = self._load()
        U, V, uid, iid = self._f32(U), self._f32(V), self._i64(uid), self._i64(iid)
        out = np.empty(len(uid), dtype=np.float32)
        l.oracle_pair_scores_fma(U.ctypes.data, V.ctypes.data, uid.ctypes.data, iid.ctypes.data, len(uid),
                                 U.shape[1], out.ctypes.data)
        return out

    def topk(self, scores, hist_indptr, hist_indices, k):
        l = self._load()
        scores = self._f32(scores)
        hp, hi = self._i64(hist_indptr), self._i64(hist_indices)
        nu, ni = scores.shape
        ids = np.empty((nu, k), dtype=np.int64)
        sc = np.empty((nu, k), dtype=np.float32)
        l.oracle_topk(scores.ctypes.data, nu, ni, hp.ctypes.data, hi.ctypes.data, k, ids.ctypes.data, sc.ctypes.data)
        return ids, sc

    def fullsort_topk(self, U, V, users, hist_indptr, hist_indices, k):
        l = self._load()
        U, V, users = self._f32(U), self._f32(V), self._i64(users)
        hp, hi = self._i64(hist_indptr), self._i64(hist_indices)
        nu = len(users)
        ids = np.empty((nu, k), dtype=np.int64)
        sc = np.empty((nu, k), dtype=np.float32)
        l.oracle_fullsort_topk(U.ctypes.data, V.ctypes.data, users.ctypes.data, nu, V.shape[0], U.shape[1],
                               hp.ctypes.data, hi.ctypes.data, k, ids.ctypes.data, sc.ctypes.data)
        return ids, sc


lib = _Lib()
