"""ctypes front-end of the CPU oracle (oracle/gmm_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, bench.py's cpu_baseline /
--impl reference legs and __graft_entry__.smoke(); never from the product.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_FP = C.POINTER(C.c_float)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])


class Oracle:
    """precision: 'f32' = reference arithmetic (the timed CPU port),
    'f64' = ground truth."""

    def __init__(self, precision="f64", clusters_struct=None):
        path = os.path.join(_HERE, f"libgmm_oracle_{precision}.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        self.precision = precision
        cs = clusters_struct
        P = C.POINTER(cs)
        L.oracle_transpose.argtypes = [_FP, C.c_int, C.c_int, _FP]
        L.oracle_invert.argtypes = [_FP, C.c_int, _FP, C.c_int]
        L.oracle_constants.argtypes = [P, C.c_int, C.c_int]
        L.oracle_seed.argtypes = [_FP, C.c_int, C.c_int, C.c_int, P]
        L.oracle_estep.argtypes = [_FP, C.c_int, C.c_int, C.c_int, P]
        L.oracle_estep.restype = C.c_float
        L.oracle_mstep.argtypes = [_FP, C.c_int, C.c_int, C.c_int, P]
        L.oracle_epsilon.argtypes = [C.c_int, C.c_int]
        L.oracle_epsilon.restype = C.c_float
        L.oracle_em.argtypes = [_FP, C.c_int, C.c_int, C.c_int, P, C.c_int, C.c_int, C.c_float,
                                C.POINTER(C.c_int)]
        L.oracle_em.restype = C.c_float
        L.oracle_rissanen.argtypes = [C.c_float, C.c_int, C.c_int, C.c_int]
        L.oracle_rissanen.restype = C.c_float
        L.oracle_reduce_order.argtypes = [P, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_reduce_order.restype = C.c_int
        L.oracle_fit.argtypes = [_FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P, P,
                                 C.POINTER(C.c_float)]
        L.oracle_fit.restype = C.c_int

    @staticmethod
    def _fp(a):
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(_FP)

    def transpose(self, aos):
        N, D = aos.shape
        soa = np.empty((D, N), np.float32)
        self.lib.oracle_transpose(self._fp(aos), N, D, self._fp(soa))
        return soa

    def invert(self, m, use_log10=False):
        a = np.ascontiguousarray(m, np.float32).copy()
        ld = C.c_float()
        self.lib.oracle_invert(self._fp(a), a.shape[0], C.byref(ld), int(use_log10))
        return a, ld.value

    def seed(self, aos, K, cl):
        N, D = aos.shape
        s = cl.struct()
        self.lib.oracle_seed(self._fp(aos), N, D, K, C.byref(s))

    def constants(self, cl, K):
        s = cl.struct()
        self.lib.oracle_constants(C.byref(s), K, cl.D)

    def estep(self, soa, cl, K):
        D, N = soa.shape
        s = cl.struct()
        return float(self.lib.oracle_estep(self._fp(soa), N, D, K, C.byref(s)))

    def mstep(self, soa, cl, K):
        D, N = soa.shape
        s = cl.struct()
        self.lib.oracle_mstep(self._fp(soa), N, D, K, C.byref(s))

    def epsilon(self, D, N):
        return float(self.lib.oracle_epsilon(D, N))

    def em(self, soa, cl, K, min_iters, max_iters, epsilon=None):
        D, N = soa.shape
        if epsilon is None:
            epsilon = self.epsilon(D, N)
        it = C.c_int()
        s = cl.struct()
        ll = self.lib.oracle_em(self._fp(soa), N, D, K, C.byref(s), min_iters, max_iters, epsilon, C.byref(it))
        return float(ll), it.value

    def rissanen(self, ll, K, D, N):
        return float(self.lib.oracle_rissanen(ll, K, D, N))

    def reduce_order(self, cl, K):
        c1, c2 = C.c_int(), C.c_int()
        s = cl.struct()
        newK = self.lib.oracle_reduce_order(C.byref(s), K, cl.D, C.byref(c1), C.byref(c2))
        return newK, (c1.value, c2.value)

    def fit(self, aos, K0, target_K, min_iters, max_iters, cl, saved):
        N, D = aos.shape
        mr = C.c_float()
        s, sv = cl.struct(), saved.struct()
        ideal = self.lib.oracle_fit(self._fp(aos), N, D, K0, target_K, min_iters, max_iters,
                                    C.byref(s), C.byref(sv), C.byref(mr))
        return ideal, mr.value
