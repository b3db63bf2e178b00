"""clusters_t mirror for ctypes callers.

Layout is the reference's ``clusters_t`` (gaussian.h:62-76): a struct of eight
``float*``; memberships are cluster-major ``[K][N]``.  ``Clusters`` owns the
numpy arrays and hands out the C struct that points into them.
"""
import ctypes as C
import numpy as np

_FP = C.POINTER(C.c_float)


class clusters_t(C.Structure):
    _fields_ = [("N", _FP), ("pi", _FP), ("constant", _FP), ("avgvar", _FP),
                ("means", _FP), ("R", _FP), ("Rinv", _FP), ("memberships", _FP)]


class Clusters:
    """Host-side cluster parameters for up to ``K`` clusters in ``D`` dims.

    ``n_events`` > 0 also allocates the ``[K][n_events]`` memberships array
    (reference: gaussian.cu:243-275).
    """
    FIELDS = ("N", "pi", "constant", "avgvar", "means", "R", "Rinv")

    def __init__(self, K, D, n_events=0):
        self.K, self.D, self.n_events = int(K), int(D), int(n_events)
        self.N = np.zeros(K, np.float32)
        self.pi = np.zeros(K, np.float32)
        self.constant = np.zeros(K, np.float32)
        self.avgvar = np.zeros(K, np.float32)
        self.means = np.zeros((K, D), np.float32)
        self.R = np.zeros((K, D, D), np.float32)
        self.Rinv = np.zeros((K, D, D), np.float32)
        self.memberships = np.zeros((K, n_events), np.float32) if n_events else None

    def struct(self):
        s = clusters_t()
        for f in self.FIELDS:
            setattr(s, f, getattr(self, f).ctypes.data_as(_FP))
        s.memberships = self.memberships.ctypes.data_as(_FP) if self.memberships is not None else _FP()
        return s

    def copy(self):
        o = Clusters(self.K, self.D, self.n_events)
        for f in self.FIELDS:
            getattr(o, f)[...] = getattr(self, f)
        if self.memberships is not None:
            o.memberships = self.memberships.copy()
            o.n_events = self.memberships.shape[1]
        return o

    def view(self, K):
        """Arrays restricted to the first K clusters (dict of views)."""
        d = {f: getattr(self, f)[:K] for f in self.FIELDS}
        if self.memberships is not None:
            d["memberships"] = self.memberships[:K]
        return d
