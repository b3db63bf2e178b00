"""Synthetic inputs of SURVEY.md §8(d): seeded Gaussian blobs in the reference's
``.bin`` format (readData.cpp:35-47: int32 N, int32 D, float32[N][D])."""
import numpy as np

SEED = 20260921

CONFIGS = {            # BASELINE.json configs
    "c1": dict(N=10_000, D=4, K=8),
    "c2": dict(N=1_000_000, D=16, K=32),
    "c3": dict(N=10_000_000, D=24, K=64),
    "c5": dict(N=10_000_000, D=24, K=128, K_true=16, target=16),
}


def make_blobs(N, D, K_true, seed=SEED, dtype=np.float32):
    """K_true blobs: centres ~ U(-10,10)^D, covariance A A^T / D + 0.5 I with
    A ~ N(0,1)^{DxD}, weights ~ Dirichlet(5), events shuffled."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(-10.0, 10.0, size=(K_true, D))
    w = rng.dirichlet(np.full(K_true, 5.0))
    counts = np.floor(w * N).astype(np.int64)
    counts[-1] += N - counts.sum()
    out = np.empty((N, D), dtype=dtype)
    pos = 0
    for j in range(K_true):
        A = rng.standard_normal((D, D))
        cov = A @ A.T / D + 0.5 * np.eye(D)
        L = np.linalg.cholesky(cov)
        n = int(counts[j])
        # chunked to bound peak memory at N = 1e7
        for s in range(0, n, 1 << 20):
            m = min(1 << 20, n - s)
            z = rng.standard_normal((m, D))
            out[pos + s:pos + s + m] = (z @ L.T + centres[j]).astype(dtype)
        pos += n
    perm = rng.permutation(N)
    return out[perm]


def write_bin(path, events):
    events = np.ascontiguousarray(events, dtype=np.float32)
    with open(path, "wb") as f:
        np.array([events.shape[0], events.shape[1]], dtype=np.int32).tofile(f)
        events.tofile(f)


def read_bin(path):
    with open(path, "rb") as f:
        n, d = np.fromfile(f, dtype=np.int32, count=2)
        return np.fromfile(f, dtype=np.float32, count=int(n) * int(d)).reshape(int(n), int(d))
