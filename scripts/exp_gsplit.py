"""Experiment: tensor M-step with/without the FP16 hi/lo pair for the responsibilities —
per-call and run-level deviation against the f64 oracle, and kernel time at config 3 / 5 shapes."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as e
pkg = e.load_package(); o64 = e.load_oracle("f64")
out = {}
def dev(got, ref, K):
    dg = float(np.abs(got.memberships - ref.memberships).max()) if got.memberships is not None and ref.memberships is not None else None
    dN = float((np.abs(got.N - ref.N) / np.maximum(ref.N, 1)).max())
    dR = float(max(np.abs(got.R[k] - ref.R[k]).max() / np.abs(ref.R[k]).max() for k in range(K)))
    dm = float(np.abs(got.means - ref.means).max())
    return dict(dgamma=dg, dN=dN, dR=dR, dmeans=dm)
# run level
for (N, D, K, iters, nb, sd) in [(10000, 4, 8, 100, 8, None), (100000, 16, 32, 10, 32, None), (40000, 24, 128, 5, 32, 55)]:
    ev = pkg.synth.make_blobs(N, D, nb, seed=sd) if sd else pkg.synth.make_blobs(N, D, nb)
    ref = pkg.Clusters(K, D, N); o64.seed(ev, K, ref); o64.em(o64.transpose(ev), ref, K, iters, iters)
    for gs in (0, 1):
        with pkg.Engine(ev, K) as eng:
            eng.set_option("path", pkg.PATH_AUTO); eng.set_option("mstep_gamma_split", gs); eng.seed(K); eng.em(K, iters, iters)
            got = eng.get_clusters(K, with_memberships=True)
        out[f"run_N{N}_D{D}_K{K}_gs{gs}"] = dev(got, ref, K)
# per call (one M-step on oracle responsibilities), incl. tiny clusters (K much larger than blobs)
for (N, D, K) in [(200000, 24, 64), (20000, 24, 64), (3000, 16, 32)]:
    ev = pkg.synth.make_blobs(N, D, min(K, 16), seed=300 + D)
    ref = pkg.Clusters(K, D, N); o64.seed(ev, K, ref); soa = o64.transpose(ev); o64.em(soa, ref, K, 1, 1)
    start = ref.copy()
    o64.estep(soa, ref, K); o64.mstep(soa, ref, K); o64.constants(ref, K)
    for gs in (0, 1):
        with pkg.Engine(ev, K) as eng:
            eng.set_option("path", pkg.PATH_TENSOR); eng.set_option("mstep_gamma_split", gs); eng.seed(K)
            eng.set_clusters(K, start); eng.estep(K); eng.mstep(K); eng.constants(K)
            got = eng.get_clusters(K)
        d = dev(got, ref, K); d["minN"] = float(ref.N.min())
        out[f"call_N{N}_D{D}_K{K}_gs{gs}"] = d
# timing
import torch
for name in ("c3", "c5"):
    cfg = pkg.synth.CONFIGS[name]; N, D, K = cfg["N"], cfg["D"], cfg["K"]
    if name == "c5": N = 2_000_000
    ev = pkg.synth.make_blobs(N, D, K)
    with pkg.Engine(ev, K) as eng:
        eng.seed(K); eng.estep(K)
        for gs in (0, 1):
            eng.set_option("mstep_gamma_split", gs)
            eng.em_iterations(K, 3); eng.profile(reset=True)
            eng.em_iterations(K, 10)
            p = eng.profile(reset=True)
            out[f"time_{name}_N{N}_gs{gs}"] = {k: float(v) for k, v in p.items()}
print(json.dumps(out, indent=1))
