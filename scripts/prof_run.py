"""Minimal driver for profilers: seed + initial E-step + a few EM iterations at config-3 shape (N from GMM_EXP_N)."""
import os, sys
sys.path.insert(0, os.getcwd())
import __graft_entry__ as e
pkg = e.load_package()
cfg = pkg.synth.CONFIGS["c3"]; N, D, K = int(os.environ.get("GMM_EXP_N", cfg["N"])), cfg["D"], cfg["K"]
ev = pkg.synth.make_blobs(N, D, K)
with pkg.Engine(ev, K) as eng:
    eng.seed(K); eng.estep(K)
    eng.em_iterations(K, int(os.environ.get("GMM_EXP_ITERS", "3")))
    print(eng.profile())
