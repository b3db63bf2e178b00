import sys, os, json
import numpy as np
np.set_printoptions(linewidth=250, precision=1, suppress=False)
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import __graft_entry__ as e
pkg = e.load_package(); o64 = e.load_oracle("f64")
from conftest import fitted_params
print("lib", os.environ.get("GMM_B200_LIB", "default"), flush=True)
N, D, K = 70_001, 16, 32
ev = pkg.synth.make_blobs(N, D, 16, seed=300 + D)
ref = fitted_params(pkg, o64, ev, K, iters=1)
soa = o64.transpose(ev)
exact = ref.copy()
o64.estep(soa, exact, K); o64.mstep(soa, exact, K)
for rep in range(2):
  for gs in (1, 0):
    with pkg.Engine(ev, K) as eng:
        eng.set_option("path", pkg.PATH_TENSOR); eng.set_option("mstep_gamma_split", gs)
        eng.seed(K); eng.set_clusters(K, ref)
        eng.estep(K); g1 = eng.get_clusters(K, with_memberships=True).memberships.copy()
        eng.mstep(K)
        got = eng.get_clusters(K)
        dm = np.abs(got.means - exact.means).max(0)
        dR = np.abs(got.R - exact.R).reshape(K, D, D).max(0)
        print("rep", rep, "gs", gs, "memb err vs oracle", float(np.abs(g1 - exact.memberships).max()))
        print(" mean err per dim:", dm)
        print(" R err (max over clusters), lower triangle rows:")
        for i in range(D): print("  ", i, dR[i, :i + 1])
