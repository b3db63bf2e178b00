"""Experiment: run-level deviation of the tensor M-step for different TMEM flush cadences."""
import sys, os, json, time
import numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as e
pkg = e.load_package(); o64 = e.load_oracle("f64")
out = {}
for (N, D, K, iters) in [(10000, 4, 8, 100), (100000, 16, 32, 10)]:
    ev = pkg.synth.make_blobs(N, D, K)
    ref = pkg.Clusters(K, D, N); o64.seed(ev, K, ref); o64.em(o64.transpose(ev), ref, K, iters, iters)
    for path in (pkg.PATH_SIMT, pkg.PATH_AUTO):
        with pkg.Engine(ev, K) as eng:
            eng.set_option("path", path); eng.seed(K); eng.em(K, iters, iters)
            got = eng.get_clusters(K, with_memberships=True)
        dg = np.abs(got.memberships - ref.memberships).max()
        dN = (np.abs(got.N - ref.N) / np.maximum(ref.N, 1)).max()
        dR = max(np.abs(got.R[k] - ref.R[k]).max() / np.abs(ref.R[k]).max() for k in range(K))
        dm = np.abs(got.means - ref.means).max()
        out[f"N{N}_D{D}_K{K}_it{iters}_path{path}"] = dict(dgamma=float(dg), dN=float(dN), dR=float(dR), dmeans=float(dm))
print(json.dumps(dict(lib=os.environ.get("GMM_B200_LIB", "default"), res=out)))
