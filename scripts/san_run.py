"""Small tensor-path run for compute-sanitizer (memcheck / synccheck / racecheck): seed + 2 EM iterations."""
import os, sys
sys.path.insert(0, os.getcwd())
import __graft_entry__ as e
pkg = e.load_package()
N, D, K = int(os.environ.get("SAN_N", "6000")), int(os.environ.get("SAN_D", "24")), int(os.environ.get("SAN_K", "64"))
ev = pkg.synth.make_blobs(N, D, min(K, 8), seed=5)
with pkg.Engine(ev, K) as eng:
    eng.set_option("path", pkg.PATH_TENSOR)
    eng.seed(K)
    ll, it = eng.em(K, 2, 2)
    print("loglik", ll, "iters", it)
