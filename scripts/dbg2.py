import sys, os
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import __graft_entry__ as e
pkg = e.load_package()
N, D, K = 20_001, 16, 32
ev = pkg.synth.make_blobs(N, D, 16, seed=316)
with pkg.Engine(ev, K) as eng:
    eng.set_option("path", pkg.PATH_TENSOR)
    eng.seed(K)
    eng.estep(K)
    eng.mstep(K)
    eng.constants(K)
    eng.estep(K)
    eng.mstep(K)
    got = eng.get_clusters(K)
    print("N sum", float(got.N.sum()), flush=True)
