"""Error budget of the tensor path (round 2): per-call and run-level deviation from the f64 oracle plus kernel time,
for library variants (GMM_B200_LIB) and for mixed SIMT/tensor step assignments.

  python scripts/exp_acc.py ref                      # oracle references -> gpurun_out/exp_acc_ref.npz (once)
  python scripts/exp_acc.py run <tag> [time]         # one library (GMM_B200_LIB or default), prints one JSON line
  python scripts/exp_acc.py all <lib|default> ...    # ref (if missing) + run for every library, one line each

Test infrastructure: uses the oracle; nothing here is product code."""
import json
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.getcwd())
import __graft_entry__ as e  # noqa: E402

REF = "gpurun_out/exp_acc_ref.npz"
RUNS = [("c1", 10000, 4, 8, 100, 8, None), ("c2s", 100000, 16, 32, 10, 32, None), ("c5s", 40000, 24, 128, 5, 32, 55)]
CALLS = [("k64", 200000, 24, 64), ("k32", 150001, 16, 32)]
FIELDS = ("N", "pi", "constant", "avgvar", "means", "R", "Rinv")


def blobs(pkg, N, D, nb, sd):
    return pkg.synth.make_blobs(N, D, nb, seed=sd) if sd else pkg.synth.make_blobs(N, D, nb)


def pack(cl, prefix, out, memb=True):
    for f in FIELDS:
        out[prefix + f] = getattr(cl, f).copy()
    if memb and cl.memberships is not None:
        out[prefix + "memb"] = cl.memberships.copy()


def unpack(pkg, z, prefix, K, D, N=0):
    cl = pkg.Clusters(K, D, N)
    for f in FIELDS:
        getattr(cl, f)[...] = z[prefix + f]
    if N:
        cl.memberships[...] = z[prefix + "memb"]
    return cl


def make_ref():
    pkg = e.load_package(); o64 = e.load_oracle("f64")
    out = {}
    for (name, N, D, K, iters, nb, sd) in RUNS:
        ev = blobs(pkg, N, D, nb, sd)
        ref = pkg.Clusters(K, D, N); o64.seed(ev, K, ref); o64.em(o64.transpose(ev), ref, K, iters, iters)
        pack(ref, f"run_{name}_", out)
    for (name, N, D, K) in CALLS:
        ev = pkg.synth.make_blobs(N, D, min(K, 16), seed=300 + D)
        ref = pkg.Clusters(K, D, N); o64.seed(ev, K, ref); soa = o64.transpose(ev); o64.em(soa, ref, K, 2, 2)
        pack(ref, f"call_{name}_start_", out, memb=False)
        o64.estep(soa, ref, K); o64.mstep(soa, ref, K); o64.constants(ref, K)
        pack(ref, f"call_{name}_", out, memb=False)
    os.makedirs("gpurun_out", exist_ok=True)
    np.savez(REF, **out)


def dev(got, ref, K):
    d = {}
    if got.memberships is not None and ref.memberships is not None:
        d["dgamma"] = float(np.abs(got.memberships - ref.memberships).max())
    d["dN"] = float((np.abs(got.N[:K] - ref.N[:K]) / np.maximum(ref.N[:K], 1)).max())
    d["dR"] = float(max(np.abs(got.R[k] - ref.R[k]).max() / np.abs(ref.R[k]).max() for k in range(K)))
    d["dmeans"] = float(np.abs(got.means[:K] - ref.means[:K]).max())
    return d


def run(tag, with_time):
    pkg = e.load_package()
    z = np.load(REF)
    out = {"tag": tag}
    mixes = {"tt": (-1, -1), "st": (pkg.PATH_SIMT, -1), "ts": (-1, pkg.PATH_SIMT)}      # (E-step, M-step)
    for (name, N, D, K, iters, nb, sd) in RUNS:
        ev = blobs(pkg, N, D, nb, sd)
        ref = unpack(pkg, z, f"run_{name}_", K, D, N)
        for mix, (pe, pm) in mixes.items():
            if D < 8 and mix != "tt":
                continue                                  # no tensor E-step below D = 8: the mixes coincide
            with pkg.Engine(ev, K) as eng:
                eng.set_option("path", pkg.PATH_AUTO); eng.set_option("estep_path", pe); eng.set_option("mstep_path", pm)
                eng.seed(K); eng.em(K, iters, iters)
                got = eng.get_clusters(K, with_memberships=True)
            out[f"run_{name}_{mix}"] = dev(got, ref, K)
    for (name, N, D, K) in CALLS:
        ev = pkg.synth.make_blobs(N, D, min(K, 16), seed=300 + D)
        start = unpack(pkg, z, f"call_{name}_start_", K, D)
        ref = unpack(pkg, z, f"call_{name}_", K, D)
        with pkg.Engine(ev, K) as eng:
            eng.set_option("path", pkg.PATH_TENSOR); eng.set_option("estep_path", pkg.PATH_SIMT)
            eng.seed(K); eng.set_clusters(K, start); eng.estep(K); eng.mstep(K); eng.constants(K)
            got = eng.get_clusters(K)
        out[f"call_{name}"] = dev(got, ref, K)
    if with_time:
        N, D, K = 4_000_000, 24, 64
        ev = pkg.synth.make_blobs(N, D, K)
        with pkg.Engine(ev, K) as eng:
            eng.seed(K); eng.estep(K)
            eng.em_iterations(K, 3); eng.profile(reset=True)
            eng.em_iterations(K, 10)
            p = eng.profile(reset=True)
            out["time_4M"] = {k: round(float(v) / 10, 4) for k, v in p.items() if k in ("estep_ms", "mstep_ms")}
    print(json.dumps(out))


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "ref":
        make_ref()
    elif mode == "run":
        run(sys.argv[2], len(sys.argv) > 3)
    else:
        if not os.path.exists(REF):
            make_ref()
        for lib in sys.argv[2:]:
            env = dict(os.environ)
            if lib != "default":
                env["GMM_B200_LIB"] = os.path.abspath(lib)
            r = subprocess.run([sys.executable, __file__, "run", os.path.basename(lib), "time"], env=env, capture_output=True, text=True, timeout=900)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("FAILED " + lib + " " + r.stderr[-800:]), flush=True)
