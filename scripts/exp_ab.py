"""A/B timing of library variants at config 3: E-step / M-step kernel ms per launch (gmm_get_profile)."""
import sys, os, json, subprocess
libs = sys.argv[1:]
code = r'''
import sys, os, json
sys.path.insert(0, os.getcwd())
import __graft_entry__ as e
pkg = e.load_package()
cfg = pkg.synth.CONFIGS["c3"]; N, D, K = int(os.environ.get("GMM_EXP_N", cfg["N"])), cfg["D"], cfg["K"]
ev = pkg.synth.make_blobs(N, D, K)
with pkg.Engine(ev, K) as eng:
    eng.seed(K); eng.estep(K)
    eng.em_iterations(K, 3); eng.profile(reset=True)
    eng.em_iterations(K, 10)
    p = eng.profile(reset=True)
print(json.dumps({k: v / 10 for k, v in p.items()}))
'''
for lib in libs:
    env = dict(os.environ)
    if lib != "default": env["GMM_B200_LIB"] = os.path.abspath(lib)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    lines = r.stdout.strip().splitlines()
    for ln in lines[:-1]:
        if "profile" in ln:
            print("   ", ln)
    print(lib, lines[-1] if lines else r.stderr[-500:])
