#!/bin/bash
# Generates tests/golden/ref_c1.* by running the UNMODIFIED reference program
# (oracle/_ref/gaussianMPI_ref: /root/reference sources compiled for sm_100a by
# `make -C oracle ref`, with the cutil.h / single-rank mpi.h shims) on a B200.
# Run on the GPU box:   gpurun -- 'bash tests/golden/make_ref_golden.sh gpurun_out/golden'
# then copy gpurun_out/golden/* into tests/golden/ and commit.
set -euo pipefail
OUT=${1:-gpurun_out/golden}
ITERS=${GMM_REF_ITERS:-100}
mkdir -p "$OUT"
python - "$OUT" <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import __graft_entry__ as e
pkg = e.load_package()
cfg = pkg.synth.CONFIGS["c1"]
ev = pkg.synth.make_blobs(cfg["N"], cfg["D"], cfg["K"])
pkg.synth.write_bin(os.path.join(sys.argv[1], "c1.bin"), ev)
PY
export OMP_NUM_THREADS=1 CUDA_VISIBLE_DEVICES=0 GMM_REF_ITERS=$ITERS
./oracle/_ref/gaussianMPI_ref 8 "$OUT/c1.bin" "$OUT/ref_c1" 8 > "$OUT/ref_c1.stdout" 2>&1
head -n 512 "$OUT/ref_c1.results" > "$OUT/ref_c1.results.head"
rm -f "$OUT/ref_c1.results"
printf "N=10000\nD=4\nK=8\nseed=20260921\niters=%s\n" "$ITERS" > "$OUT/ref_c1.meta"
ls -la "$OUT"
