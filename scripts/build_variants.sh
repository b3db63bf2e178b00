# builds the experiment variants of libgmm_b200.so used by scripts/gpu_r2*.sh (cuda-gmm-mpi_b200/variants/, git-ignored)
cd "$(dirname "$0")/../cuda-gmm-mpi_b200/csrc" || exit 1
make -s -j8 || exit 1
for v in "$@"; do
  n=${v%%:*}; d=${v#*:}
  (make -s variant NAME=$n DEFS="$d" 2>&1 | grep -i "error\|warning") &
done
wait
ls ../variants/
