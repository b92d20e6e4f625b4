#!/bin/bash
# compute-sanitizer over the small-size GPU parity tests that exercise every hand-written kernel family
# (tcgen05/TMA sweeps, factor updates, SC3 distances / Jacobi / tridiagonal eigensolver / k-means / linkage, transfer).
# usage: tools/sanitize.sh memcheck|racecheck|synccheck|initcheck   -> gpurun_out/sanitizer_<tool>.log
tool=${1:-memcheck}
mkdir -p gpurun_out
sel="f16_sweeps or tcgen05_sweeps or operand_shadow or single_update or block_jacobi or eigh_solver or tridiagonal or hclust or kmeans_vs or consensus_matrix or pipeline_labels or transfer_vs_oracle or euclidean_bit_exact_shapes or pearson"
timeout 1500 compute-sanitizer --tool "$tool" --error-exitcode 9 --log-file gpurun_out/sanitizer_${tool}.log \
  python -m pytest tests/test_gpu_nmf.py tests/test_gpu_sc3.py tests/test_gpu_transfer.py -q -m gpu -x -k "$sel" \
  > gpurun_out/sanitizer_${tool}.pytest.log 2>&1
rc=$?
echo "rc=$rc"
tail -3 gpurun_out/sanitizer_${tool}.pytest.log
tail -5 gpurun_out/sanitizer_${tool}.log
