#!/bin/bash
# compute-sanitizer evidence for profiles/ (run on a GPU box: gpurun -- bash tools/run_sanitizers.sh [ngpus])
# racecheck: shared-memory hazards of the sort / pair / histogram kernels; memcheck: every access,
# including the peer-store exchange kernel's stores into the other ranks' buffers (2 GPUs).
set -u
out=gpurun_out
mkdir -p $out
S=/usr/local/cuda/bin/compute-sanitizer
for tool in racecheck memcheck; do
  timeout 900 $S --tool $tool --print-limit 20 --error-exitcode 3 python tools/sanitize_driver.py \
    > $out/${TAG:-r3}_sanitizer_${tool}_1gpu.txt 2>&1
  echo "[$tool 1 GPU] exit $?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${TAG:-r3}_sanitizer_${tool}_1gpu.txt | tail -1)"
done
if [ "${1:-1}" -ge 2 ]; then
  timeout 900 $S --tool memcheck --target-processes all --print-limit 20 --error-exitcode 3 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    tools/sanitize_driver.py > $out/${TAG:-r3}_sanitizer_memcheck_2gpu.txt 2>&1
  echo "[memcheck 2 GPUs] exit $?: $(grep -E 'ERROR SUMMARY' $out/${TAG:-r3}_sanitizer_memcheck_2gpu.txt | tail -3 | tr '\n' ' ')"
fi
