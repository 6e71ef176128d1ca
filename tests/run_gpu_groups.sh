#!/bin/bash
# Runs the GPU kernel tests group by group, each in its own process, so that one faulting
# kernel (sticky CUDA error) does not hide the results of the others.  Logs -> gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for grp in "layernorm or l2norm" "gemm_identity" "gemm_plain" "gemm_epilogues" "gemm_lora" "attention" "search"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$grp" -p no:cacheprovider \
      --timeout 300 > "gpurun_out/test_${name}.log" 2>&1
  echo "[$grp] exit $?  $(tail -n 1 gpurun_out/test_${name}.log)"
done
