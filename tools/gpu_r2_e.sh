#!/usr/bin/env bash
# Round 2, multi-GPU call: bash tools/gpu_r2_e.sh N [tests]  -- c5 (65,536 notes over N GPUs) and c2 (1,024 notes per GPU) lines
# under torchrun, plus (optionally) the whole GPU test suite incl. the multi-device tests.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$2" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_pytest_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2e_pytest_${N}gpu.log
fi
NCCL_DEBUG=INFO bash tools/bench_c5.sh $N
