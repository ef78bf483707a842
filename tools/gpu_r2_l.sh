#!/usr/bin/env bash
mkdir -p gpurun_out
bash tools/bench_variants.sh
bash tools/bench_c5.sh 1
