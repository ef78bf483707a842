#!/usr/bin/env bash
mkdir -p gpurun_out
bash tools/bench_variants.sh --no-e2e 2>&1 | sed 's/ e2e.*src_env/ src_env/'
