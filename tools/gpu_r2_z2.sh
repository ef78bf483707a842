#!/usr/bin/env bash
bash tools/bench_variants.sh --no-e2e 2>&1 | cut -c1-150
bash tools/bench_variants.sh --no-e2e --workload c3 --notes 256 2>&1 | cut -c1-150
bash tools/bench_variants.sh --no-e2e --workload c3 --notes 1024 2>&1 | cut -c1-150
