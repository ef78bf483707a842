#!/usr/bin/env bash
# pulse-onset scan fast path: parity (incl. stage tests), then A/B against the sequential-only build on c2 and c3
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -x -s > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "worst|passed|failed|Error|error" gpurun_out/r2r_pytest.log | tail -14
bash tools/bench_variants.sh
bash tools/bench_variants.sh --workload c3 --notes 256
