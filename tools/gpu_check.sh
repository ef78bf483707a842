#!/usr/bin/env bash
# What a round-end check runs on a B200 box (from the repo root, e.g. `gpurun -- 'bash tools/gpu_check.sh'`):
# GPU parity tests, smoke(), the parity table of the 16 pinned cases, and the default bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python tests/gpu_debug.py > gpurun_out/parity_table.txt 2>&1; tail -16 gpurun_out/parity_table.txt
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
cat gpurun_out/bench_default.json
