#!/usr/bin/env bash
# Round 2, GPU call J (1 GPU): sub-batched host entry point: parity with S = 2, timing with S = 1 / 2 / 4.
mkdir -p gpurun_out
GOOFER_HOST_SUBBATCHES=2 timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_parity.py -m gpu -q -x -k "device_drawn or host_entry or edge_lengths or noise_phases or pcm16" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest S=2 rc=$?"; tail -2 gpurun_out/r2j_pytest.log
for S in 1 2 4; do
  for P in default 0.125,0.25,0.375,0.5,0.625,0.75,0.875; do
    if [ "$P" = "default" ]; then unset GOOFER_HOST_PARTS; else export GOOFER_HOST_PARTS=$P; fi
    GOOFER_HOST_SUBBATCHES=$S python bench.py --steps 20 --warmup 5 --cpu-sample 0 --verify 0 --e2e-variants prod > gpurun_out/r2j_S${S}.json 2>/dev/null
    python - "$S" "$P" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2j_S{sys.argv[1]}.json"))
print("S =", sys.argv[1], "parts", sys.argv[2], "e2e", round(d["e2e"]["ms_per_step"], 3), d["e2e"]["variants"]["device_phases_pcm16"]["rank0_call_ms"])
PY
  done
done
