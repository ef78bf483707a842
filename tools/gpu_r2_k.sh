#!/usr/bin/env bash
# Round 2, 8-GPU call K: end-to-end scaling against the number of sub-batches of the host entry point.
N=${1:-8}
mkdir -p gpurun_out
for S in 2 3 4 1; do
  GOOFER_HOST_SUBBATCHES=$S python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 --cpu-sample 0 --verify 0 --e2e-variants prod > gpurun_out/r2k_c2_${N}gpu_S${S}.json 2> gpurun_out/r2k_c2_${N}gpu_S${S}.err
  python - "$S" "$N" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r2k_c2_{sys.argv[2]}gpu_S{sys.argv[1]}.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("S =", sys.argv[1], "value", round(d["value"]), "e2e", round(e["value"]), "notes/s", round(e["ms_per_step"], 3), "ms by rank", e["ms_per_step_by_rank"], e["host_numa_binding_rank0"].get("how"))
PY
done
