#!/usr/bin/env bash
# BASELINE.json configs[4]: the whole-voicebank sweep, 65,536 notes in total, sharded by note over N GPUs of this box
# (N = 1: all of them on one GPU, in waves of 2,048).  Also the c2 line at the same N.  Usage: bash tools/bench_c5.sh N
N=${1:-1}
mkdir -p gpurun_out
PER=$((65536 / N))
if [ "$N" = "1" ]; then
  RUN="python bench.py"
else
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N"
fi
$RUN --workload c5 --notes $PER --noise device --steps 3 --warmup 3 --cpu-sample 0 --verify 4 > gpurun_out/r2_bench_c5_65536notes_${N}gpu.json 2> gpurun_out/r2_bench_c5_${N}gpu.err
echo "c5 N=$N rc=$?"
$RUN --steps 20 --warmup 5 --cpu-sample 0 > gpurun_out/r2_bench_c2_${N}gpu.json 2> gpurun_out/r2_bench_c2_${N}gpu.err
echo "c2 N=$N rc=$?"
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
for w in ("c5_65536notes", "c2"):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_{w}_{n}gpu.json").read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print(f"{w} N={n}: value {d['value']:.0f} notes/s ({d['ms_per_step']:.2f} ms/step), e2e {e.get('value', 0):.0f} notes/s ({e.get('ms_per_step', 0):.2f} ms/step) by rank {e.get('ms_per_step_by_rank')}, verify {d.get('verify', {}).get('ok')} {d.get('verify', {}).get('worst_max_abs')}")
    except Exception as ex:
        print(w, "FAILED", ex)
PY
grep -c "NCCL INFO" gpurun_out/r2_bench_c2_${N}gpu.err 2>/dev/null
