#!/bin/bash
# peer-memory (fused) vs NCCL collectives on N GPUs of one box: bench.py value + phase times
N=${1:-2}
for c in peer nccl; do
  VICAN_B200_COLLECTIVE=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29700+N)) bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline --no-e2e \
      > gpurun_out/bench_${N}gpu_$c.json 2> gpurun_out/bench_${N}gpu_$c.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${N}gpu_$c.json").read().strip().splitlines()[-1])
    print("N=$N $c", "iter/s", round(d["value"], 1), {k: round(v, 2) for k, v in d["phase_ms"].items()},
          {k[:18]: round(v["ms"], 3) for k, v in d["roofline"]["kernels"].items()})
except Exception as ex:
    print("N=$N $c failed", ex)
PY
done
