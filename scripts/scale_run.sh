#!/bin/bash
# strong-scaling sweep of bench.py on one box (the driver does the same at round end)
NGPU=$(nvidia-smi -L 2>/dev/null | wc -l)
for n in 8 4 2; do
  [ "$n" -le "$NGPU" ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_${n}gpu.json').read().strip().splitlines()[-1])
    print('N=$n', 'iter/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), d['phase_ms'], 'e2e', round(d['e2e']['value'], 1) if d['e2e'] else None,
          {k: round(v['ms'], 3) for k, v in d['roofline']['kernels'].items()})
except Exception as ex:
    print('N=$n failed', ex)
PY
  tail -2 gpurun_out/bench_${n}gpu.err | cut -c1-200
done
