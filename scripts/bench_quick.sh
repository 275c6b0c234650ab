#!/bin/bash
# quick GPU check: tests + bench summary (used during kernel iteration)
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -2
python bench.py --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_q.json'))
print('iter/s', round(d['value'], 2), 'ms/step', round(d['ms_per_step'], 2), d['phase_ms'], 'cg', d['cg_iters'])
print('e2e', d['e2e'])
for k, v in d['roofline']['kernels'].items():
    print(k, 'in-step', round(v['ms'], 4), 'ms', round(v['frac'], 3), ' isolated', round(v['isolated_ms'], 4), round(v['isolated_frac'], 3))
PY
tail -3 gpurun_out/bench_q.err
