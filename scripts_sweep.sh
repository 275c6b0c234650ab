#!/bin/bash
for c in 4 5 6; do
  echo "== VB_PASS_CTAS=$c"
  VB_PASS_CTAS=$c python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 1 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_c$c.json')); print('iter/s',round(d['value'],2),'ms/step',round(d['ms_per_step'],2)); [print(k, round(v['ms'],4), 'ms', round(v['frac'],3)) for k,v in d['roofline']['kernels'].items()]"
  tail -2 gpurun_out/bench_c$c.err
done
