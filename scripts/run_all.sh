#!/bin/bash
# Reproduce the README numbers on a B200 box (1 GPU unless noted).  Every step is bounded by `timeout`.
set -u
mkdir -p gpurun_out
echo "== GPU parity tests";            timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -2
echo "== smoke";                       timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-160
echo "== bench (cfg4, 1 GPU)";         timeout 300 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_1gpu.json").read().strip().splitlines()[-1])
print("iter/s", round(d["value"], 1), "solve ms", round(d["ms_per_step"], 2), d["phase_ms"], "e2e", d["e2e"])
for k, v in d["roofline"]["kernels"].items():
    print(" ", k, "in-step %.3f ms (%.0f %% of measured peak), alone %.3f ms" % (v["ms"], 100 * v["frac"], v["isolated_ms"]))
PY
echo "== BASELINE configs through the dict API"; timeout 900 python scripts/bench_configs.py cfg1 cfg2 cfg3 2>&1 | tail -3 | cut -c1-400
echo "== launch list under ncu";       timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k regex:"vb::|cub::" -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv 2>/dev/null | head -14
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  echo "== multi-GPU: sharded == single, collectives, scaling"
  timeout 400 python -m pytest tests/test_dist.py -m gpu -q 2>&1 | tail -2
  scripts/scale_run.sh 2>&1 | grep "^N="
fi
