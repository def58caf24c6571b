#!/bin/bash
# usage: tools/sweep_env.sh "VAR1=a VAR2=b" "VAR1=c" ...   -- one short bench per environment, prints SpMV launch times
for envs in "$@"; do
  out=gpurun_out/sweep_$(echo "$envs" | tr ' =' '__').json
  env $envs timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 3 > "$out" 2>/dev/null
  python - "$out" "$envs" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[2], "| cells/s", round(d["value"]), "ms", round(d["ms_per_step"], 1), "| spmv us", round(r["avg_launch_us"], 1),
          "A", round(r["cells_out"]["avg_us"], 1), "B", round(r["genes_out"]["avg_us"], 1))
except Exception as e:
    print(sys.argv[2], "ERR", e)
PY
done
