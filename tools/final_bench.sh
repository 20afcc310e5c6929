#!/bin/bash
# Round-end bench lines on one B200 (run under gpurun): configs[1] (headline), configs[2] at N = 1, configs[3], configs[4], train mode.
out=gpurun_out; tag=${1:-r2g}
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python bench.py --global-batch 512 --no-cpu-baseline --no-gpu-baseline > $out/${tag}_bench_gb512.json 2> $out/${tag}_gb512.err
python bench.py --config wide --no-cpu-baseline --no-gpu-baseline > $out/${tag}_bench_wide.json 2> $out/${tag}_wide.err
python bench.py --config vae --no-cpu-baseline > $out/${tag}_bench_vae.json 2> $out/${tag}_vae.err
python bench.py --mode train --no-cpu-baseline --no-gpu-baseline > $out/${tag}_bench_train.json 2> $out/${tag}_train.err
for f in $out/${tag}_bench*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d.get("value"), d.get("e2e", {}).get("value"), d.get("ms_per_step"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
