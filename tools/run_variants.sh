#!/bin/bash
# bench.py under a list of tunable settings (development aid): tools/run_variants.sh "--tune a=1" "--tune b=2" ...
export B200F_ALLOW_PROBES=1
for v in "$@"; do
  echo "== $v"; timeout 200 python bench.py --no-gallery --no-cpu-baseline --steps 200 --warmup 20 $v 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('kernel_ms'))"
done
