#!/bin/bash
# usage: tools/sweep.sh "ENV1=.. ENV2=.." "..." : bench ms/step for each environment setting
for cfg in "$@"; do
  r=$(env $cfg timeout 200 python bench.py --steps 300 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['roofline']['achieved'],1), round(d['value']/1e6,1))" 2>&1)
  echo "$cfg => $r"
done
