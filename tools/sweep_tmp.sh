#!/bin/bash
# scratch sweep: staged vs flow, lanes
for cfg in "0 2" "0 3" "0 4" "0 8" "1 2"; do
  set -- $cfg
  echo "flow=$1 lanes=$2"
  AVB_FLOW=$1 python bench.py --lanes $2 --steps 10 --no-extras --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
