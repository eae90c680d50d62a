#!/bin/bash
# GPU parity tests + one bench line with the per-phase split (development helper)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-extras --no-cpu-baseline "$@" > gpurun_out/bench_dev.log 2>&1
tail -1 gpurun_out/bench_dev.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3))
print(d['roofline']['kernel_ms_per_step'])
print(d['roofline'].get('flow_task_share'))
print(d['roofline'].get('flow_phase_share'))
" || tail -5 gpurun_out/bench_dev.log
