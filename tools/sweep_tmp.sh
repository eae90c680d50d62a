timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for fl in 1 0; do AVB_FLOW=$fl timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), r['kernel_ms_per_step'], r.get('flow_task_share'))
"; done
