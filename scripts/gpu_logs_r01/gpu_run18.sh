timeout 900 python -m pytest tests -x -q -m gpu -k "cut" 2>&1 | tail -3
for c in host device; do timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --cut $c 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$c', 'value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['breakdown_ms_per_step'])"; done
