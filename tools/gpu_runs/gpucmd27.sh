p() { python -c "
import sys,json
l=[x for x in sys.stdin if x.startswith('{')][-1]; j=json.loads(l); print('$1', j['value'], j['ksd']['ms'], j['ksd']['roofline']['frac'])"; }
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | p "no cpu baseline, no extras:"
python bench.py --steps 5 --warmup 3 --no-extras --cpu-budget 4 2>/dev/null | p "with cpu baseline:"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 20 2>/dev/null | p "short e2e:"
