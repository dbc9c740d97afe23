p() { python -c "
import sys,json
l=[x for x in sys.stdin if x.startswith('{')][-1]; j=json.loads(l); print('$1', round(j['value']/1e6,3), j['clocks']['reasons'], j['clocks']['sm_mhz'], round(j['ksd']['ms'],3), round(j['ksd']['roofline']['frac'],3))"; }
python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | p "ksd first:"
python bench.py --steps 20 --warmup 5 --no-extras --ksd-last 2>/dev/null | p "ksd last: "
python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | p "ksd first:"
python bench.py --steps 20 --warmup 5 --no-extras --ksd-last 2>/dev/null | p "ksd last: "
