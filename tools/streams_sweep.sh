for cfg in "0 2" "0 3" "1 3" "0 4" "1 4" "0 2"; do set -- $cfg; WOTB_PDL_ALWAYS=$1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 39 --warmup 3 --streams $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pdl_always=$1 streams=$2', round(d['value'],3), round(d['e2e']['value'],3), round(d['roofline']['frac'],4))"; done
