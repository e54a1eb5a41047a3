# per-shape breakdowns of several env variants in ONE call: bash tools/ab2.sh name1 "ENV=.." name2 "ENV=.." ...
while [ $# -gt 0 ]; do
  n=$1; e=$2; shift 2
  env $e timeout 300 python bench.py --steps 3 --warmup 3 --no-latency --no-cpu-baseline --no-accurate --breakdown gpurun_out/r2_bd_$n.json 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n', round(d['ms_per_step'],1), d['clocks']['sm_mhz'], d['stage_ms']['tc_conv'])"
done
