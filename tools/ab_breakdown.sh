# per-launch-shape A/B in ONE gpurun call: bash tools/ab_breakdown.sh NAME "ENV=.." ...  -> gpurun_out/bd_<i>.json
i=0
for v in "FH_NOP=1" "$@"; do
  env $v timeout 300 python bench.py --steps 3 --warmup 3 --no-latency --no-cpu-baseline --no-accurate --breakdown gpurun_out/bd_$i.json > gpurun_out/bd_$i.out 2> gpurun_out/bd_$i.err
  python -c "
import json; d=json.loads(open('gpurun_out/bd_$i.out').read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],1), 'clock', d['clocks']['sm_mhz'], {k: v for k, v in list(d['stage_ms'].items())[:3]})" || tail -3 gpurun_out/bd_$i.err
  i=$((i+1))
done
