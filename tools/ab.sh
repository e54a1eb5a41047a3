# A/B in ONE gpurun call (same box): usage  bash tools/ab.sh "ENV=1" ["ENV2=.."]  -> ms per step of the default and of each variant
run() { env $1 timeout 300 python bench.py --steps 4 --warmup 3 --no-latency --no-cpu-baseline --no-accurate > gpurun_out/ab_tmp.json 2> gpurun_out/ab_tmp.err; python -c "
import json; d=json.loads(open('gpurun_out/ab_tmp.json').read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],1), 'clock', d['clocks']['sm_mhz'], {k: v for k, v in list(d['stage_ms'].items())[:3]})" || tail -3 gpurun_out/ab_tmp.err; }
run "FH_NOP=1"
for v in "$@"; do run "$v"; done
run "FH_NOP=1"
