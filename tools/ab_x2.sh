run() { env $1 timeout 300 python bench.py --precision fp16x2 --steps 3 --warmup 3 --no-latency --no-cpu-baseline --no-accurate > gpurun_out/abx_tmp.json 2> gpurun_out/abx_tmp.err; python -c "
import json; d=json.loads(open('gpurun_out/abx_tmp.json').read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],1), 'clock', d['clocks']['sm_mhz'], {k: v for k, v in list(d['stage_ms'].items())[:2]})" || tail -3 gpurun_out/abx_tmp.err; }
for v in "$@"; do run "$v"; done
