# latency A/B in one gpurun call: default vs each env variant
run() { env $1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-accurate --latency-samples 100 > gpurun_out/abl_tmp.json 2> gpurun_out/abl_tmp.err; python -c "
import json; d=json.loads(open('gpurun_out/abl_tmp.json').read().strip().splitlines()[-1]); l=d['latency']; print('$1', 'clip10s p50 %.3f p99 %.3f | chunk1s p50 %.3f p99 %.3f' % (l['clip_10s_midpoint_p50_ms'], l['clip_10s_midpoint_p99_ms'], l['chunk_1s_euler4_p50_ms'], l['chunk_1s_euler4_p99_ms']))" || tail -3 gpurun_out/abl_tmp.err; }
run "FH_NOP=1"
for v in "$@"; do run "$v"; done
run "FH_NOP=1"
