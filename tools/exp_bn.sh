mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency --breakdown gpurun_out/bd7_$tag.json > gpurun_out/x7_$tag.log 2>&1; echo $tag; tail -c 4000 gpurun_out/x7_$tag.log | grep -o '"ms_per_step": [0-9.]*'; }
run both FH_X=0
run nov8 FH_TC_V8=0
run nobulk FH_SNAKE_BULK=0
