# usage: bash tools/run_gpu.sh <tag> [ENV=val ...]   -> pytest -m gpu, then a short bench with a per-kernel breakdown
tag=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency --breakdown gpurun_out/bd_$tag.json > gpurun_out/bench_$tag.log 2>&1
tail -c 4000 gpurun_out/bench_$tag.log | grep -o '"ms_per_step": [0-9.]*'
