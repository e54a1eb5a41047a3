mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -s -k "snake_chunked or test_vocoder or test_generate" 2>&1 | grep -E "kind=[23] \(2, 16, 700\)|kind=[23] \(3, 24|fp16.*SNR|passed|failed" | cut -c1-190
FH_SNAKE_NB=16 timeout 300 python -m pytest tests -m gpu -q -x -k "snake_chunked or test_vocoder" 2>&1 | tail -2
for nb in 16 8; do
  FH_SNAKE_NB=$nb timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency --breakdown gpurun_out/bd_nb$nb.json > gpurun_out/bench_nb$nb.log 2>&1
  echo "NB $nb: $(tail -c 4000 gpurun_out/bench_nb$nb.log | grep -o '"ms_per_step": [0-9.]*') $(grep -o '"fh_snake_aa_chunked": [0-9.]*' gpurun_out/bench_nb$nb.log | tail -1) $(grep -o '"tc_conv": [0-9.]*' gpurun_out/bench_nb$nb.log | tail -1)"
done
