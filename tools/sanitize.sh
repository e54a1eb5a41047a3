# compute-sanitizer over the GPU tests: memcheck on everything; racecheck (shared-memory hazards) on the kernels that
# synchronise with block barriers (the tcgen05 / bulk-copy kernels synchronise through mbarriers and the async proxy,
# which racecheck does not model, and their bounded waits trap under its slowdown)
mkdir -p gpurun_out
if [ "$1" != "race" ]; then
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests -m gpu -q -x \
  > gpurun_out/sanitizer_memcheck.log 2>&1
echo "exit $?" >> gpurun_out/sanitizer_memcheck.log
tail -6 gpurun_out/sanitizer_memcheck.log
fi
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 20 python -m pytest tests -m gpu -q \
  -k "snake_f32 or postprocess or logmel or resample or conv_f32 or conv_transpose_f32" \
  > gpurun_out/sanitizer_racecheck.log 2>&1
echo "exit $?" >> gpurun_out/sanitizer_racecheck.log
tail -12 gpurun_out/sanitizer_racecheck.log
