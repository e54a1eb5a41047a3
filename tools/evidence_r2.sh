# Final round-2 evidence on one B200 box: tests, smoke, bench (both arms), sanitizer over the kernels changed last,
# ncu launch list, ncu --set full of the snake, DRAM traffic of one B = 64 step.
R=${1:-r2k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -120 > gpurun_out/${R}_pytest_gpu.log; tail -2 gpurun_out/${R}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 > gpurun_out/${R}_smoke.log; tail -2 gpurun_out/${R}_smoke.log
timeout 600 python bench.py --breakdown gpurun_out/${R}_breakdown_b64.json > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; cut -c1-300 gpurun_out/${R}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/${R}_bench_reference_arm.json
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests -m gpu -q -x \
  -k "residual_ring or tc_conv1d or tc_geglu or tc_linear or snake_chunked or vocoder_16bit or large_batch" > gpurun_out/${R}_sanitizer_memcheck.log 2>&1
echo "exit $?" >> gpurun_out/${R}_sanitizer_memcheck.log; tail -4 gpurun_out/${R}_sanitizer_memcheck.log
P="python tools/profile_step.py --precision fp16"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv $P --batch 8 > /dev/null 2>&1
F="ncu --profile-from-start off --set full --clock-control none --import-source on -f"
$F -k regex:snake_aa_mma -s 96 -c 2 -o gpurun_out/${R}_prof_snake $P --batch 64 > /dev/null 2>&1
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_traffic_b64.csv $P --batch 64 > /dev/null 2>&1
ls -la gpurun_out/${R}_*
