# Round evidence on one B200 box: tests, bench (both arms), ncu launch list, ncu --set full of the top kernels, DRAM traffic.
R=${1:-r1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${R}_pytest_gpu.log; tail -2 gpurun_out/${R}_pytest_gpu.log
timeout 900 python bench.py --breakdown gpurun_out/${R}_breakdown_b64.json > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; tail -c 600 gpurun_out/${R}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2>/dev/null; cat gpurun_out/${R}_bench_reference_arm.json | cut -c1-300
P="python tools/profile_step.py --precision fp16"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv $P --batch 8 > /dev/null 2>&1
F="ncu --profile-from-start off --set full --clock-control none --import-source on -f"
$F -k regex:tc_conv_kernel -s 34 -c 2 -o gpurun_out/${R}_prof_tc768 $P --batch 8 > /dev/null 2>&1
$F -k regex:tc_conv_kernel -s 47 -c 2 -o gpurun_out/${R}_prof_tc384 $P --batch 8 > /dev/null 2>&1
$F -k regex:tc_conv_kernel -s 98 -c 2 -o gpurun_out/${R}_prof_tc48 $P --batch 8 > /dev/null 2>&1
$F -k regex:snake_aa_mma -s 54 -c 2 -o gpurun_out/${R}_prof_snake $P --batch 8 > /dev/null 2>&1
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_traffic_b64.csv $P --batch 64 > /dev/null 2>&1
ls -la gpurun_out/${R}_*
