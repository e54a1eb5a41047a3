# Last regression of the round on one B200 box: full GPU suite, smoke, bench record, memcheck over the whole GPU suite.
R=${1:-r2m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -130 > gpurun_out/${R}_pytest_gpu.log; tail -1 gpurun_out/${R}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 > gpurun_out/${R}_smoke.log; tail -1 gpurun_out/${R}_smoke.log
timeout 600 python bench.py --breakdown gpurun_out/${R}_breakdown_b64.json > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; cut -c1-200 gpurun_out/${R}_bench_n1.json
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests -m gpu -q -x \
  --deselect tests/test_gpu_kernels.py::test_big_config_batch64_matches_single > gpurun_out/${R}_sanitizer_memcheck.log 2>&1
echo "exit $?" >> gpurun_out/${R}_sanitizer_memcheck.log; tail -4 gpurun_out/${R}_sanitizer_memcheck.log
