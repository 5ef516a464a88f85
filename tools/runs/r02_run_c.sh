# round-2 run C (one B200): full GPU tests after the division-free tile coordinates + attention changes; cycle-based A/B
mkdir -p gpurun_out; export SHAPE=60,8,4096,4096,40 F=30
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r02c_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_gpu_tests.log
M=sm__cycles_elapsed.max,gpu__time_duration.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
(for poly in 4 0 2; do echo "== attention poly=$poly"; SDW_ATTN_POLY=$poly ONLY_SELF=1 timeout 100 python tools/attn_bench.py 2>&1 | head -n 1
  SDW_ATTN_POLY=$poly timeout 200 ncu --metrics $M --clock-control none -k regex:attn_ -s 3 -c 1 python tools/attn_one.py 2>&1 | grep -E "cycles_elapsed|time_duration|pipe_xu|inst_executed.sum"; done
 echo "== epi_bench (division-free tile coordinates)"; ET=2 timeout 200 python tools/epi_bench.py 2>&1 | grep "et=2 as=1" | grep "auto"
 for i in 0 1 2; do echo "== epi $i cycles"; ONLY=$i ITERS=2 ET=2 timeout 200 ncu --metrics $M --clock-control none -k regex:gemm2 -s 3 -c 1 python tools/epi_bench.py 2>&1 | grep -E "cycles_elapsed|time_duration|inst_executed.sum"; done
) > gpurun_out/r02c_ab.txt 2>&1
for f in 30 24 20 30; do timeout 400 python bench.py --steps 3 --warmup 3 --frames-per-call $f --no-cpu-baseline > gpurun_out/r02c_bench_F$f.json 2> gpurun_out/r02c_bench_F$f.err; python - <<PY
import json; d=json.load(open("gpurun_out/r02c_bench_F$f.json")); print("F=$f", d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"], d["roofline"]["us_per_launch"])
PY
done > gpurun_out/r02c_bench_F.txt 2>&1
timeout 900 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 python -m pytest tests/test_attn_gpu.py tests/test_norm_gpu.py \
    "tests/test_engine_gpu.py::test_full_sampler_tiny" -x -q -k "not subprocess" > gpurun_out/r02c_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02c_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_norm_gpu.py \
    "tests/test_engine_gpu.py::test_full_sampler_tiny[pndm-4]" -x -q > gpurun_out/r02c_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02c_sanitizer_racecheck.log
tail -n 4 gpurun_out/r02c_gpu_tests.log; cat gpurun_out/r02c_ab.txt gpurun_out/r02c_bench_F.txt; tail -n 3 gpurun_out/r02c_sanitizer_memcheck.log gpurun_out/r02c_sanitizer_racecheck.log
