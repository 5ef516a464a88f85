# round-2 check + evidence run (one B200): GPU tests, default bench, ncu captures, launch list, compute-sanitizer
mkdir -p gpurun_out; export SHAPE=60,8,4096,4096,40
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gpu_tests.log
timeout 500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_F30.json 2> gpurun_out/r02_bench_F30.err
F=30 timeout 100 python tools/attn_bench.py > gpurun_out/r02_attn_bench.txt 2>&1
F=30 timeout 200 python tools/epi_bench.py > gpurun_out/r02_epi_bench.txt 2>&1
# (1) the dominant kernel as shipped: ncu --set full (first launch after 3 warm-ups)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 3 -c 1 -f -o gpurun_out/r02_attn_self_d40 \
    python tools/attn_one.py > gpurun_out/r02_ncu_attn.log 2>&1
# (2) launch list of the bench command: a 1000-launch window inside the warm-up step (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 1000 --csv \
    --log-file gpurun_out/r02_ncu_launch_list_F30.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph \
    > gpurun_out/r02_ncu_launch_list.log 2>&1
# (3) compute-sanitizer
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_attn_gpu.py tests/test_norm_gpu.py \
    "tests/test_engine_gpu.py::test_full_sampler_tiny" -x -q -k "not subprocess" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_norm_gpu.py \
    "tests/test_engine_gpu.py::test_full_sampler_tiny[pndm-4]" -x -q > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.log
tail -n 5 gpurun_out/r02_gpu_tests.log; cat gpurun_out/r02_bench_F30.json; tail -n 3 gpurun_out/r02_bench_F30.err
cat gpurun_out/r02_attn_bench.txt | head -20
tail -n 3 gpurun_out/r02_ncu_attn.log gpurun_out/r02_sanitizer_memcheck.log gpurun_out/r02_sanitizer_racecheck.log
ls -la gpurun_out | tail -n 14
