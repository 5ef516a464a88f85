# round-2 final evidence run (one B200): tests, default bench (with the CPU baseline leg), ncu captures of the shipped kernels
mkdir -p gpurun_out; export SHAPE=60,8,4096,4096,40
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r02f_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02f_gpu_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02f_bench_F30.json 2> gpurun_out/r02f_bench_F30.err
F=30 timeout 200 python tools/op_profile.py gpurun_out/r02f_op_profile.tsv > gpurun_out/r02f_op_profile_F30.txt 2>&1
F=30 timeout 100 python tools/attn_bench.py > gpurun_out/r02f_attn_bench.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_pp -s 3 -c 1 -f -o gpurun_out/r02f_attn_self_d40 \
    python tools/attn_one.py > gpurun_out/r02f_ncu_attn.log 2>&1
F=30 ONLY=0 ITERS=2 ET=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 3 -c 1 -f \
    -o gpurun_out/r02f_epi_geglu_ew4 python tools/epi_bench.py > gpurun_out/r02f_ncu_epi0.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 1000 --csv \
    --log-file gpurun_out/r02f_ncu_launch_list_F30.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph \
    > gpurun_out/r02f_ncu_launch_list.log 2>&1
tail -n 4 gpurun_out/r02f_gpu_tests.log; cat gpurun_out/r02f_bench_F30.json; tail -n 3 gpurun_out/r02f_bench_F30.err
cat gpurun_out/r02f_attn_bench.txt; head -n 30 gpurun_out/r02f_op_profile_F30.txt; tail -n 2 gpurun_out/r02f_ncu_attn.log gpurun_out/r02f_ncu_epi0.log
