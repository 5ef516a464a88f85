# round-2 run I (one B200): streaming LayerNorm A/B (microbench, tests, bench), ncu of the two HBM-bound norm passes
mkdir -p gpurun_out
(for v in 0 2 4; do echo "== SDW_LN_STREAM=$v"; SDW_LN_STREAM=$v ONLY=ln timeout 100 python tools/norm_bench.py; done
 echo "== groupnorm"; ONLY=gn timeout 100 python tools/norm_bench.py) > gpurun_out/r02i_norm_bench.txt 2>&1
cat gpurun_out/r02i_norm_bench.txt
SDW_LN_STREAM=2 timeout 300 python -m pytest tests/test_norm_gpu.py tests/test_engine_gpu.py -q -x --timeout 200 > gpurun_out/r02i_ln_stream_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02i_ln_stream_tests.log
tail -n 3 gpurun_out/r02i_ln_stream_tests.log
for v in 0 2 0 2; do SDW_LN_STREAM=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_lns$v.json 2> gpurun_out/r02i_bench.err; python - <<PY
import json; d=json.load(open("gpurun_out/r02i_bench_lns$v.json")); print("SDW_LN_STREAM=$v", d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
PY
done > gpurun_out/r02i_bench_ab.txt 2>&1
cat gpurun_out/r02i_bench_ab.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct
(ONLY=gn ITERS=1 timeout 120 ncu --metrics $M --clock-control none -k regex:gn_ -s 6 -c 3 python tools/norm_bench.py
 ONLY=ln ITERS=1 timeout 120 ncu --metrics $M --clock-control none -k regex:layernorm -s 3 -c 1 python tools/norm_bench.py) 2>&1 | grep -vE "^==PROF|^$" > gpurun_out/r02i_ncu_norms.txt
cat gpurun_out/r02i_ncu_norms.txt | head -60
