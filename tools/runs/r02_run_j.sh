# round-2 run J (one B200): lane-group LayerNorm — tests, microbench and end-to-end A/B; smoke(); final default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/r02j_gpu_tests.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/r02j_gpu_tests.log
tail -n 4 gpurun_out/r02j_gpu_tests.log
if [ $rc -ne 0 ]; then echo "GPU tests failed: stopping"; exit 1; fi
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02j_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02j_smoke.log; tail -n 2 gpurun_out/r02j_smoke.log
(for v in 0 1; do echo "== SDW_LN_C40=$v"; SDW_LN_C40=$v ONLY=ln timeout 100 python tools/norm_bench.py; done
 M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum
 for v in 0 1; do echo "== ncu SDW_LN_C40=$v"; SDW_LN_C40=$v ONLY=ln ITERS=1 timeout 120 ncu --metrics $M --clock-control none -k regex:layernorm -s 3 -c 1 python tools/norm_bench.py 2>&1 | grep -E "layernorm_|dram__|gpu__time|sm__warps|smsp__"; done
) > gpurun_out/r02j_ln_ab.txt 2>&1
cat gpurun_out/r02j_ln_ab.txt
for v in 0 1 0 1; do SDW_LN_C40=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench_c40_$v.json 2> gpurun_out/r02j_bench.err; python - <<PY
import json; d=json.load(open("gpurun_out/r02j_bench_c40_$v.json")); print("SDW_LN_C40=$v", d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
PY
done > gpurun_out/r02j_bench_ab.txt 2>&1
cat gpurun_out/r02j_bench_ab.txt
timeout 500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02j_bench_F30.json 2> gpurun_out/r02j_bench_F30.err
cat gpurun_out/r02j_bench_F30.json | cut -c1-400
