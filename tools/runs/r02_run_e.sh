# round-2 run E (one B200): 16-warp epilogue after the setmaxnreg budget fix — guarded: stop at the first failure / hang
mkdir -p gpurun_out; export SHAPE=60,8,4096,4096,40 F=30
timeout 150 python -m pytest tests/test_gemm_gpu.py -x -q -k epilogue_width --timeout 40 > gpurun_out/r02e_ew_tests.log 2>&1
rc=$?; echo "rc=$rc" >> gpurun_out/r02e_ew_tests.log; tail -n 5 gpurun_out/r02e_ew_tests.log
if [ $rc -ne 0 ]; then echo "EW tests failed: stopping"; exit 1; fi
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/r02e_gpu_tests.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/r02e_gpu_tests.log
tail -n 4 gpurun_out/r02e_gpu_tests.log
if [ $rc -ne 0 ]; then echo "GPU tests failed: stopping"; exit 1; fi
M=sm__cycles_elapsed.max,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
(ET=2 timeout 200 python tools/epi_bench.py
 for ew in 2 4; do for i in 0 1 2 3; do echo "== epi $i ew=$ew cycles"; ONLY=$i ITERS=2 ET=2 EW=$ew timeout 90 ncu --metrics $M --clock-control none -k regex:gemm2 -s 3 -c 1 python tools/epi_bench.py 2>&1 | grep -E "cycles_elapsed|time_duration|inst_executed.sum|issue_active"; done; done
 echo "== attention (shipped: FMA-pipe share 1/4, cross attention through the two-tile kernel)"; timeout 100 python tools/attn_bench.py
) > gpurun_out/r02e_epi_ab.txt 2>&1
cat gpurun_out/r02e_epi_ab.txt
for v in 2 0 2 0; do SDW_GEMM_EW=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_ew$v.json 2> gpurun_out/r02e_bench.err; python - <<PY
import json; d=json.load(open("gpurun_out/r02e_bench_ew$v.json")); print("SDW_GEMM_EW=$v", d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
PY
done > gpurun_out/r02e_bench_ab.txt 2>&1
cat gpurun_out/r02e_bench_ab.txt
