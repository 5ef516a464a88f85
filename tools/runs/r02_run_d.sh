# round-2 run D (one B200): 16-warp epilogue (EW = 4) tests + A/B; attention softmax-loop A/B grid (cycles, same box)
mkdir -p gpurun_out; export SHAPE=60,8,4096,4096,40 F=30
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r02d_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02d_gpu_tests.log
M=sm__cycles_elapsed.max,gpu__time_duration.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
(for poly in 0 4; do for mode in 0 1 2 3; do echo "== attention poly=$poly mode=$mode (bit0 fused max, bit1 late PV wait)"
  SDW_ATTN_POLY=$poly SDW_ATTN_MODE=$mode ONLY_SELF=1 timeout 100 python tools/attn_bench.py 2>&1 | head -n 1
  SDW_ATTN_POLY=$poly SDW_ATTN_MODE=$mode timeout 200 ncu --metrics $M --clock-control none -k regex:attn_ -s 3 -c 1 python tools/attn_one.py 2>&1 | grep -E "cycles_elapsed|time_duration|pipe_xu"; done; done
 echo "== cross attention through the two-tile kernel: off / on"
 timeout 100 python tools/attn_bench.py 2>&1 | grep cross; SDW_ATTN_PP_CROSS=1 timeout 100 python tools/attn_bench.py 2>&1 | grep cross
) > gpurun_out/r02d_attn_ab.txt 2>&1
(ET=2 timeout 300 python tools/epi_bench.py
 for ew in 2 4; do for i in 0 1 2; do echo "== epi $i ew=$ew cycles"; ONLY=$i ITERS=2 ET=2 EW=$ew timeout 200 ncu --metrics $M,smsp__inst_executed.sum --clock-control none -k regex:gemm2 -s 3 -c 1 python tools/epi_bench.py 2>&1 | grep -E "cycles_elapsed|time_duration|inst_executed.sum"; done; done
) > gpurun_out/r02d_epi_ab.txt 2>&1
for v in 2 0 2 0; do SDW_GEMM_EW=$v timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_ew$v.json 2> gpurun_out/r02d_bench.err; python - <<PY
import json; d=json.load(open("gpurun_out/r02d_bench_ew$v.json")); print("SDW_GEMM_EW=$v", d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
PY
done > gpurun_out/r02d_bench_ab.txt 2>&1
tail -n 4 gpurun_out/r02d_gpu_tests.log; cat gpurun_out/r02d_attn_ab.txt gpurun_out/r02d_epi_ab.txt gpurun_out/r02d_bench_ab.txt
