# round-2 A/B run (one B200): attention with the row max fused into the exponential pass (+ FMA-pipe share), short-K GEMM captures
mkdir -p gpurun_out; export SHAPE=60,8,4096,4096,40 F=30
timeout 600 python -m pytest tests/test_attn_gpu.py -q -x > gpurun_out/r02b_attn_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_attn_tests.log
SDW_ATTN_POLY=4 timeout 600 python -m pytest tests/test_attn_gpu.py -q -x > gpurun_out/r02b_attn_tests_poly4.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_attn_tests_poly4.log
(for rep in 1 2; do
  for poly in 0 4 2; do echo "fused-max poly=$poly"; SDW_ATTN_POLY=$poly ONLY_SELF=1 timeout 100 python tools/attn_bench.py 2>&1 | head -n 1; done
done
echo "all shapes poly=0"; timeout 100 python tools/attn_bench.py
echo "all shapes poly=4"; SDW_ATTN_POLY=4 timeout 100 python tools/attn_bench.py
echo "trace poly=0"; BKV=128 timeout 100 python tools/attn_trace.py
echo "trace poly=4"; SDW_ATTN_POLY=4 BKV=128 timeout 100 python tools/attn_trace.py) > gpurun_out/r02b_attn_matrix.txt 2>&1
# short-K GEMMs at batch 60: ncu --set full of the GEGLU projection and the attention out-projection
for i in 0 1; do
  ONLY=$i ITERS=2 ET=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 3 -c 1 -f \
      -o gpurun_out/r02_epi$i python tools/epi_bench.py > gpurun_out/r02_ncu_epi$i.log 2>&1
done
# frames per call: 20 / 24 vs the default 30 (activation working set vs the 126 MB L2)
for f in 20 24; do timeout 400 python bench.py --steps 3 --warmup 3 --frames-per-call $f --no-cpu-baseline > gpurun_out/r02b_bench_F$f.json 2> gpurun_out/r02b_bench_F$f.err; done
tail -n 3 gpurun_out/r02b_attn_tests.log gpurun_out/r02b_attn_tests_poly4.log; cat gpurun_out/r02b_attn_matrix.txt
for f in 20 24; do python - <<EOF
import json; d=json.load(open("gpurun_out/r02b_bench_F$f.json")); print($f, d["value"], d["e2e"]["value"], d["ms_per_step"])
EOF
done
ls -la gpurun_out | tail -n 8
