mkdir -p gpurun_out; export F=30 SHAPE=60,8,4096,4096,40 ONLY_SELF=1
timeout 100 python tools/attn_smoke.py > gpurun_out/attn_smoke.txt 2>&1 || { cat gpurun_out/attn_smoke.txt; exit 1; }
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests_full.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_tests_full.log
(for rep in 1 2; do
  echo "pp      "; SDW_ATTN_DB=0 timeout 100 python tools/attn_bench.py 2>&1 | head -n 1
  echo "pp token"; SDW_ATTN_DB=0 SDW_ATTN_TOKEN=1 timeout 100 python tools/attn_bench.py 2>&1 | head -n 1
  echo "db      "; timeout 100 python tools/attn_bench.py 2>&1 | head -n 1
  echo "db poly4"; SDW_ATTN_POLY=4 timeout 100 python tools/attn_bench.py 2>&1 | head -n 1
  echo "db poly2"; SDW_ATTN_POLY=2 timeout 100 python tools/attn_bench.py 2>&1 | head -n 1
  echo "one-tile"; SDW_ATTN_PP=0 timeout 100 python tools/attn_bench.py 2>&1 | head -n 1
done
echo "trace pp"; SDW_ATTN_DB=0 BKV=128 timeout 100 python tools/attn_trace.py
echo "trace pp token"; SDW_ATTN_DB=0 SDW_ATTN_TOKEN=1 BKV=128 timeout 100 python tools/attn_trace.py
echo "trace db"; BKV=96 timeout 100 python tools/attn_trace.py) > gpurun_out/attn_matrix2.txt 2>&1
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_F30.json 2> gpurun_out/bench_F30.err
tail -n 6 gpurun_out/gpu_tests_full.log; cat gpurun_out/attn_matrix2.txt; cat gpurun_out/bench_F30.json; tail -n 3 gpurun_out/bench_F30.err
