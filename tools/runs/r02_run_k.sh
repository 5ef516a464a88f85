# round-2 run K (one B200): packed-arithmetic GroupNorm apply — tests, microbench, one end-to-end A/B pair
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/r02k_gpu_tests.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/r02k_gpu_tests.log
tail -n 4 gpurun_out/r02k_gpu_tests.log
(for v in 0 1; do echo "== SDW_GN_PACKED=$v"; SDW_GN_PACKED=$v ONLY=gn timeout 100 python tools/norm_bench.py; done) > gpurun_out/r02k_gn_ab.txt 2>&1
cat gpurun_out/r02k_gn_ab.txt
for v in 0 1; do SDW_GN_PACKED=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench_gnpk$v.json 2> gpurun_out/r02k_bench.err; python - <<PY
import json; d=json.load(open("gpurun_out/r02k_bench_gnpk$v.json")); print("SDW_GN_PACKED=$v", d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
PY
done > gpurun_out/r02k_bench_ab.txt 2>&1
cat gpurun_out/r02k_bench_ab.txt
