# round-2 run H (one B200): chunked unfused (VAE) attention + back-to-front norm passes: tests, then same-box A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/r02h_gpu_tests.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/r02h_gpu_tests.log
tail -n 4 gpurun_out/r02h_gpu_tests.log
if [ $rc -ne 0 ]; then echo "GPU tests failed: stopping"; exit 1; fi
for v in 0 1; do SDW_NORM_REV=$v F=30 timeout 200 python tools/op_profile.py gpurun_out/r02h_op_profile_rev$v.tsv 2>&1 | grep -E "unet:|vae:|groupnorm C320 60x64x64|layernorm C320|groupnorm C640 60x32x32|layernorm C640|attention d512|gemm conv0 C320 60x1x4096 N960|N2560 mode1" | sed "s/^/rev=$v /"; done > gpurun_out/r02h_norm_ab.txt 2>&1
cat gpurun_out/r02h_norm_ab.txt
for v in 0 1 0 1; do SDW_NORM_REV=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_rev$v.json 2> gpurun_out/r02h_bench.err; python - <<PY
import json; d=json.load(open("gpurun_out/r02h_bench_rev$v.json")); print("SDW_NORM_REV=$v", d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
PY
done > gpurun_out/r02h_bench_ab.txt 2>&1
cat gpurun_out/r02h_bench_ab.txt
