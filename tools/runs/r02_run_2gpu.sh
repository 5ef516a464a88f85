# two-GPU leg (gpurun --gpus 2): sharded walk() parity with the single-process walk (frames over sdw_nccl_gather_frames),
# then the 2-rank bench line (weights over sdw_nccl_broadcast_weights)
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_walk_multi_gpu.py -x -q --timeout 450 > gpurun_out/r02_walk_2gpu_test.log 2>&1; echo "rc=$?" >> gpurun_out/r02_walk_2gpu_test.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu_F30.json 2> gpurun_out/r02_bench_2gpu.err
tail -n 6 gpurun_out/r02_walk_2gpu_test.log; cat gpurun_out/r02_bench_2gpu_F30.json; tail -n 5 gpurun_out/r02_bench_2gpu.err
