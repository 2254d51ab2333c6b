mkdir -p gpurun_out
export P3D_SYNC_TIMEOUT_S=30
nvidia-smi topo -m > gpurun_out/r2_dp2_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_dp.py -x -q > gpurun_out/r2_dp2_tests.log 2>&1; tail -30 gpurun_out/r2_dp2_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/bench_train_dp.py 64 4096 8192 > gpurun_out/r2_dp2_bench.txt 2>&1; tail -4 gpurun_out/r2_dp2_bench.txt
P3D_TRAIN_FUSED=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tools/bench_train_dp.py 64 4096 8192 > gpurun_out/r2_dp2_bench_unfused.txt 2>&1; tail -4 gpurun_out/r2_dp2_bench_unfused.txt
