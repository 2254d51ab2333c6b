# Round 2, 8-GPU call: the bench line at N = 8 exactly as the driver launches it (train_dp + stress in `secondary`).
mkdir -p gpurun_out
export P3D_SYNC_TIMEOUT_S=30
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2n8_bench.json 2> gpurun_out/r2n8_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2n8_bench.err; head -c 600 gpurun_out/r2n8_bench.json
