# Round 2, 2-GPU call: the bench line at N = 2 exactly as the driver launches it (train_dp + stress in `secondary`), the
# reference arm under torchrun (rank 0 only), and the 2-rank data-parallel parity test.
mkdir -p gpurun_out
export P3D_SYNC_TIMEOUT_S=30
O=gpurun_out/r2n2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; tail -3 ${O}_bench.err
timeout 300 python -m pytest tests/test_gpu_dp.py -x -q -k peer-memory > ${O}_dp_test.log 2>&1; echo "dp test rc=$?"; tail -3 ${O}_dp_test.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 tools/bench_dp_parts.py 64 4096 8192 32768 > ${O}_parts.txt 2>&1; grep "^{" ${O}_parts.txt
