mkdir -p gpurun_out
export P3D_SYNC_TIMEOUT_S=30
run() { tag=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29560 + RANDOM % 100)) tools/bench_dp_parts.py 4096 > gpurun_out/r2_nccl_$tag.txt 2>&1; grep "^{" gpurun_out/r2_nccl_$tag.txt; }
run default NCCL_DEBUG=WARN
run nvls NCCL_ALGO=NVLS
run nvlstree NCCL_ALGO=NVLSTree
run tree NCCL_ALGO=Tree
run ring_ll128 NCCL_ALGO=Ring NCCL_PROTO=LL128
run ch32 NCCL_MIN_NCHANNELS=32
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29599 tools/bench_dp_parts.py 4096 2>&1 | grep -i "nvls\|algo\|channels" | head -20 > gpurun_out/r2_nccl_info.txt
tail -5 gpurun_out/r2_nccl_info.txt
