mkdir -p gpurun_out
export P3D_SYNC_TIMEOUT_S=30
nvidia-smi topo -m > gpurun_out/r2_n8_topo.txt 2>&1; lscpu | grep -E "^CPU\(s\)|NUMA" >> gpurun_out/r2_n8_topo.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/dp_check.py > gpurun_out/r2_n8_dpcheck.txt 2>&1; tail -3 gpurun_out/r2_n8_dpcheck.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; tail -2 gpurun_out/r2_bench_n8.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n8.json')); print(d['value'], d['e2e']['value'], d['e2e']['predictions_only']['value'], d['e2e']['per_rank'], d['e2e']['host_numa']); print(json.dumps(d['secondary'], indent=1))"
