mkdir -p gpurun_out
export P3D_SYNC_TIMEOUT_S=30
timeout 900 python -m pytest tests/test_gpu_dp.py -x -q > gpurun_out/r2_dp2_tests.log 2>&1; tail -3 gpurun_out/r2_dp2_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -3 gpurun_out/r2_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2.json')); print(d['value'], d['e2e']['value'], d['e2e']['per_rank'], d['e2e']['host_numa']); print(json.dumps(d['secondary'], indent=1))"
