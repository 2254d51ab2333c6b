mkdir -p gpurun_out
P3D_LAT_STAMPS=1 timeout 100 python tools/bench_latency.py 1 > gpurun_out/r2g_latency.txt 2>&1; tail -3 gpurun_out/r2g_latency.txt
P3D_LAT_GRIDLL=0 timeout 100 python tools/bench_latency.py 1 > gpurun_out/r2g_latency_cluster.txt 2>&1; tail -1 gpurun_out/r2g_latency_cluster.txt
timeout 600 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_realtime.py tests/test_gpu_mlp_golden.py -x -q -k "not training and not gradients" > gpurun_out/r2g_tests.log 2>&1; tail -4 gpurun_out/r2g_tests.log
