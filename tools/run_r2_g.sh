mkdir -p gpurun_out
for BO in 0 200 0; do P3D_LAT_BACKOFF_NS=$BO P3D_LAT_STAMPS=1 timeout 100 python tools/bench_latency.py 1 8 > gpurun_out/r2g_latency_bo$BO.txt 2>&1; echo backoff $BO; tail -3 gpurun_out/r2g_latency_bo$BO.txt; done
timeout 600 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_realtime.py -x -q > gpurun_out/r2g_tests.log 2>&1; tail -2 gpurun_out/r2g_tests.log
