# Round 2, GPU call F (1 GPU): full GPU test suite, latency stamps, traffic capture, bench.
mkdir -p gpurun_out
export P3D_SYNC_TIMEOUT_S=20
P3D_LAT_STAMPS=1 timeout 100 python tools/bench_latency.py 1 > gpurun_out/r2f_latency.txt 2>&1; tail -3 gpurun_out/r2f_latency.txt
timeout 200 python tools/capture_traffic.py > gpurun_out/r2f_traffic.txt 2>&1; tail -1 gpurun_out/r2f_traffic.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; tail -4 gpurun_out/r2f_tests.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -2 gpurun_out/r2f_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json')); print(d['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['traffic_source'], d['e2e']['value'], d['latency_batch1']); s=d['secondary']; print(s['training_step']); print(s['inference_batch_sweep']); print(s['stress_width4096']); print(s['evaluation']['procrustes_mpjpe']['roofline']['frac'], s['preprocess']['roofline']['frac'], s['realtime_frame'])"
