# ncu --set full captures of this session's kernels (after the plain runs have exited 0 in tools/run_final.sh)
set -x
timeout 100 python tools/bench_realtime.py 3000 > gpurun_out/realtime_v2.json 2>/dev/null; cat gpurun_out/realtime_v2.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:project_normalize_pair -s 6 -c 1 -o gpurun_out/prof_r1_projnorm_pair -f python tools/bench_aux.py --no-train > gpurun_out/ncu_pair.log 2>&1; tail -2 gpurun_out/ncu_pair.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:latency_cluster -s 50 -c 1 -o gpurun_out/prof_r1_realtime -f python tools/bench_realtime.py 20 > gpurun_out/ncu_rt2.log 2>&1; tail -2 gpurun_out/ncu_rt2.log
ls -la gpurun_out/*.ncu-rep | tail -3
