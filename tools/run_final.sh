# Round-end verification on one B200: GPU tests, smoke, bench line, stream probes, launch lists (ncu, after the plain runs)
set -x
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 150 python tools/bench_aux.py --no-train > gpurun_out/bench_aux_final.jsonl 2> gpurun_out/bench_aux_final.err
timeout 100 python tools/bench_realtime.py 3000 > gpurun_out/realtime_final.json 2>/dev/null
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_final.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_realtime.csv python tools/bench_realtime.py 20 > gpurun_out/ncu_rt.log 2>&1
grep -c latency_cluster gpurun_out/launches_realtime.csv
