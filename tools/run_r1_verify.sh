# Round-1 closing verification of the files edited in the last session: remaining GPU test modules, smoke(), the full
# default bench line, and the end-to-end pipeline at a second chunk size.
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_realtime.py tests/test_gpu_geometry.py -x -q > gpurun_out/last_rest_tests.log 2>&1; tail -3 gpurun_out/last_rest_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1; tail -2 gpurun_out/last_smoke.log
timeout 240 python bench.py > gpurun_out/last_bench_full.json 2> gpurun_out/last_bench_full.err; echo "bench rc=$?"
P3D_PIPE_CHUNK=32768 timeout 90 python bench.py --steps 5 --no-secondary --no-cpu-baseline > gpurun_out/last_bench_chunk32k.json 2> gpurun_out/last_bench_chunk32k.err; echo "bench rc=$?"
python - <<'PY'
import json
for n in ("full", "chunk32k"):
    try:
        d = json.load(open(f"gpurun_out/last_bench_{n}.json"))
        print(n, "value", round(d["value"] / 1e6, 1), "e2e", round(d["e2e"]["value"] / 1e6, 1), "pred-only", round(d["e2e"]["predictions_only"]["value"] / 1e6, 1), d["e2e"]["outputs_match_device_path"], d.get("clocks"))
    except Exception as e:
        print(n, "unreadable:", e)
PY
