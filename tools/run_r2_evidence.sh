# Round 2 evidence call (1 GPU): GPU tests, smoke, the bench line, then (each after its plain run exited 0) the ncu
# launch lists and one --set full capture per top kernel.  Everything lands in gpurun_out/r2ev_*; the summaries that
# are judged are copied into profiles/ afterwards (tools/ncu_summary.py on the CPU box).
#   gpurun --timeout 900 -- 'bash tools/run_r2_evidence.sh'
mkdir -p gpurun_out
export P3D_SYNC_TIMEOUT_S=20
O=gpurun_out/r2ev
timeout 400 python -m pytest tests -m gpu -x -q > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -4 ${O}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 ${O}_smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; tail -2 ${O}_bench.err
timeout 100 python tools/train_steps.py 4096 bf16 20 > ${O}_train4096.txt 2>&1; tail -1 ${O}_train4096.txt
timeout 100 python tools/train_steps.py 64 bf16 20 > ${O}_train64.txt 2>&1; tail -1 ${O}_train64.txt
P3D_LAT_STAMPS=1 timeout 100 python tools/bench_latency.py 1 8 64 1024 4096 > ${O}_latency.txt 2>&1; tail -6 ${O}_latency.txt
# ---- launch lists (gpu__time_duration only)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > ${O}_ncu_bench.log 2>&1; echo "ncu bench rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${O}_launches_train4096.csv python tools/train_steps.py 4096 bf16 2 > ${O}_ncu_train.log 2>&1; echo "ncu train rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_train64.csv python tools/train_steps.py 64 bf16 2 > ${O}_ncu_train64.log 2>&1; echo "ncu train64 rc=$?"
# ---- traffic of the shipped fused kernel (bench.py's roofline.traffic)
timeout 150 python tools/capture_traffic.py > ${O}_traffic.txt 2>&1; tail -1 ${O}_traffic.txt
# ---- one --set full capture per top kernel
timeout 150 ncu --set full --clock-control none --import-source on -k regex:mlp_forward_tc -s 2 -c 1 -o ${O}_mlp_tc -f python tools/forward_once.py 1048576 3 > ${O}_ncu_mlp_tc.log 2>&1; echo "ncu mlp_tc rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 34 -c 8 -o ${O}_train_gemm -f python tools/train_steps.py 4096 bf16 2 > ${O}_ncu_train_gemm.log 2>&1; echo "ncu train gemm rc=$?"
timeout 100 ncu --set full --clock-control none --import-source on -k regex:latency_grid -s 50 -c 1 -o ${O}_latency_grid -f python tools/bench_latency.py 1 > ${O}_ncu_lat.log 2>&1; echo "ncu lat rc=$?"
ls -la gpurun_out/*.ncu-rep
