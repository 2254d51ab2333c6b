# Round-2 opener (written at the end of round 1 without GPU access): persisting-L2 window on the activation scratch of the
# fused inference kernel (P3D_L2_PERSIST=1).  Parity tests on that path, bench A/B, and the DRAM traffic of one launch with
# ncu (dram__bytes_read.sum + dram__bytes_write.sum: 4.1 GB per 2^20-pose launch without the window, 0.34 GB algorithmic).
#   gpurun --timeout 900 -- 'bash tools/run_r2_l2persist.sh'
set -x
mkdir -p gpurun_out
P3D_L2_PERSIST=1 timeout 200 python -m pytest tests/test_gpu_mlp.py -x -q > gpurun_out/r2_l2p_tests.log 2>&1; tail -2 gpurun_out/r2_l2p_tests.log
timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; python -c "import json;d=json.load(open('gpurun_out/r2_bench_default.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'])"
P3D_L2_PERSIST=1 timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_l2p.json 2> gpurun_out/r2_bench_l2p.err; python -c "import json;d=json.load(open('gpurun_out/r2_bench_l2p.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'])"
for V in 0 1; do
  P3D_L2_PERSIST=$V timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:mlp_forward_tc -c 3 --csv \
    --log-file gpurun_out/r2_l2p_traffic_$V.csv python tools/forward_once.py 1048576 2 > gpurun_out/r2_l2p_ncu_$V.log 2>&1; tail -4 gpurun_out/r2_l2p_traffic_$V.csv
done
