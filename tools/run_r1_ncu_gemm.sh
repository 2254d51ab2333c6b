# ncu --set full of one forward-layer GEMM of a 32768-pose training step (32768 x 1024 x 1024, MN-major W, bias + column
# sums): the default one-tile-per-CTA kernel and the two-CTAs-per-SM instantiation.  Plain runs first.
set -x
mkdir -p gpurun_out
timeout 60 python tools/gemm_once.py 32768 1024 1024 0 1 1 > gpurun_out/gemm_once_default.txt 2>&1; cat gpurun_out/gemm_once_default.txt
P3D_GEMM_OCC2=1 timeout 60 python tools/gemm_once.py 32768 1024 1024 0 1 1 > gpurun_out/gemm_once_occ2.txt 2>&1; cat gpurun_out/gemm_once_occ2.txt
timeout 120 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -o gpurun_out/prof_r1_gemm32k_default -f python tools/gemm_once.py 32768 1024 1024 0 1 1 > gpurun_out/ncu_gemm_default.log 2>&1; tail -2 gpurun_out/ncu_gemm_default.log
P3D_GEMM_OCC2=1 timeout 120 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -o gpurun_out/prof_r1_gemm32k_occ2 -f python tools/gemm_once.py 32768 1024 1024 0 1 1 > gpurun_out/ncu_gemm_occ2.log 2>&1; tail -2 gpurun_out/ncu_gemm_occ2.log
ls -la gpurun_out/*.ncu-rep | tail -3
