# Round-2 opener (written at the end of round 1 without GPU access): host-buffer step with x rounded to bf16 on the host
# (P3D_PIPE_XBF16=1, 256 instead of 320 B per pose across PCIe).  Bit-identity check + end-to-end A/B on the same box,
# host thread counts 2 / 4 / 8 / 16 (the rounding has to keep up with ~150 M poses/s).
#   gpurun --timeout 400 -- 'bash tools/run_r2_xbf16.sh'
set -x
mkdir -p gpurun_out
timeout 120 python tools/check_e2e_xbf16.py > gpurun_out/r2_e2e_default.txt 2>&1; tail -4 gpurun_out/r2_e2e_default.txt
for T in 2 4 8 16; do
  P3D_PIPE_XBF16=1 P3D_PIPE_THREADS=$T timeout 120 python tools/check_e2e_xbf16.py > gpurun_out/r2_e2e_xbf16_t$T.txt 2>&1; tail -4 gpurun_out/r2_e2e_xbf16_t$T.txt
done
