# Round 2, first GPU call: the four opt-in paths written at the end of round 1, one after the other on one box.
#   gpurun --timeout 1700 -- 'bash tools/run_r2_openers.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_box.txt 2>&1
lscpu | grep -E "^CPU\(s\)|NUMA|Model name" >> gpurun_out/r2_box.txt
bash tools/run_r2_xbf16.sh > gpurun_out/r2_open_xbf16.log 2>&1
bash tools/run_r2_tmastore.sh > gpurun_out/r2_open_tmastore.log 2>&1
bash tools/run_r2_fastissue.sh > gpurun_out/r2_open_fastissue.log 2>&1
bash tools/run_r2_l2persist.sh > gpurun_out/r2_open_l2persist.log 2>&1
tail -5 gpurun_out/r2_open_*.log
