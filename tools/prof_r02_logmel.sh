# round-2 log-mel profile collection (run under gpurun, 1 GPU): both kernels timed without ncu, then one ncu --set full
# capture of each on 8 resident 3-min stereo clips (the same commands exited 0 without ncu immediately before)
set -x
python tools/bench_logmel.py 32 fp32 > gpurun_out/r02_logmel_bench.jsonl || exit 1
python tools/bench_logmel.py 32 tc >> gpurun_out/r02_logmel_bench.jsonl || exit 1
python tools/bench_logmel.py 8 fp32 >> gpurun_out/r02_logmel_bench.jsonl || exit 1
python tools/bench_logmel.py 8 tc >> gpurun_out/r02_logmel_bench.jsonl || exit 1
ncu --set full --clock-control none --import-source on -k regex:logmel -c 1 -s 3 -o gpurun_out/r02_logmel_tc -f python tools/bench_logmel.py 8 tc > gpurun_out/ncu_lm_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:logmel -c 1 -s 3 -o gpurun_out/r02_logmel_fp32 -f python tools/bench_logmel.py 8 fp32 > gpurun_out/ncu_lm_fp32.log 2>&1
tail -1 gpurun_out/ncu_lm_fp32.log
