# round-2 profile collection (run under gpurun, 1 GPU): plain bench line, ncu launch list of one step, ncu --set full
# of the tensor-core conv / wgrad launches of one step and of the block-0 forward
set -x
F="--steps 2 --warmup 3 --min-timed-s 0 --no-logmel --no-other-configs --no-cpu-baseline --no-library-baseline --no-dropin --no-fixed-global"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err || exit 1
python bench.py $F > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 360 -c 130 --csv --log-file gpurun_out/r02_launches.csv python bench.py $F > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_tc_kernel|conv0_win_fwd" -s 14 -c 7 -o gpurun_out/r02_conv_step python bench.py $F > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
