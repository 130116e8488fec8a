for g in 32 8; do
GOOEY_B200_WAVE_G=$g,$g,$g,$g ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.per_cycle_active --clock-control none -k regex:wave_kernel -s 8 -c 8 --csv --log-file gpurun_out/ncu_g$g.csv python bench.py --steps 1 --warmup 1 > /dev/null 2>&1
done
