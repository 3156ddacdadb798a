timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/train_launches.csv python scripts/gpu_train_step_once.py > gpurun_out/train_once.log 2>&1
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.err
