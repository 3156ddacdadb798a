#!/bin/bash
# Round-2 closing evidence run (1 GPU): tests, smoke, bench line, per-kernel tables of the training step.
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest_gpu_summary.log; cat gpurun_out/pytest_gpu_summary.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json; echo
timeout 300 python scripts/gpu_step_kernels.py > gpurun_out/step_kernels.md 2>/dev/null; head -3 gpurun_out/step_kernels.md
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv \
    --log-file gpurun_out/r02_train_launches.csv python scripts/gpu_train_step_once.py > gpurun_out/train_once.log 2>&1
python scripts/launch_summary.py gpurun_out/r02_train_launches.csv 3 40 > gpurun_out/r02_train_step_table.md 2>&1; head -12 gpurun_out/r02_train_step_table.md
du -sh gpurun_out | tail -1
