timeout 900 python -m pytest tests/test_gpu_mirror.py -q -m gpu -x 2>&1 | tail -30
timeout 600 python bench.py --steps 50 --warmup 5 --cpu-pairs 1 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.err
