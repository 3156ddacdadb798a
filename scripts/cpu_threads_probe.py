import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from delora_b200 import synthetic
from oracle import delora_oracle as orc
cfg = synthetic.fov_config(h=64, w=2048)
pair = synthetic.make_pair(0, w_raw=2048)
print("cores", os.cpu_count())
for th in (8, 16, 32, 64, os.cpu_count()):
    torch.set_num_threads(th)
    orc.pair_forward_backward(pair[0], pair[1], pair[3], cfg)
    tm = []
    t0 = time.perf_counter(); orc.pair_forward_backward(pair[0], pair[1], pair[3], cfg, timings=tm); dt = time.perf_counter() - t0
    print(f"threads={th}: {dt:.2f} s/pair", tm[0])
