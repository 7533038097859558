"""cProfile of bench.py's gptq_model leg (second run; the first warms allocator / module loading)"""
import os, sys, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
r = bench.leg_gptq_model(dev, 0, 1, None, layers=4)
print({k: round(v, 3) for k, v in r.items() if isinstance(v, float)})
pr = cProfile.Profile()
pr.enable()
r = bench.leg_gptq_model(dev, 0, 1, None, layers=4)
pr.disable()
print({k: round(v, 3) for k, v in r.items() if isinstance(v, float)})
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
