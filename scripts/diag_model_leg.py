"""run bench.py's gptq_model leg several times in one process (run-to-run variance of the driver-level number)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
for i in range(4):
    r = bench.leg_gptq_model(dev, 0, 1, None, layers=4)
    print(i, {k: round(v, 3) for k, v in r.items() if isinstance(v, float)})
