import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msq_b200 import _tc
n = 512
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(4 * n, n, device="cuda", generator=g)
H = (X.t() @ X) / (4 * n) + 0.01 * torch.eye(n, device="cuda")
for _ in range(2):
    _tc.cholinv_upper(H.clone())
torch.cuda.synchronize()
