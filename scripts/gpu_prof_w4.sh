cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cat > /tmp/w4p.py <<'PY'
import torch, sys
sys.path.insert(0,'.')
import msq_b200
from msq_b200 import _tc
W = torch.randn(4096,4096,device='cuda')*0.02
pw = msq_b200.pack_mx_outlier(W, 8,8,'fp4_e2m1','fp4_e2m1',5,[1],32, variant=1)
x = torch.randn(16,4096,device='cuda').to(torch.bfloat16)
for _ in range(3): y = _tc.packed_linear(x, pw)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_w4 -s 2 -c 1 -o gpurun_out/prof_w4 python /tmp/w4p.py > gpurun_out/ncu_w4.log 2>&1
