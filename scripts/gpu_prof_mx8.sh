cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cat > /tmp/mx.py <<'PY'
import torch, sys
sys.path.insert(0,'.')
import msq_b200
from msq_b200 import _tc
K = 8192
W = torch.randn(K, K, device='cuda')*0.02
pw = msq_b200.pack_mx_outlier(W, 8,8,'fp4_e2m1','fp4_e2m1',5,[1],32, variant=1)
x = torch.randn(16, K, device='cuda')
xq = msq_b200.pack_mx(x, 8, 'fp8_e4m3', axes=[-1], block_size=32)
for _ in range(3):
    y = _tc.packed_linear_mx8(xq, pw)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_mx8 -s 2 -c 1 -o gpurun_out/prof_mx8 -f python /tmp/mx.py > gpurun_out/ncu_mx8.log 2>&1
tail -3 gpurun_out/ncu_mx8.log
