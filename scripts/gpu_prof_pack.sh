cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cat > /tmp/pk.py <<'PY'
import torch, sys
sys.path.insert(0,'.')
import msq_b200
W = torch.randn(4096,4096,device='cuda')*0.02
for _ in range(3):
    p = msq_b200.pack_mx_outlier(W, 8,8,'fp4_e2m1','fp8_e4m3',2,[0],16, check=False)
    y = p.dequantize()
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_strided|k_unpack' -s 2 -c 2 -o gpurun_out/prof_pack_cur -f python /tmp/pk.py > gpurun_out/ncu_pack.log 2>&1
bash scripts/gpu_packtime.sh 2>&1 | tail -6
