cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# launch list of the default bench command (timed region + extras), device time per launch
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
# pack kernel, full set
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_strided|k_unpack' -s 2 -c 2 -o gpurun_out/prof_pack_cur python /tmp/pk.py > gpurun_out/ncu_pack.log 2>&1
ls -la gpurun_out
