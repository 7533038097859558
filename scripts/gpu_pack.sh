cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_pack.py -x -q 2>&1 | tail -20
timeout 300 python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
import msq_b200
from msq_b200 import _ops
W = (torch.randn(11008, 4096, device='cuda')*0.02)
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
n = W.numel()
for fi, fo in (('fp4_e2m1','fp8_e4m3'), ('int2','fp4')):
    ms = timeit(lambda: msq_b200.pack_mx_outlier(W, 8,8,fi,fo,2,[0],16, check=False))
    p = msq_b200.pack_mx_outlier(W, 8,8,fi,fo,2,[0],16)
    b = 4*n + p.nbytes
    print(f'pack {fi}/{fo} strided bs16: {ms:.3f} ms, {b/ms/1e6:.0f} GB/s algorithmic ({p.nbytes/n:.3f} B/elem out)')
    ms = timeit(lambda: p.dequantize()); print(f'  unpack->f32: {ms:.3f} ms {(4*n+p.nbytes)/ms/1e6:.0f} GB/s')
    ms = timeit(lambda: p.dequantize(torch.bfloat16)); print(f'  unpack->bf16: {ms:.3f} ms {(2*n+p.nbytes)/ms/1e6:.0f} GB/s')
ms = timeit(lambda: msq_b200.pack_mx_outlier(W, 8,8,'fp4_e2m1','fp4_e2m1',5,[1],32, variant=1, check=False)); print('pack inner bs32 fp4/fp4 mxops:', ms)
ms = timeit(lambda: msq_b200.pack_mx(W, 8,'fp4_e2m1',[1],32)); print(f'pack plain inner bs32 fp4: {ms:.3f} ms {(4*n+0.53125*n)/ms/1e6:.0f} GB/s')
ms = timeit(lambda: msq_b200._quantize_mx(W, 8,'fp4_e2m1',axes=[1],block_size=32)); print(f'plain fake-quant inner bs32: {ms:.3f} ms {8*n/ms/1e6:.0f} GB/s')
ms = timeit(lambda: msq_b200.quantize_mx_outlier_v1(W, 8,8,'fp4_e2m1','fp8_e4m3','max',2,[1],32)); print(f'outlier fake-quant inner bs32 (with status sync): {ms:.3f} ms {8*n/ms/1e6:.0f} GB/s')
PY
