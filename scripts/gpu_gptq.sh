cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gptq.py tests/test_gpu_gemm.py -x -q 2>&1 | tail -12
exit 0
timeout 600 python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
import msq_b200
msq_b200.GPTQ.verbose = False
for rows, cols in [(4096,4096),(11008,4096),(4096,11008)]:
    layer = torch.nn.Linear(cols, rows, bias=False).cuda().half()
    gq = msq_b200.GPTQ(layer); gq.quantizer = msq_b200.MXQuantizer(); gq.quantizer.configure(8,8,'int2','fp4',axes=[0],block_size=16)
    X = torch.randn(1, 2048, cols, device='cuda').half()
    torch.cuda.synchronize(); t0=time.time()
    for _ in range(8): gq.add_batch(X, None)
    torch.cuda.synchronize(); t1=time.time()
    H = gq.H.clone()
    t2=time.time(); gq.fasterquant(); torch.cuda.synchronize(); t3=time.time()
    # breakdown
    Hd = H + 0.01*torch.mean(torch.diag(H))*torch.eye(cols, device='cuda')
    torch.cuda.synchronize(); t4=time.time()
    Hinv = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(Hd)), upper=True)
    torch.cuda.synchronize(); t5=time.time()
    print(f'{rows}x{cols}: add_batch x8 {1e3*(t1-t0):.1f} ms; fasterquant {1e3*(t3-t2):.1f} ms (path {gq.last_path}, err {gq.last_error:.3f}); cholesky trio alone {1e3*(t5-t4):.1f} ms')
PY
