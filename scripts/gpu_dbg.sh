cd $GRAFT_REPO_ROOT
timeout 300 python - <<'PY'
import sys, torch, json
sys.path[:0] = ['.', 'oracle', 'tests']
import msq_b200, msq_oracle as O
from util import load_npz, load_json, mismatches
from msq_b200 import _native as N, _ops
msq_b200.GPTQ.verbose = False
z = load_npz("gptq.npz"); cs = load_json("gptq_cases.json")
for n in (0, 2):
    c = cs[n]
    W = z["W_%d" % n]; Hinv = z["Hinv_%d" % n]
    cfg = dict(inlier_scale_bits=8, outlier_scale_bits=8, inlier_elem_format=c["in_fmt"], outlier_elem_format=c["out_fmt"], block_size=c["bs"])
    Qo, loss_o, info = O.fasterquant(W, Hinv, cfg, c["blocksize"])
    print('case', n, c, 'oracle cols zeros<k:', info['columns_with_zeros_lt_k'], 'k[:8]', info['k'][:8], 'oracle vs ref mism', mismatches(Qo, z["Q_%d"%n]))
    layer = torch.nn.Linear(W.shape[1], W.shape[0], bias=False).cuda(); layer.weight.data = W.clone().cuda()
    gq = msq_b200.GPTQ(layer); gq.quantizer = msq_b200.MXQuantizer(); gq.quantizer.configure(8,8,c["in_fmt"],c["out_fmt"],axes=[0],block_size=c["bs"])
    # replicate _solve_fused but dump counters
    q = gq.quantizer; Wd = W.cuda().clone(); Hd = Hinv.cuda().contiguous(); rows, cols = Wd.shape
    fi, fo = _ops.fmt4(q.inlier_elem_format), _ops.fmt4(q.outlier_elem_format)
    Q = torch.empty_like(Wd); Err = torch.empty(rows,128,device='cuda'); loss = torch.zeros(1,dtype=torch.float64,device='cuda')
    counters = torch.zeros(2*cols+2, dtype=torch.int32, device='cuda')
    zeros, kcnt, status, viol = counters[:cols], counters[cols:2*cols], counters[2*cols:2*cols+1], counters[2*cols+1:]
    L, s = N.lib(), N.stream_ptr(Wd.device)
    rc = L.msq_gptq_block(N.ptr(Wd), Wd.stride(0), N.ptr(Q), Q.stride(0), N.ptr(Hd), Hd.stride(0), rows, 0, cols, N.ptr(Err), N.ptr(loss), N.ptr(zeros), N.ptr(kcnt), N.ptr(status), int(q.block_size), 8, 8, fi[0],fi[1],fi[2],fi[3],fo[0],fo[1],fo[2],fo[3], 2.0, 0, 0, s)
    torch.cuda.synchronize()
    print(' rc', rc, 'status', status.item(), 'zeros[:8]', zeros[:8].tolist(), 'kcnt[:8]', kcnt[:8].tolist(), 'mism vs oracle', mismatches(Q, Qo), 'loss', loss.item(), loss_o)
    zo = [(Qo[:, j] == 0).sum().item() for j in range(8)]
    print(' oracle zeros[:8]', zo)
PY
