cd $GRAFT_REPO_ROOT
timeout 300 python - <<'PY'
import torch, time
def t(f, n=3):
    f(); torch.cuda.synchronize(); t0=time.time()
    for _ in range(n): r=f()
    torch.cuda.synchronize(); return (time.time()-t0)/n*1e3, r
for K in (4096, 11008):
    X = torch.randn(8192, K, device='cuda'); H = X.t()@X/8192; H += 0.01*H.diag().mean()*torch.eye(K, device='cuda')
    ms1, U1 = t(lambda: torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(H)), upper=True))
    def alt():
        Lf = torch.linalg.cholesky(H.flip(0,1))
        R = Lf.flip(0,1)                      # upper, H = R R^T
        return torch.linalg.solve_triangular(R, torch.eye(K, device='cuda'), upper=True)
    ms2, U2 = t(alt)
    ms3, _ = t(lambda: torch.linalg.cholesky(H))
    print(K, 'trio', ms1, 'flip+trtri', ms2, 'one cholesky', ms3, 'rel diff', float((U1-U2).abs().max()/U1.abs().max()))
PY
