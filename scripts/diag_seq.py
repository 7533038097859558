"""where does the time of the layer-sequential driver go?  (2 LLaMA-2-7B-shaped layers, second run reported)"""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import msq_b200
from msq_b200.llm import sequential as S
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
model = bench.build_llama(layers, dev)
model.seqlen = 2048
gc = torch.Generator().manual_seed(3)
calib = [torch.randint(0, 32000, (1, 2048), generator=gc) for _ in range(128)]
qcfg = dict(inlier_scale_bits=8, outlier_scale_bits=8, inlier_elem_format="fp4_e2m1", outlier_elem_format="fp8_e4m3", axes=[0], block_size=16)
import copy
for name, kw in (("warmup", {}), ("default", {}), ("fb1", dict(forward_batch=1)), ("noshare", dict(share_hessians=False)),
                 ("overlap", dict(overlap_moves=True)), ("relcache", dict(release_cache=True))):
    m = copy.deepcopy(model)
    st = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    S.sequential_quantize(m, calib, dev, qcfg=qcfg, stats=st, **kw)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    print(name, "total %.3f" % t, {k: round(v, 3) if isinstance(v, float) else v for k, v in st.items()}, flush=True)
    del m
# pieces of one layer
def tm(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
from msq_b200 import _tc
x = torch.randn(8, 2048, 4096, device=dev).half()
H = torch.zeros(4096, 4096, device=dev)
X2 = x.reshape(-1, 4096)
print("hessian K=4096 T=16384 fp16: %.3f ms" % (1e3 * tm(lambda: _tc.hessian_accumulate(H, X2, 0.1, 0.9))))
H2 = torch.zeros(11008, 11008, device=dev)
X3 = torch.randn(16384, 11008, device=dev).half()
print("hessian K=11008 T=16384 fp16: %.3f ms" % (1e3 * tm(lambda: _tc.hessian_accumulate(H2, X3, 0.1, 0.9))))
for n in (4096, 11008):
    Hs = torch.randn(n, 2 * n, device=dev)
    Hs = (Hs @ Hs.t()) / (2 * n) + 0.01 * torch.eye(n, device=dev)
    print("cholinv n=%d: %.3f ms" % (n, 1e3 * tm(lambda: _tc.cholinv_upper(Hs.clone(), check=False), 3)))
    print("  torch trio n=%d: %.3f ms" % (n, 1e3 * tm(lambda: torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(Hs))), 2)))
