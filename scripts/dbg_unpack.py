import os, sys
sys.path.insert(0, '/root/repo')
import torch, msq_b200
for fi, fo in (("fp4_e2m1", "fp8_e4m3"), ("fp4_e2m1", "fp4_e2m1"), ("fp4_e2m1", "fp6_e3m2"), ("int4", "fp8_e4m3")):
    for shp in ((16, 256), (64, 1024), (528, 768), (100, 256), (4096, 4096), (1000, 11008), (4096, 11008)):
        W = torch.randn(shp, device='cuda') * 0.02
        W.view(-1)[::97] *= 30
        for tma in ("1", "0"):
            os.environ["MSQ_QUANT_TMA"] = tma
            a = msq_b200.pack_mx_outlier(W, 8, 8, fi, fo, 2, [0], 16)
            try:
                y = a.dequantize(); torch.cuda.synchronize()
                print(fi, fo, shp, "tma", tma, "path", a.path, "ok", flush=True)
            except Exception as e:
                print(fi, fo, shp, "tma", tma, "path", a.path, "FAIL", str(e)[:80], flush=True)
                torch.cuda.synchronize()
