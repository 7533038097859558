"""PCIe ceilings on this box: pinned H2D, D2H, and both at once (chunked like csrc/msq_host.cu)."""
import time, torch
dev = torch.device("cuda", 0)
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device=dev); d_b = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def up():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def down():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    up(); down()
print("H2D  %.1f GB/s" % (n / t(up) / 1e9))
print("D2H  %.1f GB/s" % (n / t(down) / 1e9))
print("both %.1f GB/s per direction" % (n / t(both) / 1e9))
for chunk in (4 << 20, 16 << 20, 64 << 20):
    def chunked():
        for o in range(0, n, chunk):
            with torch.cuda.stream(s1): d_a[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
            with torch.cuda.stream(s2): h_out[o:o + chunk].copy_(d_b[o:o + chunk], non_blocking=True)
    print("both, %d MiB chunks: %.1f GB/s per direction" % (chunk >> 20, n / t(chunked) / 1e9))
