"""Build libmsq_b200.so in-tree with nvcc for sm_100a (no torch dependency in the library).

    python microscopiq-llm-quantization_b200/build.py [--force]

The .so stays next to this file so it travels to the GPU box with the snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmsq_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# per-file extra flags: the quantizers mirror the reference's op-by-op fp32 arithmetic
EXTRA = {"msq_quant_strided.cu": ["-fmad=false"], "msq_quant_strided_big.cu": ["-fmad=false"], "msq_quant_inner.cu": ["-fmad=false"], "msq_quant_misc.cu": ["-fmad=false"], "msq_pack.cu": ["-fmad=false"], "msq_gptq.cu": ["-fmad=false"],
         "msq_uniform.cu": ["-fmad=false"]}


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path, flags):
    h = hashlib.sha1(" ".join(flags).encode())
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cuh", ".h")) or os.path.join(CSRC, f) == path:
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def _compile_one(src, force):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src[:-3] + ".o")
    flags = ARCH + COMMON + EXTRA.get(src, [])
    stamp = obj + ".sha1"
    dig = _digest(path, flags)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False
    cmd = ["nvcc"] + flags + ["-I", CSRC, "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp, "w") as fh:
        fh.write(dig)
    return obj, True


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile_one(s, force), srcs))
    objs = [o for o, _ in res]
    rebuilt = any(c for _, c in res)
    if rebuilt or not os.path.exists(LIB):
        cmd = ["nvcc"] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print("libmsq_b200.so: %s (%d sources, rebuilt=%s)" % (LIB, len(srcs), rebuilt))
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
