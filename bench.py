#!/usr/bin/env python
"""bench.py -- MicroScopiQ quantization hot path on B200 (contract: see the task statement).

A *step* = one pass of the MX+outlier quantizer (utils/quant.py:147 quantize_mx_outlier_v1,
the call the reference's RTN driver makes per linear, llm/llama.py:238-253) over the seven
linears of one LLaMA-2-7B decoder layer (q,k,v,o 4096x4096; gate,up 11008x4096; down
4096x11008; fp32), MX-FP4 inliers / FP8-E4M3 outliers, blocks of 16 along the output-row
axis -- BASELINE.json configs[1].  metric = GB/s of algorithmic bytes (4 B read + 4 B
written per element), whole job.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

LAYER_SHAPES = [("q_proj", 4096, 4096), ("k_proj", 4096, 4096), ("v_proj", 4096, 4096), ("o_proj", 4096, 4096),
                ("gate_proj", 11008, 4096), ("up_proj", 11008, 4096), ("down_proj", 4096, 11008)]
QCFG = dict(in_sb=8, out_sb=8, in_fmt="fp4_e2m1", out_fmt="fp8_e4m3", std_dev=2, axis=0, bs=16)
WORKLOAD = ("LLaMA-2-7B decoder-layer linears (4x4096x4096, 2x11008x4096, 4096x11008) fp32, "
            "MX-FP4 inlier / FP8-E4M3 outlier fake-quant, block 16 along rows (quantize_mx_outlier_v1)")
BYTES_PER_ELEM = 8.0
METRIC, UNIT = "mx_quant_GBps", "GB/s"


def layer_elems():
    return sum(r * c for _, r, c in LAYER_SHAPES)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def _nvml(self):
        """NVML (the library behind nvidia-smi) answers in microseconds, so the short timed region gets many samples"""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self.stop_flag:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = get_reasons(h)
            self.rows.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for _, b in bits])
            time.sleep(0.002)

    def run(self):
        try:
            return self._nvml()
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "reasons": reasons, "samples": len(self.rows)}


def bind_to_gpu_numa_node(index):
    """Pin this process (and therefore its pinned-host allocations, first touch) to the CPUs of the NUMA node the GPU
    hangs off: with one rank per GPU the host side of the e2e copies otherwise crosses the socket interconnect."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # nvml prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def _time_ms(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _time_ms_device(fn, iters=20, reps=5):
    """device time per call for calls SHORTER than the host side of their own launch (decode-size GEMMs: ~8 us of GPU
    work behind ~20 us of Python + ctypes): `iters` calls are captured into one CUDA graph and the replays timed with
    events.  Falls back to the plain event timing if the capture is refused."""
    try:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (iters * reps)
        del g
        return ms
    except Exception:
        torch.cuda.synchronize()
        return _time_ms(fn, iters)


def extras(dev, dev_w, nel):
    """secondary metrics of BASELINE.json's metric string, rank 0 only, each a few hundred ms:
    quantize+pack GB/s, GPTQ seconds for one LLaMA-2-7B decoder layer (x32 = model), packed GEMM TFLOP/s."""
    import msq_b200
    from msq_b200 import _tc
    out = {}
    # (1) quantize + pack, same 7 linears, kernel time through the C ABI with preallocated planes;
    #     algorithmic bytes = 4 B read + packed bytes written (hi stream: compacted prefix only)
    import ctypes as C
    from msq_b200 import _native as N, _ops, packed
    fi, fo = _ops.fmt4(QCFG["in_fmt"]), _ops.fmt4(QCFG["out_fmt"])
    calls, pbytes = [], 0
    for w in dev_w:
        pre, L, post = _ops.geometry(w.shape, QCFG["axis"])
        pl, wi, hi_w, stride, nseg = packed._alloc_planes(w, pre, L, post, QCFG["bs"], fi, fo, False)
        st = torch.zeros(1, dtype=torch.int32, device=dev)
        calls.append((w, pre, L, post, pl, st))
        p = msq_b200.pack_mx_outlier(w, QCFG["in_sb"], QCFG["out_sb"], QCFG["in_fmt"], QCFG["out_fmt"], QCFG["std_dev"],
                                     [QCFG["axis"]], QCFG["bs"])
        pbytes += p.nbytes
        del p
    path = C.c_int(0)

    def pack_all():
        for w, pre, L, post, pl, st in calls:
            rc = N.lib().msq_quantize_mx_outlier_f32(
                N.ptr(w), None, pre, L, post, QCFG["bs"], QCFG["in_sb"], QCFG["out_sb"], fi[0], fi[1], fi[2], fi[3],
                fo[0], fo[1], fo[2], fo[3], float(QCFG["std_dev"]), 0, 0, 0, None, None, N.ptr(pl["codes"]), N.ptr(pl["e_in"]),
                N.ptr(pl["e_out"]), N.ptr(pl["bitmap"]), N.ptr(pl["hi"]), N.ptr(pl["seg_cnt"]), None, None, N.ptr(st),
                C.byref(path), N.stream_ptr(dev))
            assert rc == 0
    ms_host = _time_ms(pack_all, 30, warm=10)      # the leg follows the PCIe-bound e2e phase: let the clocks come back up first
    # seven launches of 40-120 us each behind ~15 us of Python + ctypes per call: device time from CUDA-graph replays
    # (the event-timed number with the Python launches in the loop is kept beside it)
    ms = _time_ms_device(pack_all, 10, 5)
    bpe = pbytes / nel
    out["pack"] = {"GBps": nel * (4 + bpe) / (ms * 1e-3) / 1e9, "ms_per_layer": ms, "packed_bytes_per_elem": bpe,
                   "frac_of_hbm_peak": nel * (4 + bpe) / (ms * 1e-3) / 1e9 / peaks()[0],
                   "ms_per_layer_with_python_launches": ms_host, "timing": "CUDA-graph replays of the seven launches"}
    del calls
    # (2) GPTQ: Hessians over 128 x 2048 synthetic fp16 tokens + fasterquant for the 7 linears of one layer, the way the
    #     driver runs them: linears fed by the same activations (q/k/v, gate/up) share ONE Hessian and ONE factorization
    msq_b200.GPTQ.verbose = False
    nsamples, seqlen = 128, 2048
    X = {}
    for _, r, c in LAYER_SHAPES:
        if c not in X:
            X[c] = torch.randn(1, seqlen, c, device=dev, generator=torch.Generator(device=dev).manual_seed(c)).half()
    GROUPS = [[0, 1, 2], [3], [4, 5], [6]]          # indices into LAYER_SHAPES: (q, k, v) (o) (gate, up) (down)

    def gptq_layer(share, method):
        msq_b200.GPTQ.hinv_method = method
        t_h = t_s = 0.0
        flops = 0.0
        for grp in (GROUPS if share else [[i] for i in range(len(LAYER_SHAPES))]):
            gs = []
            for i in grp:
                _, r, c = LAYER_SHAPES[i]
                layer = torch.nn.Linear(c, r, bias=False, device=dev, dtype=torch.float16)
                layer.weight.data = dev_w[i].half()
                g = msq_b200.GPTQ(layer)
                g.hessian_batch = 8
                g.release_cache_on_free = False
                g.quantizer = msq_b200.MXQuantizer()
                g.quantizer.configure(8, 8, QCFG["in_fmt"], QCFG["out_fmt"], axes=[0], block_size=QCFG["bs"])
                if gs:
                    g.share_hessian_with(gs[0])
                gs.append(g)
            c = LAYER_SHAPES[grp[0]][2]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(nsamples):
                gs[0].add_batch(X[c], None)
            _ = gs[0].H
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for g in gs:
                g.fasterquant()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            for g in gs:
                g.free()
            t_h += t1 - t0
            t_s += t2 - t1
            flops += 2.0 * nsamples * seqlen * c * c
        msq_b200.GPTQ.hinv_method = "native"
        return t_h, t_s, flops
    gptq_layer(True, "native")                                            # warm-up (allocator, module loading)
    t_h, t_s, flops_h = gptq_layer(True, "native")
    t_h7, t_s7, flops_h7 = gptq_layer(False, "native")
    _w = torch.eye(512, device=dev)
    torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(_w)), upper=True)        # cuSOLVER handles, untimed
    _, t_lib, _ = gptq_layer(False, "reference")
    out["gptq_layer"] = {"hessian_s": t_h, "solve_s": t_s, "layer_s": t_h + t_s,
                         "solve_s_unshared": t_s7, "hessian_s_unshared": t_h7, "solve_s_cusolver_route_unshared": t_lib,
                         "hessian_TFLOPs": flops_h / t_h / 1e12,
                         "note": "kernel-level view of one decoder layer (7 linears, 128x2048 fp16 tokens): 4 distinct Hessians and "
                                 "factorizations (q/k/v and gate/up share theirs, as in the driver), 7 column loops one after the "
                                 "other; *_unshared = 7 Hessians / factorizations as llm/llama.py:139-159 builds them; the "
                                 "driver-level number incl. forwards and moves is extra.gptq_model"}
    # (3) packed-weight GEMM, K = N = 4096, MX-FP4 + outliers, bf16 activations
    gemm = {}
    W = dev_w[0]
    pw = msq_b200.pack_mx_outlier(W, 8, 8, "fp4_e2m1", "fp4_e2m1", 5, [1], 32, variant=1)
    for M in (16, 256, 4096, 8192):
        x = torch.randn(M, 4096, device=dev).to(torch.bfloat16)
        ms = (_time_ms_device if M <= 256 else _time_ms)(lambda: _tc.packed_linear(x, pw), 10)
        gemm["M%d" % M] = {"ms": ms, "TFLOPs": 2.0 * M * 4096 * 4096 / (ms * 1e-3) / 1e12}
        if M >= 4096:
            # the same with a dequantized bf16 copy of the weight kept next to the planes (2 B / weight) + kind::f16 GEMM
            _tc.LARGE_M_WITHOUT_BF16_COPY = False
            ms2 = _time_ms(lambda: _tc.packed_linear(x, pw), 10)
            _tc.LARGE_M_WITHOUT_BF16_COPY = True
            pw._dense_bf16 = None
            gemm["M%d" % M]["with_bf16_copy_of_the_weight"] = {"ms": ms2, "TFLOPs": 2.0 * M * 4096 * 4096 / (ms2 * 1e-3) / 1e12}
    gemm["note"] = ("bf16 activations x packed FP4 weight: activations split into two block-scaled E4M3 terms, tcgen05 "
                    "kind::mxf8f6f4 on the packed codes at every M (no dequantized copy of the weight)")
    out["packed_gemm_k4096_n4096"] = gemm
    # (3b) the artefact the quantizers / GPTQ emit (row blocks of 16 along the output features, FP4 inliers / FP8 outliers,
    # hi stream), bf16 activations: its exact K-blocked MX-FP8 form on the block-scaled kernels vs the row-block kernel
    rb = {}
    pwr = msq_b200.pack_mx_outlier(W, 8, 8, QCFG["in_fmt"], QCFG["out_fmt"], QCFG["std_dev"], [0], QCFG["bs"])
    rb["exact_mxfp8_form"] = _tc.rowblock_mxfp8_operand(pwr) is not None
    for M in (16, 256):
        x = torch.randn(M, 4096, device=dev).to(torch.bfloat16)
        ms = _time_ms_device(lambda: _tc.packed_linear(x, pwr), 10)
        _tc.ROWBLOCK_VIA_MXFP8 = False
        ms_rb = _time_ms_device(lambda: _tc.packed_linear(x, pwr), 10)
        _tc.ROWBLOCK_VIA_MXFP8 = True
        rb["M%d" % M] = {"ms": ms, "TFLOPs": 2.0 * M * 4096 * 4096 / (ms * 1e-3) / 1e12, "rowblock_kernel_ms": ms_rb}
    del pwr
    out["quantizer_artefact_gemm_k4096_n4096"] = rb
    # (4) same weight on the native block-scaled path (tcgen05 kind::mxf8f6f4), MX-FP8 activations
    mx8 = {}
    if _tc.mx8_weight_ok(pw):
        for M in (16, 256, 4096):
            xq = msq_b200.pack_mx(torch.randn(M, 4096, device=dev), 8, "fp8_e4m3", axes=[-1], block_size=32)
            ms = (_time_ms_device if M <= 256 else _time_ms)(lambda: _tc.packed_linear_mx8(xq, pw), 10)
            mx8["M%d" % M] = {"ms": ms, "TFLOPs": 2.0 * M * 4096 * 4096 / (ms * 1e-3) / 1e12}
        # decode size against the library GEMM on the dequantized weight (bf16, 4x the bytes), both as device time
        wb = pw.dequantize(torch.bfloat16)
        xb = torch.randn(16, 4096, device=dev).to(torch.bfloat16)
        mx8["M16_cublas_bf16_ms"] = _time_ms_device(lambda: torch.matmul(xb, wb.t()), 10)
        del wb
        # compute-bound shape the north star quotes the tensor-core fraction on: 8192^3 on the TMA-fed kernel
        # (msq_gemm_mxfp8_tma), against the nominal dense kind::mxf8f6f4 peak and against 2x the measured cuBLAS bf16 rate
        W8 = torch.randn(8192, 8192, device=dev) * 0.02
        pw8 = msq_b200.pack_mx_outlier(W8, 8, 8, "fp4_e2m1", "fp4_e2m1", 5, [1], 32, variant=1)
        del W8
        if _tc.mx8_weight_ok(pw8):
            xq8 = msq_b200.pack_mx(torch.randn(8192, 8192, device=dev), 8, "fp8_e4m3", axes=[-1], block_size=32)
            ms = _time_ms(lambda: _tc.packed_linear_mx8(xq8, pw8), 10)
            tf = 2.0 * 8192.0 ** 3 / (ms * 1e-3) / 1e12
            bf16 = 1357.9                   # B200_PROFILING.md fallback (sustained cuBLAS bf16)
            _pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
            if os.path.exists(_pp):
                bf16 = float(json.load(open(_pp)).get("bf16_tflops_sustained", bf16))
            mx8["M8192_K8192_N8192"] = {"ms": ms, "TFLOPs": tf, "frac_of_nominal_4500_TFLOPs": tf / 4500.0,
                                        "frac_of_2x_measured_bf16": tf / (2.0 * bf16),
                                        "kernel": "k_gemm_mx8_tma (token tile %d)" % _tc.N.lib().msq_gemm_mxfp8_tma_tile_tokens(8192, 8192, 8192)}
            del xq8
        del pw8
        torch.cuda.empty_cache()
    out["blockscaled_mxfp8_gemm_k4096_n4096"] = mx8
    # (5) W4A4: both operands packed FP4 codes on tcgen05 kind::mxf4.block_scale (msq_gemm_mxfp4_tma)
    mx4 = {}
    pw4 = msq_b200.pack_mx(W, 8, "fp4_e2m1", axes=[1], block_size=32)
    for M in (256, 4096):
        xq = msq_b200.pack_mx(torch.randn(M, 4096, device=dev), 8, "fp4_e2m1", axes=[-1], block_size=32)
        ms = _time_ms_device(lambda: _tc.packed_linear_mx4(xq, pw4), 10)
        mx4["M%d" % M] = {"ms": ms, "TFLOPs": 2.0 * M * 4096 * 4096 / (ms * 1e-3) / 1e12}
    del pw4
    W8 = torch.randn(8192, 8192, device=dev) * 0.02
    pw8 = msq_b200.pack_mx(W8, 8, "fp4_e2m1", axes=[1], block_size=32)
    del W8
    xq8 = msq_b200.pack_mx(torch.randn(8192, 8192, device=dev), 8, "fp4_e2m1", axes=[-1], block_size=32)
    ms = _time_ms(lambda: _tc.packed_linear_mx4(xq8, pw8), 10)
    tf = 2.0 * 8192.0 ** 3 / (ms * 1e-3) / 1e12
    mx4["M8192_K8192_N8192"] = {"ms": ms, "TFLOPs": tf, "frac_of_nominal_9000_TFLOPs": tf / 9000.0,
                                "kernel": "k_gemm_mx8_tma2<F4> (kind::mxf4.block_scale.scale_vec::2X, cta_group::2, 256 x %d tiles)"
                                          % _tc.N.lib().msq_gemm_mxfp4_tma_tile_tokens(8192, 8192, 8192)}
    del xq8, pw8
    torch.cuda.empty_cache()
    out["blockscaled_mxfp4_gemm_w4a4"] = mx4
    # (6) the activation side of those linears: [M, K] activations -> packed codes + shared exponents (msq_quantize_mx_act),
    # 8192 x 8192 so that the input does not sit in L2 between replays
    act = {}
    peak = 6543.1
    _pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(_pp):
        peak = float(json.load(open(_pp)).get("hbm_gbs", peak))
    for dt, nm in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        xa = torch.randn(8192, 8192, device=dev).to(dt)
        for fmt, wbytes in (("fp4_e2m1", 0.5), ("fp8_e4m3", 1.0)):
            ms = _time_ms_device(lambda: msq_b200.pack_mx(xa, 8, fmt, axes=[-1], block_size=32), 10)
            gb = 8192.0 * 8192.0 * (xa.element_size() + wbytes + 1.0 / 32) / (ms * 1e-3) / 1e9
            act["%s_to_%s" % (nm, fmt)] = {"ms": ms, "GBps": gb, "frac_of_hbm_peak": gb / peak}
        del xa
    act["note"] = "8192 x 8192 activations, blocks of 32 along K; algorithmic bytes = input + codes + 1 exponent byte per block"
    torch.cuda.empty_cache()
    out["activation_quant"] = act
    # (7) the two together: MXLinear.from_packed(..., fast_activations=True) on bf16 activations (one-launch activation
    # quantizer + block-scaled GEMM on the packed FP4 weight), device time of the whole forward, next to the bf16 library
    # linear on the dequantized weight
    srv = {}
    pws = msq_b200.pack_mx(W, 8, "fp4_e2m1", axes=[1], block_size=32)
    wb = pws.dequantize(torch.bfloat16)
    for a_fmt, nm in (("fp4_e2m1", "w4a4"), ("fp8_e4m3", "w4a8")):
        lin = msq_b200.MXLinear.from_packed(pws, mx_specs={"w_elem_format": "fp4_e2m1", "a_elem_format": a_fmt, "block_size": 32,
                                                           "scale_bits": 8}, fast_activations=True)
        for M in (16, 256, 4096):
            xs = torch.randn(M, 4096, device=dev).to(torch.bfloat16)
            ms = _time_ms_device(lambda: lin(xs), 10)
            srv["%s_M%d" % (nm, M)] = {"ms": ms, "TFLOPs": 2.0 * M * 4096 * 4096 / (ms * 1e-3) / 1e12}
    for M in (16, 256, 4096):
        xs = torch.randn(M, 4096, device=dev).to(torch.bfloat16)
        srv["cublas_bf16_M%d_ms" % M] = _time_ms_device(lambda: torch.nn.functional.linear(xs, wb), 10)
    srv["note"] = "K = N = 4096, bf16 activations in, bf16 out; activation quantization included in the timed forward"
    del wb, pws
    torch.cuda.empty_cache()
    out["mxlinear_serving_k4096_n4096"] = srv
    return out


def _dist_max(dist, dev, *vals):
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def build_llama(layers, dev, hidden=4096, inter=11008, heads=32, vocab=32000, seed=0):
    """random-init LLaMA-2-7B-shaped HF model in fp16 on the host (no checkpoint available), same weights on every rank"""
    import transformers
    cfg = transformers.LlamaConfig(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers, num_attention_heads=heads,
                                   num_key_value_heads=heads, vocab_size=vocab, max_position_embeddings=4096)
    skip = lambda *x, **k: None
    saved = (torch.nn.init.kaiming_uniform_, torch.nn.init.uniform_, torch.nn.init.normal_)
    torch.nn.init.kaiming_uniform_ = torch.nn.init.uniform_ = torch.nn.init.normal_ = skip
    torch.set_default_dtype(torch.float16)
    try:
        model = transformers.LlamaForCausalLM(cfg)
    finally:
        torch.set_default_dtype(torch.float32)
        torch.nn.init.kaiming_uniform_, torch.nn.init.uniform_, torch.nn.init.normal_ = saved
    g = torch.Generator(device=dev).manual_seed(seed)
    with torch.no_grad():
        for _, p in model.named_parameters():
            if p.dim() >= 2:
                p.data = (torch.randn(p.shape, device=dev, generator=g) * 0.02).half().cpu()
            else:
                p.data = torch.ones_like(p.data)
    model.eval()
    return model


def leg_gptq_model(dev, rank, world, dist, layers=4, nsamples=128, seqlen=2048, qcfg=None, tag="fp4_e2m1/fp8_e4m3 block 16"):
    """STRONG scaling of the real calibration path (llm/llama.py:62-173 through msq_b200.llm.sequential): `layers`
    LLaMA-2-7B decoder layers, 128 x 2048 synthetic tokens split over the ranks, per-rank tcgen05 Hessians, NCCL
    allreduce, solves spread over the ranks, quantized weights broadcast.  Wall time, max over ranks."""
    from msq_b200.llm import sequential as S
    qcfg = qcfg or dict(inlier_scale_bits=8, outlier_scale_bits=8, inlier_elem_format=QCFG["in_fmt"],
                        outlier_elem_format=QCFG["out_fmt"], axes=[0], block_size=QCFG["bs"])
    # untimed warm-up on one layer / 8 samples per rank: CUDA module loading, cuBLAS / NCCL initialisation and the
    # allocator's first growth would otherwise be multiplied by 32 / layers in the extrapolation
    warm = build_llama(1, dev)
    warm.seqlen = seqlen
    gw = torch.Generator().manual_seed(4)
    S.sequential_quantize(warm, [torch.randint(0, 32000, (1, seqlen), generator=gw) for _ in range(8 * world)], dev, qcfg=qcfg)
    del warm
    model = build_llama(layers, dev)
    model.seqlen = seqlen
    gc = torch.Generator().manual_seed(3)
    calib = [torch.randint(0, 32000, (1, seqlen), generator=gc) for _ in range(nsamples)]
    stats = {}
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    S.sequential_quantize(model, calib, dev, qcfg=qcfg, stats=stats)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    tq = time.perf_counter() - t0
    agree = True
    if dist is not None:
        w = model.model.layers[-1].mlp.down_proj.weight.data.to(dev).float()
        chk = torch.stack([w.sum(), w.abs().sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        agree = bool(torch.equal(lo, hi))
    keys = ["hessian_s", "solve_s", "forward_s", "comm_s", "move_s", "capture_s"]
    mx = _dist_max(dist, dev, tq, *[stats.get(k, 0.0) for k in keys])
    del model
    cap = mx[-1]
    out = {"quantizer": tag, "layers": layers, "nsamples": nsamples, "seqlen": seqlen, "quantize_s": mx[0],
           "layer_s": (mx[0] - cap) / layers, "llama2_7b_quantize_s": cap + (mx[0] - cap) / layers * 32,
           "llama2_7b_quantize_s_whole_run_scaled": mx[0] / layers * 32,
           "note": "%d decoder layers measured end to end (capture, forwards, Hessians, collectives, solves, host<->device "
                   "layer moves); 32-layer model = the embedding / input capture once (capture_s) + 32 x the per-layer time; "
                   "*_whole_run_scaled = the whole %d-layer run x %g (charges the capture %g times)"
                   % (layers, layers, 32.0 / layers, 32.0 / layers), "ranks_agree": agree}
    for k, v in zip(keys, mx[1:]):
        out[k] = v
    out["allreduce_s"] = out.pop("comm_s")
    nbytes = stats.get("comm_bytes", 0)
    out["allreduce_GBps"] = (nbytes / out["allreduce_s"] / 1e9) if out["allreduce_s"] > 0 and world > 1 else None
    out["linears_solved_rank0"] = stats.get("linears")
    return out


def leg_cfg3(dev, rank, world, dist, dev_w):
    """BASELINE configs[2]: MX-FP6 (fp6_e3m2 in/out, scale_bits 4, block 32, the run_mx_fp6.sh element format) on the
    LLaMA-2-7B linears + the Hessian allreduce across the ranks."""
    import msq_b200
    from msq_b200 import _ops, dist as D
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    outs = [torch.empty_like(w) for w in dev_w]

    def step():
        for w, o in zip(dev_w, outs):
            _ops.mx_outlier(w, 4, 4, "fp6_e3m2", "fp6_e3m2", 2, 0, 32, want=("deq",), check=False, out=o, status=st)
    ms = _time_ms(step, 10)
    nel = sum(w.numel() for w in dev_w)
    res = {"quant_GBps_per_gpu": nel * 8 / (ms * 1e-3) / 1e9, "quantizer": "fp6_e3m2/fp6_e3m2 scale_bits 4 block 32 axes=[0]"}
    del outs
    if dist is not None:
        ar = {}
        for K in (4096, 11008):
            H = torch.randn(K, K, device=dev)
            for _ in range(2):
                D.combine_hessians(H, 16, upper_only=False)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                D.combine_hessians(H, 16, upper_only=False)
            e1.record()
            torch.cuda.synchronize()
            (ms_ar,) = _dist_max(dist, dev, e0.elapsed_time(e1) / 5)
            ar["K%d" % K] = {"ms": ms_ar, "algo_GBps": K * K * 4 / (ms_ar * 1e-3) / 1e9,
                             "bus_GBps": K * K * 4 * 2 * (world - 1) / world / (ms_ar * 1e-3) / 1e9}
            del H
        res["hessian_allreduce"] = ar
    return res


def leg_cfg4(dev, rank, world, dist):
    """BASELINE configs[3]: LLaMA-3-70B linears, one of each big shape, rows sharded over the ranks
    (dist.fasterquant_row_sharded): 8192x8192 and 28672x8192 (K = 8192) and 8192x28672 (K = 28672)."""
    import msq_b200
    from msq_b200 import dist as D
    res = {}
    tokens = 128 * 2048
    for name, rows, cols in (("q_proj", 8192, 8192), ("gate_proj", 28672, 8192), ("down_proj", 8192, 28672)):
        g = torch.Generator(device=dev).manual_seed(rows + cols)
        layer = torch.nn.Linear(cols, rows, bias=False, device=dev, dtype=torch.float16)
        layer.weight.data = (torch.randn(rows, cols, device=dev, generator=g) * 0.02).half()
        gq = msq_b200.GPTQ(layer)
        gq.verbose = False
        gq.release_cache_on_free = False
        gq.quantizer = msq_b200.MXQuantizer()
        gq.quantizer.configure(8, 8, QCFG["in_fmt"], QCFG["out_fmt"], axes=[0], block_size=QCFG["bs"])
        per_rank = tokens // world
        X = torch.randn(1, 8192, cols, device=dev, generator=torch.Generator(device=dev).manual_seed(rank)).half()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(per_rank // 8192):
            gq.add_batch(X, None)
        _ = gq.H
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if dist is not None:
            H, n = D.combine_hessians(gq.H, gq.nsamples, upper_only=False)
            gq.H, gq.nsamples = H, n
            torch.cuda.synchronize()
        t2 = time.perf_counter()
        path = D.fasterquant_row_sharded(gq)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        th, ta, ts = _dist_max(dist, dev, t1 - t0, t2 - t1, t3 - t2)
        res["%s_%dx%d" % (name, rows, cols)] = {"hessian_s": th, "allreduce_s": ta, "solve_s": ts, "path": path,
                                               "hessian_TFLOPs_aggregate": 2.0 * tokens * cols * cols / th / 1e12,
                                               "error": gq.last_error}
        gq.free()
        del layer, gq, X
        torch.cuda.empty_cache()
    return res


def leg_cfg5(dev, rank, world, dist):
    """BASELINE configs[4]: packed MX-FP4 + outlier GEMM sweep, output features split over the ranks (no collective).
    W is packed along K with block 32 (the mx.linear operand layout); A = bf16 (W4A16) and MX-FP8 (block-scaled)."""
    import msq_b200
    from msq_b200 import _tc, dist as D
    res = {}
    for KN in (4096, 8192, 14336, 28672):
        lo, hi = D.shard_range(KN // 128, rank, world)
        n_loc = (hi - lo) * 128
        if n_loc == 0:
            continue
        g = torch.Generator(device=dev).manual_seed(KN + rank)
        W = torch.randn(n_loc, KN, device=dev, generator=g) * 0.02
        pw = msq_b200.pack_mx_outlier(W, 8, 8, "fp4_e2m1", "fp4_e2m1", 5, [1], 32, variant=1)
        del W
        ok8 = _tc.mx8_weight_ok(pw)
        row = {}
        for M in (1, 16, 256, 4096):
            x = torch.randn(M, KN, device=dev, generator=g).to(torch.bfloat16)
            tfn = _time_ms_device if M <= 256 else _time_ms      # decode sizes: device time (CUDA graph), not launch overhead
            ms = tfn(lambda: _tc.packed_linear(x, pw), 5)
            ent = {"a16_ms": None, "a16_TFLOPs": None}
            (ms,) = _dist_max(dist, dev, ms)
            ent["a16_ms"], ent["a16_TFLOPs"] = ms, 2.0 * M * KN * KN / (ms * 1e-3) / 1e12
            if ok8:
                xq = msq_b200.pack_mx(x.float(), 8, "fp8_e4m3", axes=[-1], block_size=32)
                ms8 = tfn(lambda: _tc.packed_linear_mx8(xq, pw), 5)
                (ms8,) = _dist_max(dist, dev, ms8)
                ent["mxfp8_ms"], ent["mxfp8_TFLOPs"] = ms8, 2.0 * M * KN * KN / (ms8 * 1e-3) / 1e12
            row["M%d" % M] = ent
        res["KN%d" % KN] = row
        del pw
        torch.cuda.empty_cache()
    return res


def make_weights(device, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _, r, c in LAYER_SHAPES:
        out.append((torch.randn(r, c, generator=g) * 0.02))
    return out


def cpu_reference_pass(W):
    """the reference's CPU implementation of the path = the torch-CPU oracle port"""
    import msq_oracle as O
    return O.quantize_mx_outlier(W, QCFG["in_sb"], QCFG["out_sb"], QCFG["in_fmt"], QCFG["out_fmt"], QCFG["std_dev"],
                                 QCFG["axis"], QCFG["bs"])


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (the torch-CPU oracle port, all host threads) on
    the SAME config as the B200 arm: one step = the seven linears of one decoder layer (202 M elements)."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    Ws = make_weights(None, 1000)
    for _ in range(args.warmup):
        cpu_reference_pass(Ws[0])                 # warm-up on one q_proj-sized linear (a full step takes seconds)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for W in Ws:
            cpu_reference_pass(W)
    dt = (time.perf_counter() - t0) / args.steps
    v = layer_elems() * BYTES_PER_ELEM / dt / 1e9
    sample = "the full step: all seven linears of the layer, %d steps" % args.steps
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "elements_per_step_per_gpu": layer_elems(), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary metrics (pack, GPTQ layer, GEMM sweep)")
    ap.add_argument("--gptq-layers", type=int, default=4, help="decoder layers of the LLaMA-2-7B-shaped model the GPTQ leg quantizes")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import msq_b200
    from msq_b200 import _ops
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # each rank owns a different decoder layer (layers quantize independently: no collective)
    host_w = [w.pin_memory() for w in make_weights(dev, 1000 + rank)]
    dev_w = [w.to(dev) for w in host_w]
    dev_out = [torch.empty_like(w) for w in dev_w]
    host_out = [torch.empty_like(w).pin_memory() for w in host_w]
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    nel = layer_elems()

    # the seven linears go through ONE launch (msq_quantize_mx_outlier_multi_f32: the RTN driver's per-layer call)
    plan = _ops.MultiQuantPlan(dev_w, dev_out, QCFG["in_sb"], QCFG["out_sb"], QCFG["in_fmt"], QCFG["out_fmt"],
                               QCFG["std_dev"], QCFG["bs"], status=status)

    def step_device():
        plan.launch()

    def step_device_per_tensor():
        for w, o in zip(dev_w, dev_out):
            _ops.mx_outlier(w, QCFG["in_sb"], QCFG["out_sb"], QCFG["in_fmt"], QCFG["out_fmt"], QCFG["std_dev"],
                            QCFG["axis"], QCFG["bs"], want=("deq",), check=False, out=o, status=status)

    def step_e2e():
        """public API with HOST buffers: msq_b200.quantize_mx_outlier_host (the reference's quantize_mx_outlier_v1
        signature on a CPU-resident weight) -> C ABI msq_host_quantize_mx_outlier_f32.  Pinned host tensors in, pinned
        host tensors out; inside the library every linear is cut into 16 MiB chunks whose upload, kernel and download
        overlap on three streams.  Every byte of every step crosses PCIe inside the timed region."""
        for hw, ho in zip(host_w, host_out):
            msq_b200.quantize_mx_outlier_host(hw, QCFG["in_sb"], QCFG["out_sb"], QCFG["in_fmt"], QCFG["out_fmt"], "max",
                                              QCFG["std_dev"], [QCFG["axis"]], QCFG["bs"], out=ho, sync=False)
        msq_b200.host_sync()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step_device()
    # keep the same load running (untimed) so the clock samples describe the steady state the
    # timed steps run in; the timed region itself is exactly --steps steps
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 0.6:
        for _ in range(20):
            step_device()
        torch.cuda.synchronize()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.summary() if rank == 0 else None
    ms_per_tensor = _time_ms(step_device_per_tensor, args.steps) if rank == 0 else 0.0     # seven launches, for comparison
    ms_direct = 0.0
    if rank == 0:      # the round-1/early-round-2 direct-load kernel (k_strided_multi) on the same step, for comparison
        os.environ["MSQ_QUANT_TMA"] = "0"
        ms_direct = _time_ms(step_device, args.steps)
        os.environ.pop("MSQ_QUANT_TMA")
    assert int(status.item()) == 0, "quantizer raised status %d" % int(status.item())

    # end to end through the public API with host buffers
    for _ in range(2):
        step_e2e()
    barrier()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps

    t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s = float(t[0]), float(t[1])
    legs = {}
    if not args.no_extras:
        # every rank takes part in these: they time the paths that shard with their collectives inside
        legs["gptq_model"] = leg_gptq_model(dev, rank, world, dist, layers=args.gptq_layers)
        legs["cfg3_mxfp6"] = leg_cfg3(dev, rank, world, dist, dev_w)
        legs["cfg4_llama3_70b_row_sharded"] = leg_cfg4(dev, rank, world, dist)
        legs["cfg5_gemm_sweep_n_split"] = leg_cfg5(dev, rank, world, dist)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    ms_step = ms / args.steps
    bytes_step = nel * BYTES_PER_ELEM
    value = world * bytes_step / (ms_step * 1e-3) / 1e9
    peak, peak_src = peaks()
    per_gpu = bytes_step / (ms_step * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "elements_per_step_per_gpu": nel, "l2": "inputs larger than L2 (810 MB read + 810 MB written per step)",
                   "sharding": "one decoder layer per rank, no collective", "host_numa_node_rank0": numa},
        "clocks": clocks,
        "e2e": {"value": world * bytes_step / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": nel * 4, "d2h_bytes_per_step": nel * 4,
                "ms_per_step": e2e_s * 1e3},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": peak, "unit": "GB/s", "frac": per_gpu / peak,
                     "traffic": None,
                     "kernel": "k_strided_tma<BS=16, TW=256> (persistent, TMA-staged; one launch = the seven linears)",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": "8 B x 202 375 168 elements = 1.619 GB",
                     "seven_separate_launches_GBps": bytes_step / (ms_per_tensor * 1e-3) / 1e9,
                     "direct_load_kernel_k_strided_multi_GBps": bytes_step / (ms_direct * 1e-3) / 1e9},
    }
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        line["roofline"]["traffic"] = json.load(open(tr)).get("k_strided_tma_bytes_per_launch")
    if not args.no_extras:
        line["extra"] = extras(dev, dev_w, nel)
        line["extra"].update(legs)
    if not args.no_cpu_baseline and world == 1:      # reported on rank 0 at N=1 only
        torch.set_num_threads(os.cpu_count())
        W = host_w[0]
        cpu_reference_pass(W)
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < 10.0 or n < 2:
            cpu_reference_pass(W)
            n += 1
        dt = (time.perf_counter() - t0) / n
        line["cpu_baseline"] = {"value": W.numel() * BYTES_PER_ELEM / dt / 1e9, "unit": UNIT, "cores": os.cpu_count(),
                                "kind": "port", "sample": "q_proj 4096x4096 (1/12 of a step), %d passes" % n}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
