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
    ms = _time_ms(pack_all, 10)
    bpe = pbytes / nel
    out["pack"] = {"GBps": nel * (4 + bpe) / (ms * 1e-3) / 1e9, "ms_per_layer": ms, "packed_bytes_per_elem": bpe,
                   "frac_of_hbm_peak": nel * (4 + bpe) / (ms * 1e-3) / 1e9 / peaks()[0]}
    del calls
    # (2) GPTQ: Hessian over 128 x 2048 synthetic fp16 tokens + fasterquant for the 7 linears of one layer
    msq_b200.GPTQ.verbose = False
    nsamples, seqlen = 128, 2048
    t_h = t_s = 0.0
    flops_h = 0.0
    X = {}
    for _, r, c in LAYER_SHAPES:
        if c not in X:
            X[c] = torch.randn(1, seqlen, c, device=dev, generator=torch.Generator(device=dev).manual_seed(c)).half()
    for i, (name, r, c) in enumerate(LAYER_SHAPES):
        layer = torch.nn.Linear(c, r, bias=False, device=dev, dtype=torch.float16)
        layer.weight.data = dev_w[i].half()
        g = msq_b200.GPTQ(layer)
        g.hessian_batch = 8
        g.quantizer = msq_b200.MXQuantizer()
        g.quantizer.configure(8, 8, QCFG["in_fmt"], QCFG["out_fmt"], axes=[0], block_size=QCFG["bs"])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(nsamples):
            g.add_batch(X[c], None)
        _ = g.H
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        g.fasterquant()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if i == 0:      # first call pays cuSOLVER initialisation: repeat it
            layer.weight.data = dev_w[i].half()
            g = msq_b200.GPTQ(layer)
            g.hessian_batch = 8
            g.quantizer = msq_b200.MXQuantizer()
            g.quantizer.configure(8, 8, QCFG["in_fmt"], QCFG["out_fmt"], axes=[0], block_size=QCFG["bs"])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(nsamples):
                g.add_batch(X[c], None)
            _ = g.H
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            g.fasterquant()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
        t_h += t1 - t0
        t_s += t2 - t1
        flops_h += 2.0 * nsamples * seqlen * c * c
        g.free()
    # same solves with the one-Cholesky route to the identical factor (GPTQ.hinv_method = "flip")
    t_flip = 0.0
    msq_b200.GPTQ.hinv_method = "flip"
    for i, (name, r, c) in enumerate(LAYER_SHAPES):
        layer = torch.nn.Linear(c, r, bias=False, device=dev, dtype=torch.float16)
        layer.weight.data = dev_w[i].half()
        g = msq_b200.GPTQ(layer)
        g.hessian_batch = 8
        g.quantizer = msq_b200.MXQuantizer()
        g.quantizer.configure(8, 8, QCFG["in_fmt"], QCFG["out_fmt"], axes=[0], block_size=QCFG["bs"])
        for _ in range(8):
            g.add_batch(X[c], None)
        _ = g.H
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        g.fasterquant()
        torch.cuda.synchronize()
        t_flip += time.perf_counter() - t1
        g.free()
    msq_b200.GPTQ.hinv_method = "reference"
    out["gptq_layer"] = {"hessian_s": t_h, "solve_s": t_s, "layer_s": t_h + t_s, "solve_s_flip_cholesky": t_flip,
                         "hessian_TFLOPs": flops_h / t_h / 1e12,
                         "llama2_7b_quantize_s_est": 32 * (t_h + t_s),
                         "note": "one decoder layer measured (7 linears, 128x2048 fp16 tokens each, layer forward excluded) x 32"}
    # (3) packed-weight GEMM, K = N = 4096, MX-FP4 + outliers, bf16 activations
    gemm = {}
    W = dev_w[0]
    pw = msq_b200.pack_mx_outlier(W, 8, 8, "fp4_e2m1", "fp4_e2m1", 5, [1], 32, variant=1)
    for M in (16, 256, 4096, 8192):
        x = torch.randn(M, 4096, device=dev).to(torch.bfloat16)
        ms = _time_ms(lambda: _tc.packed_linear(x, pw), 10)
        gemm["M%d" % M] = {"ms": ms, "TFLOPs": 2.0 * M * 4096 * 4096 / (ms * 1e-3) / 1e12}
    out["packed_gemm_k4096_n4096"] = gemm
    # (4) same weight on the native block-scaled path (tcgen05 kind::mxf8f6f4), MX-FP8 activations
    mx8 = {}
    if _tc.mx8_weight_ok(pw):
        for M in (16, 256, 4096):
            xq = msq_b200.pack_mx(torch.randn(M, 4096, device=dev), 8, "fp8_e4m3", axes=[-1], block_size=32)
            ms = _time_ms(lambda: _tc.packed_linear_mx8(xq, pw), 10)
            mx8["M%d" % M] = {"ms": ms, "TFLOPs": 2.0 * M * 4096 * 4096 / (ms * 1e-3) / 1e12}
    out["blockscaled_mxfp8_gemm_k4096_n4096"] = mx8
    return out


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
    """--impl reference: the CPU path on the host cores, bounded sample (one 4096x4096 linear per step)"""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    W = torch.randn(4096, 4096, generator=torch.Generator().manual_seed(0)) * 0.02
    for _ in range(args.warmup):
        cpu_reference_pass(W)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_pass(W)
    dt = (time.perf_counter() - t0) / args.steps
    v = W.numel() * BYTES_PER_ELEM / dt / 1e9
    sample = "one q_proj-sized linear (4096x4096 fp32 = 1/12 of a step's elements) per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
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

    def step_device():
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
        "gpu_launches": len(LAYER_SHAPES) * args.steps,
        "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": peak, "unit": "GB/s", "frac": per_gpu / peak,
                     "traffic": None, "kernel": "k_strided<BS=16, OUTLIER, fake-quant>", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": "8 B x rows x cols"},
    }
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        line["roofline"]["traffic"] = json.load(open(tr)).get("k_strided_bytes_per_launch")
    if not args.no_extras:
        line["extra"] = extras(dev, dev_w, nel)
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
