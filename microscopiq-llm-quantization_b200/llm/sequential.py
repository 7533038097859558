"""Layer-sequential GPTQ / RTN drivers and perplexity evaluation -- the callers of the hot path
(SURVEY 8f rows 1-2): llm/llama.py:62-173 (llama_sequential), :176-284 (llama_eval, incl. the
`--nearest` RTN branch :226-253), llm/opt.py:26-128 (opt_sequential).

One implementation serves every HF decoder stack: the first decoder layer is wrapped in a Catcher that
records its positional input and *all* keyword arguments (attention mask, position ids, rotary
`position_embeddings` of transformers >= 4.48, cache position ...) and those are replayed for every layer,
so the driver does not depend on one transformers version (the reference's llama.py does: SURVEY 7).
The quantizer configuration is an argument (the reference hard-codes int2/fp4/block 16, llama.py:130-137).
"""
import time
from types import SimpleNamespace

import torch
import torch.nn as nn

from .. import _ops
from .. import dist as D
from ..utils.modelutils import find_layers
from ..utils.quant import MXQuantizer, quantize_mx_outlier_v1
from .gptq import GPTQ

#: the reference's hard-coded quantizer (llm/llama.py:130-137)
REFERENCE_QCFG = dict(inlier_scale_bits=8, outlier_scale_bits=8, inlier_elem_format="int2", outlier_elem_format="fp4",
                      axes=[0], block_size=16)

LLAMA_TRUE_SEQUENTIAL = [["self_attn.k_proj", "self_attn.v_proj", "self_attn.q_proj"], ["self_attn.o_proj"],
                         ["mlp.up_proj", "mlp.gate_proj"], ["mlp.down_proj"]]


def decoder_parts(model):
    """(layers ModuleList, modules that run before layer 0, final norm or None, lm_head)"""
    if hasattr(model, "model") and hasattr(model.model, "layers"):                       # LLaMA / Mistral
        pre = [m for m in (model.model.embed_tokens, getattr(model.model, "rotary_emb", None)) if m is not None]
        return model.model.layers, pre, model.model.norm, model.lm_head
    if hasattr(model, "model") and hasattr(model.model, "decoder"):                      # OPT
        d = model.model.decoder
        pre = [m for m in (d.embed_tokens, d.embed_positions, getattr(d, "project_in", None)) if m is not None]
        post = [m for m in (getattr(d, "final_layer_norm", None), getattr(d, "project_out", None)) if m is not None]
        return d.layers, pre, (nn.Sequential(*post) if post else None), model.lm_head
    raise TypeError("unsupported model type %s" % type(model))


class _Stop(Exception):
    pass


class LayerMover(object):
    """Moves decoder layers between host and device around the per-layer work (the reference does `layer.to(dev)` /
    `layer.cpu()`, llm/llama.py:94 / :168, which allocates -- and page-faults -- a fresh host copy of every layer).
    Here the host tensors stay where they are: `fetch` uploads them (through one reusable pinned staging buffer, so
    the copy runs at PCIe speed) and `offload` copies the device values back INTO the original host storage.
    With enabled=True (sequential_quantize(overlap_moves=True)) layer i+1 is uploaded and layer i-1 downloaded by a
    background thread on a side stream while layer i is being quantized; ordering is by CUDA events.  Opt-in: on one
    GPU the moves cost 0.035 s per layer through the staging buffer, and the background copies slow the launch-bound
    solves down by more than that (measured: 4 layers 2.30 s without, 2.82 s with overlap)."""

    def __init__(self, layers, dev, enabled=True, shard=False, gather=None):
        """shard=True (under torch.distributed, every rank holding the SAME host model): an upload is split over the
        ranks -- each stages and copies 1/world of the layer's bytes through its own PCIe link and one all-gather over
        NVLink assembles the layer on every GPU (eight ranks each pushing the whole layer through the one host
        otherwise contend for the host's memory bandwidth: 0.51 s of a 1.34 s four-layer run on 8 GPUs).  `gather` =
        all_gather_into_tensor-like callable (tests)."""
        self.layers, self.dev = layers, dev
        self.cuda = dev.type == "cuda"
        self.enabled = enabled and self.cuda
        self.world, self.rank = 1, 0
        self.gather = gather
        if shard and D.is_dist() and not self.enabled:
            self.world, self.rank = D.dist.get_world_size(), D.dist.get_rank()
            if gather is None:
                self.gather = D.dist.all_gather_into_tensor
        self.pending, self.offloads = {}, []
        self.host = {}             # layer index -> {param/buffer name: host tensor} (the original storage)
        self.stage = None
        if self.enabled:
            from concurrent.futures import ThreadPoolExecutor
            self.stream = torch.cuda.Stream(dev)
            self.pool = ThreadPoolExecutor(max_workers=1)

    def _staging(self, nbytes):
        if self.stage is None or self.stage.numel() < nbytes:
            self.stage = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        return self.stage

    @staticmethod
    def shard_plan(nbytes_list, world):
        """byte layout of a sharded upload: tensors at 256-byte aligned offsets of one buffer, the buffer cut into `world`
        equal chunks (multiples of 256 bytes) -> (offsets, chunk bytes)"""
        offs, off = [], 0
        for nb in nbytes_list:
            offs.append(off)
            off += (nb + 255) // 256 * 256
        chunk = ((off + world - 1) // world + 255) // 256 * 256
        return offs, max(chunk, 256)

    @staticmethod
    def stage_shard(tensors, offs, lo, hi, stage):
        """copy the bytes [lo, hi) of the concatenated layout (tensor t at offs[t]) into stage[0 : hi - lo]"""
        for p, off in zip(tensors, offs):
            nb = p.numel() * p.element_size()
            a, b = max(lo, off), min(hi, off + nb)
            if a < b:
                src = p.contiguous().reshape(-1).view(torch.uint8)
                stage[a - lo: b - lo].copy_(src[a - off: b - off])

    def _to_device_sharded(self, i):
        layer = self.layers[i]
        named = [(n, p) for n, p in list(layer.named_parameters()) + list(layer.named_buffers()) if p.device != self.dev]
        if not named:
            return layer
        offs, chunk = self.shard_plan([p.numel() * p.element_size() for _, p in named], self.world)
        lo = self.rank * chunk
        stage = self._staging(chunk) if self.cuda else torch.empty(chunk, dtype=torch.uint8)
        self.stage_shard([p.data for _, p in named], offs, lo, lo + chunk, stage)
        buf = torch.empty(chunk * self.world, dtype=torch.uint8, device=self.dev)
        mine = buf[lo: lo + chunk]
        mine.copy_(stage[:chunk], non_blocking=True)
        self.gather(buf, mine)                                       # NVLink: every rank ends with the whole layer
        host = {}
        with torch.no_grad():
            for (n, p), off in zip(named, offs):
                nb = p.numel() * p.element_size()
                host[n] = p.data
                p.data = buf[off: off + nb].view(p.dtype).reshape(p.shape)
        if self.cuda:
            torch.cuda.current_stream(self.dev).synchronize()         # the staging buffer is reused by the next layer
        self.host[i] = host
        return layer

    def _to_device(self, i):
        layer = self.layers[i]
        if self.world > 1:
            return self._to_device_sharded(i)
        if not self.cuda:
            return layer.to(self.dev)
        host = {}
        tensors = [(n, p) for n, p in list(layer.named_parameters()) + list(layer.named_buffers()) if not p.is_cuda]
        total = sum((p.numel() * p.element_size() + 255) // 256 * 256 for _, p in tensors)
        stage = self._staging(max(total, 1))
        off = 0
        with torch.no_grad():
            for n, p in tensors:
                nb = p.numel() * p.element_size()
                view = stage[off: off + nb].view(p.dtype).reshape(p.shape)
                view.copy_(p.data)                                   # host -> pinned
                host[n] = p.data
                p.data = view.to(self.dev, non_blocking=True)        # pinned -> device at PCIe speed
                off += (nb + 255) // 256 * 256
        torch.cuda.current_stream(self.dev).synchronize()             # the staging buffer is reused by the next layer
        self.host[i] = host
        return layer

    def _to_host(self, i, layer):
        host = self.host.pop(i, None)
        if host is None or not self.cuda:
            return layer.cpu()
        with torch.no_grad():
            for n, p in list(layer.named_parameters()) + list(layer.named_buffers()):
                h = host.get(n)
                if h is None or h.shape != p.shape:
                    p.data = p.data.cpu()
                    continue
                h.copy_(p.data.to(h.dtype))                          # back into the ORIGINAL host storage
                p.data = h
        return layer

    def _upload(self, i):
        torch.cuda.set_device(self.dev)
        with torch.cuda.stream(self.stream):
            layer = self._to_device(i)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return layer, ev

    def prefetch(self, i):
        if self.enabled and 0 <= i < len(self.layers) and i not in self.pending:
            self.pending[i] = self.pool.submit(self._upload, i)

    def fetch(self, i):
        """layer i on the device, usable on the current stream"""
        if not self.enabled:
            return self._to_device(i)
        self.prefetch(i)
        layer, ev = self.pending.pop(i).result()
        torch.cuda.current_stream(self.dev).wait_event(ev)
        return layer

    def _download(self, i, layer, ev):
        torch.cuda.set_device(self.dev)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            self.layers[i] = self._to_host(i, layer)

    def offload(self, i, layer):
        """send layer i back to the host once everything queued on the current stream has used it"""
        if not self.enabled:
            self.layers[i] = self._to_host(i, layer)
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self.offloads.append(self.pool.submit(self._download, i, layer, ev))

    def finish(self):
        for f in self.offloads:
            f.result()
        self.offloads = []
        if self.enabled:
            self.pool.shutdown(wait=True)


class Catcher(nn.Module):
    def __init__(self, module, store):
        super().__init__()
        self.module, self.store = module, store

    def forward(self, inp, *args, **kwargs):
        self.store["inps"].append(inp.detach())
        self.store["args"], self.store["kwargs"] = args, kwargs
        raise _Stop


@torch.no_grad()
def capture_layer_inputs(model, batches, dev):
    """run the embedding part on every calibration batch and stop at decoder layer 0"""
    layers, pre, _, _ = decoder_parts(model)
    for m in pre:
        m.to(dev)
    layers[0] = layers[0].to(dev)
    store = {"inps": [], "args": (), "kwargs": {}}
    layers[0] = Catcher(layers[0], store)
    use_cache = model.config.use_cache
    model.config.use_cache = False
    try:
        for batch in batches:
            ids = batch[0] if isinstance(batch, (tuple, list)) else batch
            try:
                model(ids.to(dev))
            except _Stop:
                pass
    finally:
        layers[0] = layers[0].module
        model.config.use_cache = use_cache
    layers[0] = layers[0].cpu()
    for m in pre:
        m.cpu()
    inps = torch.cat(store["inps"], 0)
    return inps, store["args"], store["kwargs"]


def _run_layer(layer, x, args, kwargs):
    out = layer(x, *args, **kwargs)
    return out[0] if isinstance(out, (tuple, list)) else out


@torch.no_grad()
def sequential_quantize(model, batches, dev, qcfg=None, nearest=False, percdamp=.01, groupsize=-1, act_order=False,
                        static_groups=False, true_sequential=False, hessian_batch=8, verbose=False, stats=None, overlap_moves=False,
                        share_hessians=True, forward_batch=8, release_cache=False, packed_out=None,
                        solve_streams_n=3, shard_moves=True):
    """llama_sequential / opt_sequential (llm/llama.py:62-173, llm/opt.py:26-128).
    nearest=True is the RTN branch (llm/llama.py:226-253): every linear is replaced by
    quantize_mx_outlier_v1(W) without calibration.  Returns {layer name: quantizer}.

    Under torch.distributed (one process per GPU, every rank holding the same model and the same `batches`)
    the path shards where it shards naturally: each rank runs ITS share of the calibration samples through the
    layer, the per-rank Hessians are combined by one sum-allreduce per linear (dist.combine_hessians), the
    independent solves of a layer are spread over the ranks (dist.assign_linears) and the quantized weights are
    broadcast from their owners; RTN spreads the linears of a layer the same way.

    forward_batch: calibration samples pushed through the layer per call (the reference uses 1, llm/llama.py:147;
    the Hessian is a sum over tokens, so batching only changes how many kernels are launched).  share_hessians: one
    Hessian + one factorization for linears that receive the same tensor (q/k/v, gate/up).  release_cache: call
    torch.cuda.empty_cache() after every linear like gptq.py:191 (off: the solver workspaces are reused).
    solve_streams_n: side streams the independent column loops of a layer are queued on (0 = one after the other).
    shard_moves: under torch.distributed every rank uploads 1/world of a layer and an all-gather assembles it
    (LayerMover; needs the same host model on every rank, which the data-parallel path assumes anyway).
    packed_out: a dict that receives {"layers.<i>.<name>": PackedMX} -- the packed artefact of every linear this rank
    solved on the fused path (packed.save_packed writes them to disk, MXLinear.from_packed runs them)."""
    qcfg = dict(REFERENCE_QCFG if qcfg is None else qcfg)
    layers, _, _, _ = decoder_parts(model)
    stats = stats if stats is not None else {}
    stats.update(hessian_s=0.0, solve_s=0.0, forward_s=0.0, comm_s=0.0, move_s=0.0, capture_s=0.0, linears=0, comm_bytes=0)
    quantizers = {}
    world, rank = 1, 0
    if D.is_dist():
        world, rank = D.dist.get_world_size(), D.dist.get_rank()
        lo, hi = D.shard_range(len(batches), rank, world)
        if not nearest:
            assert hi > lo, "fewer calibration batches than ranks"
            batches = batches[lo:hi]
    mover = LayerMover(layers, dev, overlap_moves, shard=shard_moves and dev.type == "cuda")
    solve_streams = [torch.cuda.Stream(dev) for _ in range(max(0, solve_streams_n))] if dev.type == "cuda" else []
    if nearest:
        for i in range(len(layers)):
            layer = mover.fetch(i)
            mover.prefetch(i + 1)
            found = find_layers(layer)
            owners = D.assign_linears({n: tuple(m.weight.shape) for n, m in found.items()}, world)
            mine = []
            for name, lin in found.items():
                q = MXQuantizer()
                q.configure(**qcfg)
                quantizers["layers.%d.%s" % (i, name)] = q
                if owners[name] == rank:
                    mine.append((lin, q))
            axes = qcfg.get("axes", [0])
            axes = [axes] if isinstance(axes, int) else list(axes)
            one_launch = (len(mine) > 0 and axes in ([0], [-2]) and qcfg.get("block_size", 16) in (8, 16, 32)
                          and qcfg.get("inlier_elem_format") is not None and len(mine) <= 16
                          and all(l.weight.dim() == 2 and l.weight.is_cuda for l, _ in mine)
                          and mine[0][1].shared_exp_method == "max" and mine[0][1].round in ("nearest", "floor", "even"))
            if one_launch:          # every linear of the layer this rank owns in ONE launch (k_strided_multi)
                q = mine[0][1]
                outs = _ops.mx_outlier_multi([l.weight.data.float() for l, _ in mine], q.inlier_scale_bits,
                                             q.outlier_scale_bits, q.inlier_elem_format, q.outlier_elem_format, q.std_dev,
                                             q.block_size, q.round, q.flush_fp32_subnorms)
                for (lin, _), o in zip(mine, outs):
                    lin.weight.data = o.to(lin.weight.dtype)
            else:
                for lin, q in mine:
                    W = lin.weight.data
                    lin.weight.data = quantize_mx_outlier_v1(W.float(), q.inlier_scale_bits, q.outlier_scale_bits,
                                                             q.inlier_elem_format, q.outlier_elem_format, q.shared_exp_method,
                                                             q.std_dev, q.axes, q.block_size, q.round, q.flush_fp32_subnorms,
                                                             q.custom_cuda).to(W.dtype)
            stats["linears"] += len(mine)
            for name, lin in found.items():
                D.broadcast_weight(lin, owners[name])
            mover.offload(i, layer)
        mover.finish()
        return quantizers

    t0 = time.perf_counter()
    inps, args, kwargs = capture_layer_inputs(model, batches, dev)
    outs = torch.zeros_like(inps)
    torch.cuda.synchronize()
    stats["capture_s"] = stats.get("capture_s", 0.0) + time.perf_counter() - t0      # once per model, not per layer
    nsamples = inps.shape[0]
    for i in range(len(layers)):
        t0 = time.perf_counter()
        layer = mover.fetch(i)
        mover.prefetch(i + 1)
        stats["move_s"] += time.perf_counter() - t0        # time the compute thread waits for the layer
        full = find_layers(layer)
        if true_sequential and all(n in full for grp in LLAMA_TRUE_SEQUENTIAL for n in grp):
            sequential = LLAMA_TRUE_SEQUENTIAL
        else:
            sequential = [list(full.keys())]
        for names in sequential:
            subset = {n: full[n] for n in names}
            gptq = {}
            for name in subset:
                gptq[name] = GPTQ(subset[name])
                gptq[name].hessian_batch = max(1, hessian_batch // forward_batch)
                gptq[name].verbose = verbose
                gptq[name].release_cache_on_free = release_cache
                gptq[name].emit_packed = packed_out is not None
                gptq[name].quantizer = MXQuantizer()
                gptq[name].quantizer.configure(**qcfg)

            # linears fed by the SAME tensor (q/k/v, gate/up) have identical Hessians: the first one seen in a forward
            # accumulates, the others share its state (GPTQ.share_hessian_with).  `seen` holds a reference to the
            # input for the duration of one forward, so equal (pointer, shape, version) means equal contents.
            seen = {}

            def add_batch(name):
                def tmp(_, inp, out):
                    x = inp[0].data
                    if share_hessians:
                        key = (x.data_ptr(), tuple(x.shape), tuple(x.stride()), x.dtype, x._version)
                        hit = seen.get(key)
                        if hit is None:
                            seen[key] = (name, x)
                        elif hit[0] != name:
                            owner = shared_with.setdefault(name, hit[0])
                            assert owner == hit[0], "input sharing changed between samples"
                            gptq[name].share_hessian_with(gptq[owner])
                            return
                    assert name not in shared_with, "input sharing changed between samples"
                    gptq[name].add_batch(x, out.data)
                return tmp
            shared_with = {}
            handles = [subset[name].register_forward_hook(add_batch(name)) for name in subset]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for j in range(0, nsamples, forward_batch):
                seen.clear()
                outs[j:j + forward_batch] = _run_layer(layer, inps[j:j + forward_batch], args, kwargs)
            seen.clear()
            for name in subset:
                _ = gptq[name].H                     # flush buffered activations
            torch.cuda.synchronize()
            stats["hessian_s"] += time.perf_counter() - t0
            for h in handles:
                h.remove()
            owners = D.assign_linears({n: tuple(m.weight.shape) for n, m in subset.items()}, world)
            if world > 1:
                t0 = time.perf_counter()
                for name in subset:                  # one collective per DISTINCT Hessian (share groups: the owner's)
                    if name in shared_with:
                        continue
                    H, N = D.combine_hessians(gptq[name].H, gptq[name].nsamples, upper_only=False)
                    gptq[name].H, gptq[name].nsamples = H, N
                    stats["comm_bytes"] += H.numel() * 4
                torch.cuda.synchronize()
                stats["comm_s"] += time.perf_counter() - t0
            t0 = time.perf_counter()
            mine = [n for n in subset if owners[n] == rank]
            # factor every distinct Hessian once on the main stream, then queue the independent column loops on side
            # streams (a column kernel keeps one or two warps per SM busy: several of them overlap) and collect them
            main = torch.cuda.current_stream(dev)
            for name in mine:
                gptq[name]._factorize(percdamp, act_order)
            ready = torch.cuda.Event()
            ready.record(main)
            launched = []
            for k, name in enumerate(mine):
                sstream = solve_streams[k % len(solve_streams)] if solve_streams else main
                sstream.wait_event(ready)
                with torch.cuda.stream(sstream):
                    launched.append((name, sstream, gptq[name]._fq_launch(128, percdamp, groupsize, act_order, static_groups)))
            for name, sstream, st in launched:
                with torch.cuda.stream(sstream):
                    gptq[name]._fq_finish(st)
                main.wait_stream(sstream)
                stats["linears"] += 1
                if packed_out is not None and gptq[name].packed is not None:
                    packed_out["layers.%d.%s" % (i, name)] = gptq[name].packed
            for name in subset:
                quantizers["layers.%d.%s" % (i, name)] = gptq[name].quantizer
                gptq[name].free()
            torch.cuda.synchronize()
            stats["solve_s"] += time.perf_counter() - t0
            if world > 1:
                t0 = time.perf_counter()
                for name in subset:
                    D.broadcast_weight(subset[name], owners[name])
                    stats["comm_bytes"] += subset[name].weight.numel() * subset[name].weight.element_size()
                torch.cuda.synchronize()
                stats["comm_s"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        for j in range(0, nsamples, forward_batch):
            outs[j:j + forward_batch] = _run_layer(layer, inps[j:j + forward_batch], args, kwargs)
        torch.cuda.synchronize()
        stats["forward_s"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        mover.offload(i, layer)
        stats["move_s"] += time.perf_counter() - t0
        del layer, gptq
        inps, outs = outs, inps
    t0 = time.perf_counter()
    mover.finish()
    stats["move_s"] += time.perf_counter() - t0
    return quantizers


@torch.no_grad()
def perplexity(model, token_ids, dev, seqlen=None):
    """llama_eval's arithmetic (llm/llama.py:266-282): exp(mean token NLL) over consecutive windows of
    `seqlen` tokens, decoder layers streamed through the device one at a time."""
    seqlen = seqlen or getattr(model, "seqlen", 2048)
    token_ids = token_ids.reshape(1, -1)
    nsamples = token_ids.numel() // seqlen
    batches = [token_ids[:, i * seqlen:(i + 1) * seqlen] for i in range(nsamples)]
    layers, _, norm, lm_head = decoder_parts(model)
    inps, args, kwargs = capture_layer_inputs(model, batches, dev)
    outs = torch.zeros_like(inps)
    for i in range(len(layers)):
        layer = layers[i].to(dev)
        for j in range(nsamples):
            outs[j] = _run_layer(layer, inps[j].unsqueeze(0), args, kwargs)
        layers[i] = layer.cpu()
        inps, outs = outs, inps
    if norm is not None:
        norm.to(dev)
    lm_head.to(dev)
    nlls = []
    for i in range(nsamples):
        h = inps[i].unsqueeze(0)
        if norm is not None:
            h = norm(h)
        logits = lm_head(h)
        shift_logits = logits[:, :-1, :].contiguous().float()
        shift_labels = batches[i][:, 1:].to(dev)
        loss = nn.CrossEntropyLoss()(shift_logits.view(-1, shift_logits.size(-1)), shift_labels.reshape(-1))
        nlls.append(loss.float() * seqlen)
    if norm is not None:
        norm.cpu()
    lm_head.cpu()
    return float(torch.exp(torch.stack(nlls).sum() / (nsamples * seqlen)))


def llama_sequential(model, dataloader, dev, args=None):
    """signature of llm/llama.py:62; `args` replaces the reference's module-global argparse namespace"""
    a = args or SimpleNamespace()
    return sequential_quantize(model, dataloader, dev, nearest=getattr(a, "nearest", False),
                               percdamp=getattr(a, "percdamp", .01), groupsize=getattr(a, "groupsize", -1),
                               act_order=getattr(a, "act_order", False), static_groups=getattr(a, "static_groups", False),
                               true_sequential=getattr(a, "true_sequential", False))


opt_sequential = llama_sequential
