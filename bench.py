#!/usr/bin/env python
"""Benchmark of the fast_moe hot path (gate -> dispatch -> 32-expert FFN -> combine + residual).

Contract: `python bench.py --gpus N --steps K --warmup W` (under torchrun for N > 1) prints ONE JSON line on rank 0.

Workload (default `cfg3`, BASELINE.json configs[2], the largest single-GPU configuration): the MoE path of the
18-layer 32-expert 3M-ASR encoder at batch 64 x 206 frames -> 64 x 50 = 3 200 tokens per layer, BF16, top-1 3M router
with the cat-embed input, SiLU experts 512 -> 1024 -> 512, ff_scale 0.5 and the residual add.  A "step" is one pass of
the batch through the 18 MoE layers (18 distinct weight sets = 1.15 GB, far larger than the 126 MB L2, so no layer's
weights are L2-resident from the previous step).  Attention / convolution / LayerNorm of the encoder are outside this
repo's scope, so each layer's output feeds the next layer directly.

  value   MoE-layer tokens/s = layers x tokens x steps / device time, inputs resident in HBM, whole step replayed as
          one CUDA graph (the product's intended mode: no host work between kernels)
  e2e     same metric through the public Python API with HOST buffers: pinned H2D copy of the step's activations,
          the 18 layer calls, D2H of the result, all inside the timed region
  roofline  expert_ffn kernel (dominant): algorithmic bytes per launch / its CUDA-event duration vs measured HBM peak
  cpu_baseline / --impl reference  the CPU oracle (the reference's forward restated in PyTorch CPU fp32) on host cores
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "3m-asr-inference_b200"

WORKLOADS = {
    # name: (layers, utterances, tokens per utterance)
    "cfg1": dict(layers=1, utts=1, tok_per_utt=50, desc="single fast_moe layer, one 206-frame utterance (50 tokens)"),
    "cfg2": dict(layers=12, utts=1, tok_per_utt=50, desc="12-layer encoder MoE path, batch 1 x 206 frames"),
    "cfg3": dict(layers=18, utts=64, tok_per_utt=50, desc="18-layer encoder MoE path, batch 64 x 206 frames"),
    "cfg3f": dict(layers=18, utts=64, tok_per_utt=206, desc="18-layer encoder MoE path, batch 64, frames as tokens"),
    "big": dict(layers=18, utts=512, tok_per_utt=128, desc="18 layers x 65 536 tokens (compute-bound regime)"),
}
E, D, H, DEMB = 32, 512, 1024, 512


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--compute", default="bf16", choices=["bf16", "tf32"],
                    help="bf16 (default, the headline) or tf32: fp32 activations + fp32 weights, <= 1e-3 parity bar")
    ap.add_argument("--ep", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU exchange: peer-memory kernels (product) or NCCL all-to-all (comparison)")
    ap.add_argument("--block", action="store_true",
                    help="time the Conformer block's feed-forward part instead of the bare layer: norm_ff in front, "
                         "norm_final behind (SURVEY 8 f1); not the headline configuration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            p = json.load(open(path))
            return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
        except Exception:
            pass
    return 6650.0, 1590.0, "fallback"  # /opt/skills/guides/B200_PROFILING.md


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
                     0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}
            while not self._halt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.02)
        except Exception as exc:  # report rather than hide
            self.reasons.add(f"sampler_error:{type(exc).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------------------------------
def cpu_forward_step(oracle, torch, layers_cpu, x, embed):
    """One pass of the batch through the MoE layers on the host: the reference's forward restated (oracle)."""
    cur = x
    for w in layers_cpu:
        cur = oracle.moe_forward(cur, embed, w["Wr"], None, w["W1"], w["b1"], w["W2"], w["b2"], top_k=1,
                                 gate_mode=oracle.GATE_3M, act_type=oracle.ACT_SILU, residual=cur, ff_scale=0.5)["out"]
    return cur


def make_cpu_layers(torch, synth, n_layers):
    out = []
    for li in range(n_layers):
        w = synth.make_weights(20260300 + li, E, D, H, DEMB)
        out.append(dict(Wr=w.Wr, W1=w.W1, b1=w.b1, W2=w.W2, b2=w.b2))
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores.  The reference's own
    sources for the path cannot be built here (TensorRT headers / fmoe_cuda absent), so this is the oracle port."""
    if rank != 0:
        return
    import torch
    oracle = importlib.import_module("oracle.moe_oracle")
    synth = importlib.import_module(PKG + ".synth")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = WORKLOADS[args.workload]
    S = wl["utts"] * wl["tok_per_utt"] * max(args.gpus, 1)
    # bounded sample: at most 4 distinct layer weight sets are materialised (each 128 MiB fp32) and cycled
    n_sets = min(wl["layers"], 4)
    layers = make_cpu_layers(torch, synth, n_sets)
    seq = [layers[i % n_sets] for i in range(wl["layers"])]
    g = torch.Generator().manual_seed(20260003)
    x = torch.randn(S, D, generator=g).bfloat16().float()
    embed = torch.randn(S, DEMB, generator=g).bfloat16().float()
    with torch.no_grad():
        # calibrate, then bound each step's sample (a prefix of the layer stack) so that K + W steps fit in ~150 s
        cpu_forward_step(oracle, torch, seq[:1], x, embed)
        t0 = time.perf_counter()
        cpu_forward_step(oracle, torch, seq[:1], x, embed)
        t_layer = time.perf_counter() - t0
        budget = 150.0 / max(args.steps + args.warmup, 1)
        n_layers = max(1, min(wl["layers"], int(budget / max(t_layer, 1e-6))))
        seq = seq[:n_layers]
        for _ in range(args.warmup):
            cpu_forward_step(oracle, torch, seq, x, embed)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_forward_step(oracle, torch, seq, x, embed)
        dt = time.perf_counter() - t0
    value = n_layers * S * args.steps / dt
    line = {
        "impl": "reference", "metric": "moe_layer_tokens_per_sec", "value": value, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, wl, S),
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of {n_layers} of the {wl['layers']} layers x {S} tokens, "
                                   f"{n_sets} weight sets cycled"},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, wl, S_total):
    return {"workload": f"{args.workload}: {wl['desc']}" + (" + norm_ff / norm_final around every layer" if args.block
                                                            else ""),
            "layers": wl["layers"], "tokens_per_layer": S_total,
            "utterances": wl["utts"] * max(args.gpus, 1), "experts": E, "idim": D, "hidden_units": H, "embed_dim": DEMB,
            "top_k": 1, "gate": "3m softmax->max", "activation": "silu", "ff_scale": 0.5,
            "cache": "inputs larger than L2 (18 x 64 MiB weight sets cycle through a 126 MB L2)",
            "parallelism": "single GPU" if args.gpus == 1 else
            f"ep{args.gpus} (32/{args.gpus} experts per GPU), exchange: " +
            ("peer-memory stores fused into dispatch / FFN kernels" if args.ep == "p2p" else "NCCL all-to-all")}


# ----------------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    ops = importlib.import_module(PKG + ".ops")   # raises if libb200moe.so is missing: no fallback
    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    ep = importlib.import_module(PKG + (".ep" if args.ep == "nccl" else ".ep_p2p")) if world > 1 else None

    wl = WORKLOADS[args.workload]
    tf32 = args.compute == "tf32"
    if tf32 and world > 1:
        raise SystemExit("--compute tf32 is single-GPU")
    L = wl["layers"]
    S = wl["utts"] * wl["tok_per_utt"]          # tokens per layer on THIS rank (weak scaling)
    E_local = E // world

    # ---- random-init weights of the architecture, generated on the device (reference init: xavier gain 0.5, bias 0)
    gen = torch.Generator(device=dev).manual_seed(20260300 + rank)
    gen_shared = torch.Generator(device=dev).manual_seed(20260399)   # router is replicated across EP ranks

    def xavier(shape, fan_in, fan_out, g):
        bound = 0.5 * (6.0 / (fan_in + fan_out)) ** 0.5
        return ((torch.rand(shape, generator=g, device=dev) * 2 - 1) * bound)

    layers = []
    for _ in range(L):
        W1 = xavier((E_local, H, D), H * D, E * D, gen)
        W2 = xavier((E_local, D, H), D * H, E * H, gen)
        if not tf32:
            W1, W2 = W1.bfloat16(), W2.bfloat16()
        b1 = torch.zeros(E_local, H, device=dev)
        b2 = torch.zeros(E_local, D, device=dev)
        Wr = xavier((DEMB + D, E), DEMB + D, E, gen_shared).bfloat16().float()
        experts = ops.fp32_experts(W1, b1, W2, b2) if tf32 else ops.PackedExperts(W1, b1, W2, b2)
        norms = {}
        if args.block:
            norms = {"norm_ff": (1.0 + 0.1 * torch.randn(D, generator=gen_shared, device=dev),
                                 0.1 * torch.randn(D, generator=gen_shared, device=dev)),
                     "norm_final": (1.0 + 0.1 * torch.randn(D, generator=gen_shared, device=dev),
                                    0.1 * torch.randn(D, generator=gen_shared, device=dev))}
            if not tf32:
                norms["Wr_packed_ln"] = ops.pack_router_ln(Wr, *norms["norm_ff"])
        layers.append((Wr, experts, ops.pack_router(Wr), norms))

    # ---- synthetic activations: pinned host copies (for e2e) and device-resident copies (for value)
    g = torch.Generator().manual_seed(20260003 + rank)
    act_dtype = torch.float32 if tf32 else torch.bfloat16
    x_host = torch.randn(S, D, generator=g).bfloat16().to(act_dtype).pin_memory()
    e_host = torch.randn(S, DEMB, generator=g).bfloat16().to(act_dtype).pin_memory()
    out_host = torch.empty(S, D, dtype=act_dtype).pin_memory()
    x_dev = x_host.to(dev)
    e_dev = e_host.to(dev)
    x_stage = torch.empty_like(x_dev)
    e_stage = torch.empty_like(e_dev)
    bufs = [torch.empty_like(x_dev), torch.empty_like(x_dev)]

    ep_ctx = None
    if world > 1 and args.ep == "p2p":
        ep_ctx = ep.EpContext.from_process_group(E_local, D, cap=S, timeout_ms=20000)

    def step(x_in, e_in):
        cur = x_in
        for li, (Wr, experts, Wrp, norms) in enumerate(layers):
            out = bufs[li & 1]
            if ep_ctx is not None:
                ep_ctx.forward(cur, e_in, Wr, None, experts, residual=cur, top_k=1, gate_mode=ops.GATE_3M,
                               act_type=ops.ACT_SILU, ff_scale=0.5, out=out, Wr_packed=Wrp, **norms)
            elif world > 1:
                ep.ep_moe_layer(cur, e_in, Wr, None, experts, num_local_expert=E_local, group=None, top_k=1,
                                gate_mode=ops.GATE_3M, act_type=ops.ACT_SILU, ff_scale=0.5, residual=cur, out=out,
                                Wr_packed=Wrp)
            else:
                ops.moe_layer(cur, e_in, Wr, None, experts, residual=cur, top_k=1, gate_mode=ops.GATE_3M,
                              act_type=ops.ACT_SILU, ff_scale=0.5, out=out, Wr_packed=None if tf32 else Wrp,
                              compute=ops.COMPUTE_TF32 if tf32 else ops.COMPUTE_BF16, **norms)
            cur = out
        return cur

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    for _ in range(max(args.warmup, 3)):
        final = step(x_dev, e_dev)
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    step(x_dev, e_dev)
    torch.cuda.synchronize()
    launches_per_step = ops.launch_count() - n0

    # the peer-memory EP path has no host synchronisation either, so the whole step is captured on every rank
    use_graph = (world == 1 or ep_ctx is not None) and not args.no_graph
    graph = None
    if use_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            final = step(x_dev, e_dev)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()

    def timed(fn, iters, finish=None):
        barrier()
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(iters):
            fn()
        if finish is not None:
            finish()
        ev1.record(stream)
        ev1.synchronize()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- value: device-resident inputs, K steps
    K = args.steps
    ms_total = timed((lambda: graph.replay()) if use_graph else (lambda: step(x_dev, e_dev)), K)
    ms_step = ms_total / K
    tokens_per_step = L * S * world
    value = tokens_per_step / (ms_step * 1e-3)

    # ---- eager device time (no graph) and per-stage CUDA-event timing of the same steps (roofline input)
    ms_eager = timed(lambda: step(x_dev, e_dev), K) / K
    stage_ms, stage_calls = {}, {}
    if world == 1 or ep_ctx is not None:
        ops.profile_enable(True)
        kprof = min(K, 400)
        for _ in range(kprof):
            step(x_dev, e_dev)
        torch.cuda.synchronize()
        stage_ms, stage_calls = ops.profile_read()
        ops.profile_enable(False)

    # ---- e2e: host buffers in, host buffer out, EVERY step, through the public API.  The copies run on their own
    # streams so that step i+1's upload and step i's download overlap step i's / i+1's compute, the way a serving loop
    # would feed the layer; every step still uploads its own inputs from pinned memory and downloads its own result.
    h2d_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)
    st_x = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    st_e = [torch.empty_like(e_dev), torch.empty_like(e_dev)]
    st_o = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    out_hosts = [out_host, torch.empty_like(out_host).pin_memory()]
    ev_up = [torch.cuda.Event(), torch.cuda.Event()]        # upload of slot b finished
    ev_used = [torch.cuda.Event(), torch.cuda.Event()]      # compute has consumed staging slot b
    ev_out = [torch.cuda.Event(), torch.cuda.Event()]       # result of slot b is in st_o[b]
    ev_down = [torch.cuda.Event(), torch.cuda.Event()]      # download of slot b finished
    e2e_i = [0]

    def e2e_step():
        b = e2e_i[0] & 1
        e2e_i[0] += 1
        with torch.cuda.stream(h2d_stream):
            h2d_stream.wait_event(ev_used[b])
            st_x[b].copy_(x_host, non_blocking=True)
            st_e[b].copy_(e_host, non_blocking=True)
            ev_up[b].record(h2d_stream)
        stream.wait_event(ev_up[b])
        if use_graph:   # the graph reads x_dev / e_dev and leaves the result in `final`
            x_dev.copy_(st_x[b], non_blocking=True)
            e_dev.copy_(st_e[b], non_blocking=True)
            ev_used[b].record(stream)
            graph.replay()
            res = final
        else:
            res = step(st_x[b], st_e[b])
            ev_used[b].record(stream)
        stream.wait_event(ev_down[b])
        st_o[b].copy_(res, non_blocking=True)
        ev_out[b].record(stream)
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(ev_out[b])
            out_hosts[b].copy_(st_o[b], non_blocking=True)
            ev_down[b].record(d2h_stream)

    def e2e_finish():   # the last downloads belong to the timed region
        stream.wait_stream(d2h_stream)

    for _ in range(3):
        e2e_step()
    e2e_finish()
    ms_e2e = timed(e2e_step, K, e2e_finish) / K
    e2e_value = tokens_per_step / (ms_e2e * 1e-3)
    clocks = sampler.stop()
    if ep_ctx is not None:
        status = ep_ctx.status()
        if status != 0:   # a kernel gave up waiting for a peer: whatever was timed is not the workload
            print(f"rank {rank}: expert-parallel flag wait timed out (status {status}); no result", file=sys.stderr,
                  flush=True)
            sys.exit(3)

    # ---- roofline of the dominant kernel (expert FFN with the fused combine epilogue)
    hbm_peak, tf_peak, peak_kind = load_peaks()
    roofline = None
    traffic = None   # DRAM bytes per launch of the same kernel from one `ncu --set full` capture (profiles/)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_ffn_traffic.json")))
        if args.workload == tr.get("workload") and world == 1 and not tf32:
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except Exception:
        pass
    es = 4 if tf32 else 2                                                       # bytes per weight / activation element
    w_bytes = 2 * E_local * D * H * es + (E_local * H + E_local * D) * 4        # W1 + W2, fp32 biases (this rank)
    if stage_calls.get("expert_ffn"):
        t_ffn = stage_ms["expert_ffn"] / stage_calls["expert_ffn"] * 1e-3      # seconds per launch
        act_bytes = S * (3 * es * D + 8)                                        # xbuf read, residual read, out write, pos+score
        alg_bytes = w_bytes + act_bytes
        flops = S * 4 * D * H
        hbm_time = alg_bytes / (hbm_peak * 1e9)
        tc_time = flops / (tf_peak * 1e12)
        if hbm_time >= tc_time:
            ach = alg_bytes / t_ffn / 1e9
            roofline = {"kernel": "ffn_kernel", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                        "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_kind,
                        "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": t_ffn * 1e6}
        else:
            ach = flops / t_ffn / 1e12
            roofline = {"kernel": "ffn_kernel", "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                        "frac": ach / tf_peak, "traffic": traffic, "peak_source": peak_kind,
                        "algorithmic_flops_per_launch": flops, "us_per_launch": t_ffn * 1e6}
    # the whole layer (gate + dispatch + expert FFN with the fused combine) against the same peaks: SURVEY section 8(d)'s
    # 9 236 B per token (bf16, top-1) + the expert weights once, over the time a layer takes inside the timed region
    layer_bytes = w_bytes + S * 9236 * (es // 2)
    layer_flops = S * (4 * D * H + 2 * (D + DEMB) * E)
    t_layer = ms_step * 1e-3 / L
    layer_roofline = {
        "bound": "hbm" if layer_bytes / (hbm_peak * 1e9) >= layer_flops / (tf_peak * 1e12) else "tensor",
        "hbm_GBps": layer_bytes / t_layer / 1e9, "hbm_frac": layer_bytes / t_layer / 1e9 / hbm_peak,
        "tensor_TFLOPs": layer_flops / t_layer / 1e12, "tensor_frac": layer_flops / t_layer / 1e12 / tf_peak,
        "algorithmic_bytes_per_layer": layer_bytes, "us_per_layer": t_layer * 1e6}

    # ---- CPU baseline: the oracle on this box's host cores, bounded sample (rank 0, N = 1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        oracle = importlib.import_module("oracle.moe_oracle")   # bench's cpu_baseline leg: allowed oracle use
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n_sets = min(L, 2)
        cpu_layers = [dict(Wr=Wr.cpu(), W1=ex.W1.float().cpu(), b1=ex.b1.cpu(), W2=ex.W2.float().cpu(), b2=ex.b2.cpu())
                      for (Wr, ex, _, _n) in layers[:n_sets]]
        seq = [cpu_layers[i % n_sets] for i in range(L)]
        xc, ec = x_host.float(), e_host.float()
        with torch.no_grad():
            cpu_forward_step(oracle, torch, seq, xc, ec)
            reps, t0 = 0, time.perf_counter()
            while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 50):
                cpu_forward_step(oracle, torch, seq, xc, ec)
                reps += 1
            dt = (time.perf_counter() - t0) / reps
        cpu_baseline = {"value": L * S / dt, "unit": "tokens/s", "cores": cores, "kind": "port",
                        "sample": f"{reps} full steps ({L} layers x {S} tokens) of the CPU oracle, "
                                  f"{n_sets} weight sets cycled, torch {torch.__version__} fp32"}

    if rank == 0:
        line = {
            "metric": "moe_layer_tokens_per_sec", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32" if tf32 else "bf16", "data": "synthetic",
            "config": workload_config(args, wl, S * world),
            "utterances_per_sec": wl["utts"] * world / (ms_step * 1e-3),
            "mode": "cuda_graph" if use_graph else "eager",
            "ms_per_step_eager": ms_eager,
            "us_per_layer": ms_step * 1e3 / L,
            "stage_us_per_layer": {k: (stage_ms[k] / stage_calls[k] * 1e3 if stage_calls.get(k) else None)
                                   for k in stage_ms},
            "roofline": roofline,
            "layer_roofline": layer_roofline,
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "tokens/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": (x_host.numel() + e_host.numel()) * x_host.element_size(),
                    "d2h_bytes_per_step": out_host.numel() * out_host.element_size()},
            "gpu_launches": launches_per_step * K,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if ep_ctx is not None:
        dist.barrier()
        ep_ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
