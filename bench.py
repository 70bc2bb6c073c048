#!/usr/bin/env python
"""Benchmark of the fast_moe hot path (gate -> dispatch -> 32-expert FFN -> combine + residual).

Contract: `python bench.py --gpus N --steps K --warmup W` (under torchrun for N > 1) prints ONE JSON line on rank 0
(`--workload sweep` prints one line per sweep point instead).

Workload (default `cfg3`, BASELINE.json configs[2], the largest single-GPU configuration): the MoE path of the
18-layer 32-expert 3M-ASR encoder at batch 64 x 206 frames -> 64 x 50 = 3 200 tokens per layer, BF16, top-1 3M router
with the cat-embed input, SiLU experts 512 -> 1024 -> 512, ff_scale 0.5 and the residual add.  A "step" is one pass of
the batch through the 18 MoE layers (18 distinct weight sets = 1.15 GB, far larger than the 126 MB L2, so no layer's
weights are L2-resident from the previous step).  Attention / convolution of the encoder are outside this repo's scope
(SURVEY.md section 8f), so each layer's output feeds the next layer directly.

  value   MoE-layer tokens/s = layers x (valid) tokens x steps / device time, inputs resident in HBM, whole step
          replayed as one CUDA graph (the product's intended mode: no host work between kernels)
  e2e     same metric through the public Python API with HOST buffers: pinned H2D copy of the step's activations,
          the layer calls, D2H of the result, all inside the timed region
  roofline  expert_ffn kernel (dominant): algorithmic bytes (or flops) per launch / its CUDA-event duration vs the
          measured peak; roofline_stages: the same for the route / gate / dispatch / combine kernels (HBM)
  parity  (N > 1) one untimed check per run: this rank's expert-parallel output of layer 0 against the single-GPU
          all-experts layer on the same tokens -- routing bit-exact, rel-L2 <= 1e-2 -- exit code 4 on failure
  cpu_baseline / --impl reference  the CPU oracle (the reference's forward restated in PyTorch CPU fp32) on host cores

Other workloads: cfg1 / cfg2 (batch 1), cfg3f, big (compute-bound), cfg4 (BASELINE.json configs[3]: 32 utterances per
GPU with seeded lengths of 100-1000 frames, i.e. batch 256 on 8 GPUs, padded rows masked by x_len) and `sweep`
(configs[4]: global tokens 1K..1M x top-1 / top-2 (+ one Zipf-skewed point), split over the N ranks).
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
    # name: layers, utterances PER GPU, tokens per utterance (uniform) or frame range (variable length)
    "cfg1": dict(layers=1, utts=1, tok_per_utt=50, desc="single fast_moe layer, one 206-frame utterance (50 tokens)"),
    "cfg2": dict(layers=12, utts=1, tok_per_utt=50, desc="12-layer encoder MoE path, batch 1 x 206 frames"),
    "cfg3": dict(layers=18, utts=64, tok_per_utt=50, desc="18-layer encoder MoE path, batch 64 x 206 frames"),
    "cfg3x1": dict(layers=1, utts=64, tok_per_utt=50, desc="ONE layer x 3 200 tokens replayed (weights stay in L2: experiments)"),
    "cfg3f": dict(layers=18, utts=64, tok_per_utt=206, desc="18-layer encoder MoE path, batch 64, frames as tokens"),
    "big": dict(layers=18, utts=512, tok_per_utt=128, desc="18 layers x 65 536 tokens (compute-bound regime)"),
    "cfg4": dict(layers=18, utts=32, frames=(100, 1000),
                 desc="18-layer encoder MoE path, 32 utterances per GPU (256 on 8), 100-1000 frames each, x_len-masked padding"),
    "sweep": dict(layers=4, desc="fast_moe layer sweep: global tokens 1K-1M x top-1/top-2 (+ Zipf), 4 weight sets cycled"),
}
SWEEP_TOKENS = (1024, 4096, 16384, 65536, 262144, 1048576)
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
    ap.add_argument("--sustain", type=float, default=0.0,
                    help="additionally replay the step back to back for this many seconds (own clock record)")
    ap.add_argument("--sweep-max", type=int, default=SWEEP_TOKENS[-1], help="largest sweep point (global tokens)")
    ap.add_argument("--sweep-points", default="", help="run only these sweep points: 'tokens:k[:zipf],...'")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--encoder", action="store_true",
                    help="also time the whole Conformer-MoE encoder around the path (SURVEY 8 f1): utterances/s, N = 1")
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
                time.sleep(0.01)
        except Exception as exc:  # report rather than hide
            self.reasons.add(f"sampler_error:{type(exc).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------------------------------
# workload geometry (shared by both arms, so that `config` is identical)
def subsampled_len(frames: int) -> int:
    """Conv2dSubsampling4 of the reference encoder (trainer_3m_fix/layer/subsampling.py): 206 frames -> 50 tokens."""
    return ((frames - 1) // 2 - 1) // 2


def rank_geometry(torch, wl, rank):
    """-> (B, T, x_len or None, valid tokens) of one rank's batch."""
    if "frames" in wl:
        g = torch.Generator().manual_seed(20260004)
        lo, hi = wl["frames"]
        frames = torch.randint(lo, hi + 1, (wl["utts"] * 8,), generator=g)   # the 256 utterances of the 8-GPU batch
        mine = frames[(rank % 8) * wl["utts"]:(rank % 8 + 1) * wl["utts"]]
        lens = torch.tensor([subsampled_len(int(f)) for f in mine], dtype=torch.int32)
        T = int(lens.max())
        return wl["utts"], T, lens, int(lens.sum())
    return wl["utts"], wl["tok_per_utt"], None, wl["utts"] * wl["tok_per_utt"]


def workload_config(args, wl, tokens, padded, utts, top_k=1, extra=None):
    cfg = {"workload": f"{args.workload}: {wl['desc']}" + (" + norm_ff / norm_final around every layer" if args.block
                                                            else ""),
           "layers": wl["layers"], "tokens_per_layer": tokens, "padded_tokens_per_layer": padded,
           "utterances": utts, "experts": E, "idim": D, "hidden_units": H, "embed_dim": DEMB,
           "top_k": top_k, "gate": "3m softmax->max" if top_k == 1 else "naive top-k -> softmax", "activation": "silu",
           "ff_scale": 0.5,
           "cache": "inputs larger than L2 (64 MiB weight sets cycle through a 126 MB L2)",
           "parallelism": "single GPU" if args.gpus == 1 else
           f"ep{args.gpus} (32/{args.gpus} experts per GPU), exchange: " +
           ("peer-memory stores fused into dispatch / FFN kernels" if args.ep == "p2p" else "NCCL all-to-all")}
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------------------------------------------
def cpu_forward_step(oracle, torch, layers_cpu, x, embed, x_len=None, T=None):
    """One pass of the batch through the MoE layers on the host: the reference's forward restated (oracle)."""
    cur = x
    for w in layers_cpu:
        cur = oracle.moe_forward(cur, embed, w["Wr"], None, w["W1"], w["b1"], w["W2"], w["b2"], top_k=1,
                                 gate_mode=oracle.GATE_3M, act_type=oracle.ACT_SILU, residual=cur, ff_scale=0.5,
                                 x_len=x_len, T=T)["out"]
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
    wl = dict(WORKLOADS[args.workload])
    if args.workload == "sweep":   # one representative point: 65 536 global tokens, top-1
        wl.update(utts=1, tok_per_utt=65536 // max(args.gpus, 1))
    n = max(args.gpus, 1)
    geo = [rank_geometry(torch, wl, r) for r in range(n)]
    B, T = sum(g[0] for g in geo), max(g[1] for g in geo)
    x_len = None
    if geo[0][2] is not None:
        x_len = torch.cat([g[2] for g in geo])
    valid = sum(g[3] for g in geo)
    padded = sum(g[0] * g[1] for g in geo)
    S = B * T
    # bounded sample: at most 4 distinct layer weight sets are materialised (each 128 MiB fp32) and cycled
    n_sets = min(wl["layers"], 4)
    layers = make_cpu_layers(torch, synth, n_sets)
    seq = [layers[i % n_sets] for i in range(wl["layers"])]
    g = torch.Generator().manual_seed(20260003)
    x = torch.randn(S, D, generator=g).bfloat16().float()
    embed = torch.randn(S, DEMB, generator=g).bfloat16().float()
    with torch.no_grad():
        # calibrate, then bound each step's sample (a prefix of the layer stack) so that K + W steps fit in ~150 s
        cpu_forward_step(oracle, torch, seq[:1], x, embed, x_len, T)
        t0 = time.perf_counter()
        cpu_forward_step(oracle, torch, seq[:1], x, embed, x_len, T)
        t_layer = time.perf_counter() - t0
        budget = 150.0 / max(args.steps + args.warmup, 1)
        n_layers = max(1, min(wl["layers"], int(budget / max(t_layer, 1e-6))))
        seq = seq[:n_layers]
        for _ in range(args.warmup):
            cpu_forward_step(oracle, torch, seq, x, embed, x_len, T)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_forward_step(oracle, torch, seq, x, embed, x_len, T)
        dt = time.perf_counter() - t0
    value = n_layers * valid * args.steps / dt
    line = {
        "impl": "reference", "metric": "moe_layer_tokens_per_sec", "value": value, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, wl, valid, padded, B),
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of {n_layers} of the {wl['layers']} layers x {valid} tokens, "
                                   f"{n_sets} weight sets cycled"},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
def pin_to_gpu_numa_node(torch, index):
    """One process per GPU: run this rank's host threads -- and so first-touch its pinned staging buffers -- on the NUMA
    node its GPU hangs off, instead of wherever the launcher left it (with 8 ranks on node 0 the host copies of the
    GPUs behind the other socket cross the inter-socket link).  Best effort: returns the node or None."""
    try:
        props = torch.cuda.get_device_properties(index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


class Bench:
    """One rank's state: device, process group, EP context, peaks."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.ops = importlib.import_module(PKG + ".ops")   # raises if libb200moe.so is missing: no fallback
        assert torch.cuda.is_available(), "bench.py needs a CUDA device"
        torch.cuda.set_device(self.local_rank)
        self.numa = pin_to_gpu_numa_node(torch, self.local_rank) if self.world > 1 else None
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.ep_mod = (importlib.import_module(PKG + (".ep" if args.ep == "nccl" else ".ep_p2p"))
                       if self.world > 1 else None)
        self.tf32 = args.compute == "tf32"
        if self.tf32 and self.world > 1:
            raise SystemExit("--compute tf32 is single-GPU")
        self.E_local = E // self.world
        self.hbm_peak, self.tf_peak, self.peak_kind = load_peaks()
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)

    # ---- random-init weights of the architecture, generated on the device (reference init: xavier gain 0.5, bias 0)
    def make_layers(self, L, top_k=1, router_bias=None):
        torch, ops, dev = self.torch, self.ops, self.dev
        gen = torch.Generator(device=dev).manual_seed(20260300 + self.rank)
        gen_shared = torch.Generator(device=dev).manual_seed(20260399)   # router is replicated across EP ranks

        def xavier(shape, fan_in, fan_out, g):
            bound = 0.5 * (6.0 / (fan_in + fan_out)) ** 0.5
            return ((torch.rand(shape, generator=g, device=dev) * 2 - 1) * bound)

        demb = DEMB if top_k == 1 else 0   # NaiveGate (top-k > 1) has no cat-embed input (fmoe/gates.py:51-66)
        layers = []
        for _ in range(L):
            W1 = xavier((self.E_local, H, D), H * D, E * D, gen)
            W2 = xavier((self.E_local, D, H), D * H, E * H, gen)
            if not self.tf32:
                W1, W2 = W1.bfloat16(), W2.bfloat16()
            b1 = torch.zeros(self.E_local, H, device=dev)
            b2 = torch.zeros(self.E_local, D, device=dev)
            Wr = xavier((demb + D, E), demb + D, E, gen_shared).bfloat16().float()
            experts = ops.fp32_experts(W1, b1, W2, b2) if self.tf32 else ops.PackedExperts(W1, b1, W2, b2)
            norms = {}
            if self.args.block:
                norms = {"norm_ff": (1.0 + 0.1 * torch.randn(D, generator=gen_shared, device=dev),
                                     0.1 * torch.randn(D, generator=gen_shared, device=dev)),
                         "norm_final": (1.0 + 0.1 * torch.randn(D, generator=gen_shared, device=dev),
                                        0.1 * torch.randn(D, generator=gen_shared, device=dev))}
                if not self.tf32:
                    norms["Wr_packed_ln"] = ops.pack_router_ln(Wr, *norms["norm_ff"])
            layers.append(dict(Wr=Wr, br=router_bias, experts=experts, Wrp=ops.pack_router(Wr), norms=norms))
        return layers

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, iters, finish=None):
        torch = self.torch
        self.barrier()
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record(self.stream)
        for _ in range(iters):
            fn()
        if finish is not None:
            finish()
        ev1.record(self.stream)
        ev1.synchronize()
        self.barrier()
        ms = ev0.elapsed_time(ev1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def all_sum(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([float(v)], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t)
        return float(t.item())


class Case:
    """One (workload point) on one rank: inputs, layer stack, step function, graph."""

    def __init__(self, b: Bench, L, B, T, x_len, top_k=1, router_bias=None, cap=None):
        torch, ops = b.torch, b.ops
        self.b, self.L, self.B, self.T, self.top_k = b, L, B, T, top_k
        self.S = B * T
        self.valid = int(x_len.sum()) if x_len is not None else self.S
        self.gate_mode = ops.GATE_3M if top_k == 1 else ops.GATE_NAIVE
        self.layers = b.make_layers(L, top_k, router_bias)
        g = torch.Generator().manual_seed(20260003 + b.rank)
        self.act_dtype = torch.float32 if b.tf32 else torch.bfloat16
        S = self.S
        self.x_host = torch.randn(S, D, generator=g).bfloat16().to(self.act_dtype).pin_memory()
        self.e_host = (torch.randn(S, DEMB, generator=g).bfloat16().to(self.act_dtype).pin_memory()
                       if top_k == 1 else None)
        self.out_host = torch.empty(S, D, dtype=self.act_dtype).pin_memory()
        self.x_dev = self.x_host.to(b.dev)
        self.e_dev = None if self.e_host is None else self.e_host.to(b.dev)
        self.x_len = None if x_len is None else x_len.to(b.dev)
        self.bufs = [torch.empty_like(self.x_dev), torch.empty_like(self.x_dev)]
        self.ep_ctx = None
        if b.world > 1 and b.args.ep == "p2p":
            # symmetric buffers: one capacity for all ranks (their padded batches differ in length)
            t = torch.tensor([cap or max(S * top_k, 1)], device=b.dev, dtype=torch.int64)
            b.dist.all_reduce(t, op=b.dist.ReduceOp.MAX)
            self.ep_ctx = b.ep_mod.EpContext.from_process_group(b.E_local, D, cap=int(t.item()), timeout_ms=20000)
        if self.ep_ctx is not None:
            # outputs inside the symmetric buffer: the owners' epilogue writes the finished rows there (folded combine)
            self.bufs = [self.ep_ctx.out_buffer(0, S), self.ep_ctx.out_buffer(1, S)]
        self.graph = None
        self.final = None

    def layer_call(self, li, cur, e_in, out, **kw):
        b, ops, ly = self.b, self.b.ops, self.layers[li]
        common = dict(residual=cur, top_k=self.top_k, gate_mode=self.gate_mode, act_type=ops.ACT_SILU, ff_scale=0.5,
                      out=out, x_len=self.x_len, seq_len=self.T)
        if self.ep_ctx is not None:
            return self.ep_ctx.forward(cur, e_in, ly["Wr"], ly["br"], ly["experts"], Wr_packed=ly["Wrp"], **common,
                                       **ly["norms"], **kw)
        if b.world > 1:
            return b.ep_mod.ep_moe_layer(cur, e_in, ly["Wr"], ly["br"], ly["experts"], num_local_expert=b.E_local,
                                         group=None, top_k=self.top_k, gate_mode=self.gate_mode,
                                         act_type=ops.ACT_SILU, ff_scale=0.5, residual=cur, out=out, Wr_packed=ly["Wrp"])
        return ops.moe_layer(cur, e_in, ly["Wr"], ly["br"], ly["experts"], Wr_packed=None if b.tf32 else ly["Wrp"],
                             compute=ops.COMPUTE_TF32 if b.tf32 else ops.COMPUTE_BF16, **common, **ly["norms"], **kw)

    def step(self, x_in=None, e_in=None):
        cur = self.x_dev if x_in is None else x_in
        e_in = self.e_dev if e_in is None else e_in
        for li in range(self.L):
            out = self.bufs[li & 1]
            if self.ep_ctx is not None:   # the wait for the owners' rows is left to the next layer's call; the last waits
                self.layer_call(li, cur, e_in, out, wait=li == self.L - 1, out_slot=li & 1)
            else:
                self.layer_call(li, cur, e_in, out)
            cur = out
        return cur

    def warm_and_capture(self, warmup):
        b, torch, ops = self.b, self.b.torch, self.b.ops
        for _ in range(max(warmup, 3)):
            self.final = self.step()
        torch.cuda.synchronize()
        n0 = ops.launch_count()
        self.step()
        torch.cuda.synchronize()
        self.launches_per_step = ops.launch_count() - n0
        # the peer-memory EP path has no host synchronisation either, so the whole step is captured on every rank
        self.use_graph = (b.world == 1 or self.ep_ctx is not None) and not b.args.no_graph
        if self.use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=b.stream):
                self.final = self.step()
            for _ in range(3):
                self.graph.replay()
            torch.cuda.synchronize()

    def run_once(self):
        if self.use_graph:
            self.graph.replay()
        else:
            self.step()

    def stage_profile(self, iters):
        """Per-stage CUDA-event timing of eager steps (roofline input): {stage: us per call}."""
        b, ops = self.b, self.b.ops
        if not (b.world == 1 or self.ep_ctx is not None):
            return {}
        ops.profile_enable(True)
        for _ in range(iters):
            self.step()
        b.torch.cuda.synchronize()
        ms, calls = ops.profile_read()
        ops.profile_enable(False)
        return {k: (ms[k] / calls[k] * 1e3 if calls.get(k) else None) for k in ms}

    def kernel_spans(self):
        """In-graph kernel durations from the kernels' own %globaltimer marks (b200moe_debug_timeline): a second graph of
        the same step is captured with the marks on and replayed; span = first CTA start -> last CTA end, averaged over
        the launches of one replay.  Unlike CUDA events around a launch it holds no launch gap, and unlike an eager run
        it sees the kernel next to its real neighbours (programmatic dependent launch).  Single GPU, graph mode only."""
        b, torch = self.b, self.b.torch
        if not self.use_graph or b.world != 1:
            return None
        lib = importlib.import_module(PKG + "._lib").load()
        n_slots = self.launches_per_step + 4
        tl = torch.zeros(n_slots, 148, 8, dtype=torch.int64, device=b.dev)
        lib.b200moe_debug_timeline(tl.data_ptr(), n_slots)
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=b.stream):
                self.step()
            kinds = [lib.b200moe_debug_timeline_kind(i) for i in range(n_slots)]
        finally:
            lib.b200moe_debug_timeline(None, 0)
        with torch.cuda.stream(b.stream):
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            tl.zero_()
            torch.cuda.synchronize()
            g.replay()
            torch.cuda.synchronize()
        t = tl.cpu().numpy().astype("float64")
        spans = {1: [], 2: []}
        for i, k in enumerate(kinds):
            if k in spans:
                a = t[i]
                used = a[:, 0] > 0
                if used.any() and (a[used, 5] > 0).all():
                    spans[k].append((a[used, 5].max() - a[used, 0].min()) / 1e3)
        del g
        out = {}
        for k, name in ((1, "route"), (2, "expert_ffn")):
            if len(spans[k]) > 2:
                v = sorted(spans[k][1:])   # (the first launch of a replay has no predecessor to overlap with)
                out[name] = {"mean_us": sum(v) / len(v), "median_us": v[len(v) // 2], "launches": len(v)}
        return out or None

    def nonempty_experts(self):
        """Mean number of LOCAL experts per layer that receive at least one row (their weights are what a layer call
        has to stream): from one untimed pass with the routing returned."""
        b, torch = self.b, self.b.torch
        if self.ep_ctx is not None or b.world > 1:
            return float(b.E_local)   # counted per source rank on the owner; assume all local experts are hit
        cur, tot = self.x_dev, 0
        for li in range(self.L):
            out = self.bufs[li & 1]
            res = self.layer_call(li, cur, self.e_dev, out, return_routing=True)
            tot += int((res.counts > 0).sum())
            cur = out
        torch.cuda.synchronize()
        return tot / self.L

    def close(self):
        if self.ep_ctx is not None:
            self.b.dist.barrier()
            self.ep_ctx.close()
            self.ep_ctx = None

    def check_status(self):
        if self.ep_ctx is not None:
            status = self.ep_ctx.status()
            if status != 0:   # a kernel gave up waiting for a peer: whatever was timed is not the workload
                print(f"rank {self.b.rank}: expert-parallel flag wait timed out (status {status}); no result",
                      file=sys.stderr, flush=True)
                sys.exit(3)
        st = self.b.ops.status()
        if st != 0:
            print(f"rank {self.b.rank}: device status word {st}; no result", file=sys.stderr, flush=True)
            sys.exit(3)


def ep_parity(b: Bench, case: Case):
    """Untimed, once per multi-GPU run: layer 0 through the expert-parallel path on this rank's tokens against the
    single-GPU all-experts layer on the same tokens (the other ranks' experts gathered over NCCL for the check).
    Routing (expert assignment, counts, scatter indices) must be identical, outputs within the BF16 bar."""
    torch, ops, dist = b.torch, b.ops, b.dist
    ly = case.layers[0]
    ex = ly["experts"]

    def gather(t):
        parts = [torch.empty_like(t) for _ in range(b.world)]
        dist.all_gather(parts, t.contiguous())
        return torch.cat(parts, 0)

    full = ops.PackedExperts(gather(ex.W1), gather(ex.b1), gather(ex.W2), gather(ex.b2))
    out_ep = case.bufs[0]   # (the timed configuration: output in the symmetric buffer, combine folded into the owners)
    common = dict(residual=case.x_dev, top_k=case.top_k, gate_mode=case.gate_mode, act_type=ops.ACT_SILU, ff_scale=0.5,
                  x_len=case.x_len, seq_len=case.T, return_routing=True)
    r_ep = case.ep_ctx.forward(case.x_dev, case.e_dev, ly["Wr"], ly["br"], ex, Wr_packed=ly["Wrp"], out_slot=0,
                               **common)
    torch.cuda.synchronize()
    dist.barrier()
    r_1 = ops.moe_layer(case.x_dev, case.e_dev, ly["Wr"], ly["br"], full, Wr_packed=ly["Wrp"], **common)
    torch.cuda.synchronize()
    _, idx, _score, counts, mapping = r_ep
    if case.top_k == 1:
        routing = bool(torch.equal(idx, r_1.idx) and torch.equal(counts, r_1.counts) and
                       torch.equal(mapping, r_1.mapping))
    else:   # the order of the k experts of a token is not part of the contract; the set, counts and rows are
        routing = bool(torch.equal(torch.sort(idx, 1).values, torch.sort(r_1.idx, 1).values) and
                       torch.equal(counts, r_1.counts))
    a, c = out_ep.float(), r_1.out.float()
    rel = float((a - c).norm() / c.norm().clamp_min(1e-30)) if a.numel() else 0.0
    finite = bool(torch.isfinite(a).all())
    t = torch.tensor([0.0 if routing else 1.0, rel, 0.0 if finite else 1.0], device=b.dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = {"routing_exact": t[0].item() == 0.0, "rel_l2": float(t[1].item()), "finite": t[2].item() == 0.0,
           "checked": "layer 0, every rank: EP output vs single-GPU all-experts layer on the same tokens", "bar": 1e-2}
    res["ok"] = bool(res["routing_exact"] and res["finite"] and res["rel_l2"] <= 1e-2)
    return res


def rooflines(b: Bench, case: Case, stage_us, ms_step, nonempty=None):
    """-> (roofline of the expert-FFN kernel, per-stage rooflines, whole-layer roofline)."""
    es = 4 if b.tf32 else 2                                                      # bytes per weight / activation element
    k = case.top_k
    rows = case.valid * k                                                        # rows through the experts (this rank)
    if b.world > 1:                                                              # EP: rows received ~ rows sent (uniform)
        rows = b.all_sum(rows) / b.world
    ne = b.E_local if nonempty is None else nonempty
    w_bytes = 2 * ne * D * H * es + ne * (H + D) * 4                             # W1 + W2 + fp32 biases of the hit experts
    demb = DEMB if k == 1 else 0
    roofline = None
    traffic = None   # DRAM bytes per launch of the same kernel from one `ncu --set full` capture (profiles/)
    for name in ("r02_ffn_traffic.json", "r01_ffn_traffic.json"):
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", name)))
            if b.args.workload == tr.get("workload") and b.world == 1 and not b.tf32:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                break
        except Exception:
            pass
    t_ffn_us = stage_us.get("expert_ffn")
    if t_ffn_us:
        t_ffn = t_ffn_us * 1e-6
        fused = k == 1 and b.world == 1
        # xbuf read + (fused top-1 epilogue = the combine: residual read, out write, pos + score) or ybuf write
        act_bytes = rows * (3 * es * D + 8) if fused else rows * 2 * es * D
        alg_bytes = w_bytes + act_bytes
        flops = rows * 4 * D * H
        if alg_bytes / (b.hbm_peak * 1e9) >= flops / (b.tf_peak * 1e12):
            ach = alg_bytes / t_ffn / 1e9
            roofline = {"kernel": "ffn_kernel", "bound": "hbm", "achieved": ach, "peak": b.hbm_peak, "unit": "GB/s",
                        "frac": ach / b.hbm_peak, "traffic": traffic, "peak_source": b.peak_kind,
                        "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": t_ffn_us,
                        "experts_with_rows": ne}
        else:
            ach = flops / t_ffn / 1e12
            roofline = {"kernel": "ffn_kernel", "bound": "tensor", "achieved": ach, "peak": b.tf_peak,
                        "unit": "TFLOP/s", "frac": ach / b.tf_peak, "traffic": traffic, "peak_source": b.peak_kind,
                        "algorithmic_flops_per_launch": flops, "us_per_launch": t_ffn_us,
                        "frac_of_burst_peak": None}
            try:
                burst = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
                roofline["frac_of_burst_peak"] = ach / burst
            except Exception:
                pass
    # HBM-bound stages, bytes per SURVEY section 8(d) (scaled by es / 2 for fp32 activations)
    S_pad, S_val = case.S, case.valid
    gate_b = S_pad * (es * (D + demb)) + S_val * 8 * k
    disp_b = S_val * k * 2 * (es * D + 4)
    comb_b = S_val * (es * D * k + es * D + 4 * k + es * D)
    stage_bytes = {"gate": gate_b, "dispatch": disp_b, "combine": comb_b}
    stages = {}
    route_fused = stage_us.get("gate") and not stage_us.get("dispatch")
    for name, us in stage_us.items():
        if not us or name == "expert_ffn":
            continue
        nbytes = stage_bytes[name] + (disp_b if name == "gate" and route_fused else 0)
        label = "route (gate + dispatch fused)" if name == "gate" and route_fused else name
        stages[label] = {"us": us, "algorithmic_bytes": nbytes, "achieved_GBps": nbytes / us / 1e3,
                         "frac_of_hbm_peak": nbytes / us / 1e3 / b.hbm_peak}
    # 9 236 B per token at bf16, k = 1: gate read 2 048 + [8 + 2 056 + 2 048 + 1 028] k + 2 048
    layer_bytes = w_bytes + S_pad * es * (D + demb) + S_val * ((8 + 2056 + 2048 + 1028) * k + 2048) * (es // 2)
    layer_flops = rows * 4 * D * H + S_pad * 2 * (D + demb) * E
    t_layer = ms_step * 1e-3 / case.L
    layer_roofline = {
        "bound": "hbm" if layer_bytes / (b.hbm_peak * 1e9) >= layer_flops / (b.tf_peak * 1e12) else "tensor",
        "hbm_GBps": layer_bytes / t_layer / 1e9, "hbm_frac": layer_bytes / t_layer / 1e9 / b.hbm_peak,
        "tensor_TFLOPs": layer_flops / t_layer / 1e12, "tensor_frac": layer_flops / t_layer / 1e12 / b.tf_peak,
        "algorithmic_bytes_per_layer": layer_bytes, "us_per_layer": t_layer * 1e6}
    return roofline, stages, layer_roofline


# ----------------------------------------------------------------------------------------------------------------------
def run_sweep(b: Bench):
    """BASELINE.json configs[4]: global tokens 1K-1M x top-1 / top-2, plus one Zipf-skewed run (router bias).  The
    global token count is split evenly over the ranks (strong scaling); one JSON line per point."""
    torch, args = b.torch, b.args
    wl = WORKLOADS["sweep"]
    points = [(s, k, False) for s in SWEEP_TOKENS if s <= args.sweep_max for k in (1, 2)]
    points.append((min(65536, args.sweep_max), 1, True))
    if args.sweep_points:
        points = []
        for item in args.sweep_points.split(","):
            f = item.split(":")
            points.append((int(f[0]), int(f[1]) if len(f) > 1 else 1, len(f) > 2 and f[2] == "zipf"))
    for s_glob, k, zipf in points:
        s_loc = s_glob // b.world
        br = None
        if zipf:   # expert e's logit is lowered by ln(e + 1): a Zipf-like load with expert 0 the hottest
            br = -torch.log(torch.arange(1, E + 1, device=b.dev, dtype=torch.float32))
        L = wl["layers"]
        case = Case(b, L, 1, s_loc, None, top_k=k, router_bias=br)
        case.warm_and_capture(3)
        iters = max(3, min(args.steps, int(0.25 / max(1e-6, (s_loc * 2.5e-9 + 3e-5) * L))))
        ms_step = b.timed(case.run_once, iters) / iters
        stage_us = case.stage_profile(min(iters, 20))
        case.check_status()
        roofline, stages, layer_roofline = rooflines(b, case, stage_us, ms_step)
        load = None
        if b.world == 1:
            res = case.layer_call(0, case.x_dev, case.e_dev, case.bufs[1], return_routing=True)
            c = res.counts.float()
            load = {"max_over_mean": float(c.max() / c.mean().clamp_min(1e-9)), "empty_experts": int((c == 0).sum())}
        parity = ep_parity(b, case) if case.ep_ctx is not None else None
        if b.rank == 0:
            line = {"metric": "moe_layer_tokens_per_sec", "value": L * s_loc * b.world / (ms_step * 1e-3),
                    "unit": "tokens/s", "n_gpus": b.world, "steps": iters, "warmup": 3, "ms_per_step": ms_step,
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
                    "data": "synthetic",
                    "config": workload_config(args, wl, s_loc * b.world, s_loc * b.world, 1, top_k=k,
                                              extra={"routing": "zipf (router_bias = -ln(e + 1))" if zipf
                                                     else "natural (near-uniform)"}),
                    "sweep_point": {"global_tokens": s_glob, "top_k": k, "zipf": zipf, "expert_load": load},
                    "mode": "cuda_graph" if case.use_graph else "eager", "us_per_layer": ms_step * 1e3 / L,
                    "stage_us_per_layer": stage_us, "roofline": roofline, "roofline_stages": stages,
                    "layer_roofline": layer_roofline, "parity": parity,
                    "gpu_launches": case.launches_per_step * iters}
            print(json.dumps(line), flush=True)
        case.close()
        bad = parity is not None and not parity["ok"]
        del case
        b.ops.clear_workspaces()
        torch.cuda.empty_cache()
        if bad:
            sys.exit(4)


def encoder_bench(b, wl, steps, warmup, moe_ms_step):
    """The whole encoder around the path (BASELINE metric "18L encoder utts/sec"): embed net (6 dense Conformer
    blocks), Conv2dSubsampling4, `layers` FmoeConformerLayers whose feed-forward block runs through b200moe_block_forward,
    after_norm + output linear.  Everything outside the fast_moe block is library code (torch: cuDNN / cuBLAS / SDPA).
    Synthetic 206-frame, 40-dim features, random-init weights of the repo configuration, bf16, CUDA graph."""
    torch, dev = b.torch, b.dev
    enc = importlib.import_module(PKG + ".encoder")
    L, B, frames = wl["layers"], wl["utts"], 206
    torch.manual_seed(7)
    model = enc.ConformerMoEEncoder(
        40, 5000, attention_heads=8, attention_dim=512, num_blocks=L,
        embed_conf=dict(attention_heads=4, attention_dim=512, linear_units=1024, num_blocks=6),
        moe_conf=dict(num_experts=32, hidden_units=1024, rand_init_router=True)).to(dev)
    model.to_inference(torch.bfloat16)
    feats = (torch.randn(B, frames, 40, device=dev) * 0.5).bfloat16()
    lens = torch.full((B,), frames, dtype=torch.int64, device=dev)
    feats_host = feats.cpu().pin_memory()
    ops = b.ops
    stream = torch.cuda.Stream(device=dev)
    with torch.no_grad(), torch.cuda.stream(stream):
        for _ in range(2):
            out = model(feats, lens)
        stream.synchronize()
        n0 = ops.launch_count()
        out = model(feats, lens)
        launches = ops.launch_count() - n0
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            out = model(feats, lens)
        out_host = torch.empty(out.shape, dtype=out.dtype).pin_memory()

        def timed(fn, n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            stream.synchronize()
            e0.record(stream)
            for _ in range(n):
                fn()
            e1.record(stream)
            e1.synchronize()
            return e0.elapsed_time(e1) / n

        def e2e():
            feats.copy_(feats_host, non_blocking=True)
            graph.replay()
            out_host.copy_(out, non_blocking=True)

        for _ in range(warmup):
            graph.replay()
        ms = timed(graph.replay, steps)
        for _ in range(2):
            e2e()
        ms_e2e = timed(e2e, steps)
    finite = bool(torch.isfinite(out.float()).all())
    del graph, model
    torch.cuda.empty_cache()
    return {"utterances_per_sec": B / (ms * 1e-3), "ms_per_step": ms, "layers": L, "utterances": B, "frames": frames,
            "tokens_per_utterance": 50, "moe_path_share": moe_ms_step / ms, "own_kernel_launches_per_step": int(launches),
            "e2e": {"utterances_per_sec": B / (ms_e2e * 1e-3), "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": feats_host.numel() * 2, "d2h_bytes_per_step": out_host.numel() * 2},
            "output_finite": finite, "mode": "cuda_graph", "dtype": "bf16",
            "note": "fast_moe blocks: this repository's kernels (C ABI); embed net, attention, convolution, subsampling, "
                    "output linear: torch library ops"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    b = Bench(args)
    torch, ops = b.torch, b.ops
    if args.workload == "sweep":
        run_sweep(b)
        if b.world > 1:
            b.dist.destroy_process_group()
        return

    wl = WORKLOADS[args.workload]
    L = wl["layers"]
    B, T, x_len, valid = rank_geometry(torch, wl, b.rank)
    case = Case(b, L, B, T, x_len)
    S = case.S
    x_dev, e_dev, stream, dev = case.x_dev, case.e_dev, b.stream, b.dev
    case.warm_and_capture(args.warmup)
    use_graph = case.use_graph

    parity = None
    if case.ep_ctx is not None:
        parity = ep_parity(b, case)
        if not parity["ok"]:
            if b.rank == 0:
                print(json.dumps({"error": "expert-parallel parity check failed", "parity": parity}), flush=True)
            sys.exit(4)

    sampler = ClockSampler(b.local_rank)
    sampler.start()

    # ---- value: device-resident inputs, K steps
    K = args.steps
    ms_total = b.timed(case.run_once, K)
    ms_step = ms_total / K
    tokens_global = b.all_sum(valid)
    padded_global = b.all_sum(S)
    tokens_per_step = L * tokens_global
    value = tokens_per_step / (ms_step * 1e-3)

    # ---- eager device time (no graph) and per-stage CUDA-event timing of the same steps (roofline input)
    ms_eager = b.timed(lambda: case.step(), K) / K
    stage_us = case.stage_profile(min(K, 400))

    # ---- e2e: host buffers in, host buffer out, EVERY step, through the public API.  The copies run on their own
    # streams so that step i+1's upload and step i's download overlap step i's / i+1's compute, the way a serving loop
    # would feed the layer; every step still uploads its own inputs from pinned memory and downloads its own result.
    x_host, e_host, out_host = case.x_host, case.e_host, case.out_host
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
        s = e2e_i[0] & 1
        e2e_i[0] += 1
        with torch.cuda.stream(h2d_stream):
            h2d_stream.wait_event(ev_used[s])
            st_x[s].copy_(x_host, non_blocking=True)
            st_e[s].copy_(e_host, non_blocking=True)
            ev_up[s].record(h2d_stream)
        stream.wait_event(ev_up[s])
        if use_graph:   # the graph reads x_dev / e_dev and leaves the result in `final`
            x_dev.copy_(st_x[s], non_blocking=True)
            e_dev.copy_(st_e[s], non_blocking=True)
            ev_used[s].record(stream)
            case.graph.replay()
            res = case.final
        else:
            res = case.step(st_x[s], st_e[s])
            ev_used[s].record(stream)
        stream.wait_event(ev_down[s])
        st_o[s].copy_(res, non_blocking=True)
        ev_out[s].record(stream)
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(ev_out[s])
            out_hosts[s].copy_(st_o[s], non_blocking=True)
            ev_down[s].record(d2h_stream)

    def e2e_finish():   # the last downloads belong to the timed region
        stream.wait_stream(d2h_stream)

    for _ in range(3):
        e2e_step()
    e2e_finish()
    ms_e2e = b.timed(e2e_step, K, e2e_finish) / K
    e2e_value = tokens_per_step / (ms_e2e * 1e-3)
    clocks = sampler.stop()

    # ---- sustained load: the same graph replayed back to back for --sustain seconds, with its own clock record
    sustained = None
    if args.sustain > 0:
        x_dev.copy_(x_host, non_blocking=True)
        s2 = ClockSampler(b.local_rank)
        s2.start()
        chunk = max(1, int(0.25 / max(ms_step * 1e-3, 1e-6)))
        n_done, t_ms, t0 = 0, 0.0, time.perf_counter()
        while time.perf_counter() - t0 < args.sustain:
            t_ms += b.timed(case.run_once, chunk)
            n_done += chunk
        sustained = {"seconds": time.perf_counter() - t0, "steps": n_done, "ms_per_step": t_ms / n_done,
                     "value": tokens_per_step / (t_ms / n_done * 1e-3), "unit": "tokens/s", "clocks": s2.stop()}
    case.check_status()

    nonempty = case.nonempty_experts()
    spans = None if os.environ.get("B200MOE_AB_OLD_LIB") else case.kernel_spans()
    roofline, stages, layer_roofline = rooflines(b, case, stage_us, ms_step, nonempty)
    if roofline is not None and spans and "expert_ffn" in spans:
        # the same algorithmic bytes / flops over the kernel's own in-graph duration (see Case.kernel_spans)
        us = spans["expert_ffn"]["mean_us"]
        work = roofline.get("algorithmic_bytes_per_launch") or roofline.get("algorithmic_flops_per_launch")
        scale = 1e3 if roofline["bound"] == "hbm" else 1e6
        roofline["in_graph_span_us"] = us
        roofline["achieved_in_graph"] = work / us / scale
        roofline["frac_in_graph"] = work / us / scale / roofline["peak"]

    # ---- CPU baseline: the oracle on this box's host cores, bounded sample (rank 0, N = 1 only)
    cpu_baseline = None
    if b.rank == 0 and b.world == 1 and not args.no_cpu_baseline:
        oracle = importlib.import_module("oracle.moe_oracle")   # bench's cpu_baseline leg: allowed oracle use
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n_sets = min(L, 2)
        cpu_layers = [dict(Wr=ly["Wr"].cpu(), W1=ly["experts"].W1.float().cpu(), b1=ly["experts"].b1.cpu(),
                           W2=ly["experts"].W2.float().cpu(), b2=ly["experts"].b2.cpu())
                      for ly in case.layers[:n_sets]]
        seq = [cpu_layers[i % n_sets] for i in range(L)]
        xc, ec = x_host.float(), e_host.float()
        xl = None if x_len is None else x_len
        # bounded: a prefix of the layer stack when one full step would take more than a few seconds
        with torch.no_grad():
            t0 = time.perf_counter()
            cpu_forward_step(oracle, torch, seq[:1], xc, ec, xl, T)
            t_layer = time.perf_counter() - t0
            n_layers = max(1, min(L, int(4.0 / max(t_layer, 1e-6))))
            seq = seq[:n_layers]
            cpu_forward_step(oracle, torch, seq, xc, ec, xl, T)
            reps, t0 = 0, time.perf_counter()
            while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 50):
                cpu_forward_step(oracle, torch, seq, xc, ec, xl, T)
                reps += 1
            dt = (time.perf_counter() - t0) / reps
        cpu_baseline = {"value": n_layers * valid / dt, "unit": "tokens/s", "cores": cores, "kind": "port",
                        "sample": f"{reps} passes of {n_layers} of the {L} layers x {valid} tokens of the CPU oracle, "
                                  f"{n_sets} weight sets cycled, torch {torch.__version__} fp32"}

    encoder = None
    if args.encoder and b.world == 1 and "tok_per_utt" in wl and wl["tok_per_utt"] == 50 and not b.tf32:
        encoder = encoder_bench(b, wl, min(K, 30), max(3, min(args.warmup, 5)), ms_step)

    if b.rank == 0:
        line = {
            "metric": "moe_layer_tokens_per_sec", "value": value, "unit": "tokens/s", "n_gpus": b.world, "steps": K,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32" if b.tf32 else "bf16", "data": "synthetic",
            "config": workload_config(args, wl, int(tokens_global), int(padded_global), wl["utts"] * b.world),
            "utterances_per_sec": wl["utts"] * b.world / (ms_step * 1e-3),
            "mode": "cuda_graph" if use_graph else "eager",
            "ms_per_step_eager": ms_eager,
            "us_per_layer": ms_step * 1e3 / L,
            "stage_us_per_layer": stage_us,
            "roofline": roofline,
            "roofline_stages": stages,
            "kernel_spans_in_graph": spans,
            "layer_roofline": layer_roofline,
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "tokens/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": (x_host.numel() + e_host.numel()) * x_host.element_size(),
                    "d2h_bytes_per_step": out_host.numel() * out_host.element_size()},
            "parity": parity,
            "sustained": sustained,
            "encoder": encoder,
            "host_numa_node": b.numa,
            "gpu_launches": case.launches_per_step * K,
            "gpu_launches_per_step": case.launches_per_step,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    case.close()
    if b.world > 1:
        b.dist.destroy_process_group()


if __name__ == "__main__":
    main()
