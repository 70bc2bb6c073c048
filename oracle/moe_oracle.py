"""CPU oracle for the fast_moe expert layer -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product path (the `3m-asr-inference_b200` package) never does and fails loudly without its CUDA library.

It restates, in plain PyTorch CPU ops (fp32, with fp64 for the router logits), what the reference computes on this
path.  Citations are relative to the upstream tree (/root/reference):

  gate_3m        trainer_3m_fix/model/dfsmn_base_fmoe_localComm_catEmbed.py:166-181,210-211 (live torch code) ==
                 trainer_3m_fix/layer/positionwise_feed_forward.py:151-157 (commented twin), TRT form :169-207,225 and
                 TRTAPI++/plugin/softmax_topk_plugin/softmax_topk_kernel.cu:26-89 (value = 1 / sum(exp(l - max)))
  gate_naive     trainer_3m_fix/fmoe/gates.py:51-66 (NaiveGate.forward)
  prepare        trainer_3m_fix/fmoe/functions.py:29-35 (sort + counts) and
                 TRTAPI++/plugin/fmoe_expert_plugin/fmoe_expert_kernel.cu:25-73 (histogram, scan, mapping)
  scatter/gather trainer_3m_fix/fmoe/functions.py:72,194; fmoe_expert_kernel.cu:103,202
  expert_ffn     TRTAPI++/plugin/fmoe_expert_plugin/fmoe_expert_plugin.cpp:82-128; fmoe/functions.py:142-148;
                 fmoe/transformer.py:22-30; activation trainer_3m_fix/utils/common.py:24-28
  combine        positionwise_feed_forward.py:257-258 (x gate_value), fmoe/layers.py:199-206 (top-k bmm),
                 layer/fmoe_transformer.py:155-158 and layer/fmoeExMarc_transformer.py:150-154 (x ff_scale + residual)
  layer_norm     nn.LayerNorm(size, eps=1e-12) of trainer_3m_fix/layer/fmoe_transformer.py:54-65 (norm_ff, norm_final);
                 the TensorRT build swaps in TRTAPI++/plugin/layer_norm_plugin/layer_norm_kernel.cu, same formula in
                 fp32 with eps dropped -- indistinguishable at eps = 1e-12
  moe_block_forward  layer/fmoe_transformer.py:144-166: residual = x; x = norm_ff(x); x = feed_forward(x, embed);
                 x = residual + ff_scale * x; x = norm_final(x)

Pinning.  The reference ships no tests, golden vectors or fixtures for this path, and its arithmetic lives in the
un-vendored, un-pinned third-party extension `fmoe_cuda` (laekov/fastmoe, Tencent-modified; call sites
fmoe/functions.py:27,38,72,74,94,103,115,128,145,158,185,194,205,207).  What CAN run here is the reference's own Python
orchestration (NaiveGate, moe_prepare_forward, MOEScatter, MOEbiasLinear, MOEGather, the 3M router gate) once
`fmoe_cuda`'s four primitives are stubbed with their published fastmoe semantics; tests/golden/make_golden.py does
exactly that and the resulting vectors under tests/golden/ pin this oracle (tests/test_oracle_golden.py).  The two
places where the reference's result is not a function of its inputs are pinned by definition instead: order within an
expert (torch.sort is unstable / the plugin ranks with atomicAdd) is defined as stable by entry index, and exact ties
in the router go to the lowest expert index.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

ACT_SILU, ACT_RELU, ACT_GELU = 0, 1, 2
GATE_3M, GATE_NAIVE = 0, 1


def activation(h: torch.Tensor, act_type: int) -> torch.Tensor:
    if act_type == ACT_SILU:  # Swish: x * sigmoid(x)   (utils/common.py:24-28)
        return h * torch.sigmoid(h)
    if act_type == ACT_RELU:
        return torch.relu(h)
    if act_type == ACT_GELU:  # torch.nn.GELU() default = exact erf form (fmoe/transformer.py:45)
        return 0.5 * h * (1.0 + torch.erf(h / math.sqrt(2.0)))
    raise ValueError(f"bad act_type {act_type}")


def router_logits(x: torch.Tensor, embed: Optional[torch.Tensor], Wr: torch.Tensor,
                  br: Optional[torch.Tensor]) -> torch.Tensor:
    """cat([embed, x]) @ Wr (+ br), accumulated in fp64 so that the arg-max does not depend on summation order."""
    r = x if embed is None else torch.cat([embed, x], dim=-1)  # embed FIRST (positionwise_feed_forward.py:225)
    logits = r.double() @ Wr.double()
    if br is not None:
        logits = logits + br.double()
    return logits


def gate_3m(x, embed, Wr, br=None):
    """softmax over all experts, then max: returns idx [S] int64, value [S] float32, logits [S, E] float64."""
    logits = router_logits(x, embed, Wr, br)
    probs = torch.softmax(logits, dim=-1)
    value, idx = probs.max(dim=-1)  # CPU max returns the first (lowest) index among exact ties
    return idx, value.float(), logits


def gate_naive(x, W, b, top_k: int):
    """NaiveGate: top-k of Linear(x), softmax over the k selected logits.  W [d, E] (= gate.weight.T), b [E].
    Order within the k is canonicalised to descending logit, ties to the lower index (sorted=False leaves it open)."""
    logits = router_logits(x, None, W, b)
    # stable descending sort == "largest first, lowest index among equals"
    order = torch.sort(logits, dim=-1, descending=True, stable=True).indices[:, :top_k]
    vals = torch.gather(logits, 1, order)
    score = torch.softmax(vals, dim=-1)
    return order, score.float(), logits


def prepare(idx_flat: torch.Tensor, num_expert: int) -> Dict[str, torch.Tensor]:
    """Counting sort of entries by expert, stable in the entry index.  Entries with idx < 0 are dropped (mapping -1).
    counts [E], offsets [E+1] (= the plugin's acc_histogram), pos [n_valid] (fastmoe's `pos`), mapping [n]."""
    idx_flat = idx_flat.reshape(-1).long()
    n = idx_flat.numel()
    valid = (idx_flat >= 0) & (idx_flat < num_expert)
    key = torch.where(valid, idx_flat, torch.full_like(idx_flat, num_expert))
    pos_all = torch.sort(key, stable=True).indices
    n_valid = int(valid.sum())
    pos = pos_all[:n_valid]
    counts = torch.bincount(idx_flat[valid], minlength=num_expert)[:num_expert]
    offsets = torch.zeros(num_expert + 1, dtype=torch.long)
    offsets[1:] = torch.cumsum(counts, 0)
    mapping = torch.full((n,), -1, dtype=torch.long)
    mapping[pos] = torch.arange(n_valid)
    return {"counts": counts, "offsets": offsets, "pos": pos, "mapping": mapping}


def expert_ffn(xbuf: torch.Tensor, counts: torch.Tensor, W1, b1, W2, b2, act_type: int = ACT_SILU) -> torch.Tensor:
    """ybuf[rows of e] = act(xbuf_e @ W1[e].T + b1[e]) @ W2[e].T + b2[e];  W1 [E,H,D], W2 [E,D,H] (FMoELinear layout)."""
    out = torch.empty(xbuf.shape[0], W2.shape[1], dtype=xbuf.dtype)
    base = 0
    for e, c in enumerate(counts.tolist()):
        if c == 0:
            continue
        xe = xbuf[base:base + c]
        h = xe @ W1[e].t()
        if b1 is not None:
            h = h + b1[e]
        h = activation(h, act_type)
        y = h @ W2[e].t()
        if b2 is not None:
            y = y + b2[e]
        out[base:base + c] = y
        base += c
    return out


def moe_forward(x, embed, Wr, br, W1, b1, W2, b2, *, top_k=1, gate_mode=GATE_3M, act_type=ACT_SILU, residual=None,
                ff_scale=1.0, keep_expert_output=False, x_len=None, T=None,
                bf16_intermediate=False) -> Dict[str, torch.Tensor]:
    """The whole layer on [S, D] activations.  x_len/T: rows t >= x_len[b] of each length-T sequence are padding:
    they are not routed and their output is the residual (or zero).
    bf16_intermediate=True rounds the scattered input and the hidden activations to bf16 like the GPU path does
    (used only to separate rounding from logic errors when debugging; parity is judged against the fp32 form)."""
    S, D = x.shape
    E = W1.shape[0]
    x = x.float()
    if gate_mode == GATE_3M:
        assert top_k == 1
        idx, value, logits = gate_3m(x, None if embed is None else embed.float(), Wr, br)
        idx = idx.view(S, 1)
        score = value.view(S, 1)
    else:
        idx, score, logits = gate_naive(x, Wr, br, top_k)
    if x_len is not None:
        t_in_seq = torch.arange(S) % T
        valid = t_in_seq < x_len.long().repeat_interleave(T)
        idx = torch.where(valid[:, None], idx, torch.full_like(idx, -1))
        score = torch.where(valid[:, None], score, torch.zeros_like(score))
    prep = prepare(idx.reshape(-1), E)
    pos = prep["pos"]
    xbuf = x[pos // top_k]  # repeat_interleave(top_k) then local_scatter (fmoe/layers.py:199, functions.py:72)
    if bf16_intermediate:
        xbuf = xbuf.bfloat16().float()
    if bf16_intermediate:
        ybuf = _expert_ffn_bf16(xbuf, prep["counts"], W1, b1, W2, b2, act_type)
    else:
        ybuf = expert_ffn(xbuf, prep["counts"], W1.float(), b1, W2.float(), b2, act_type)
    # local_gather + weighted sum over the k slots
    y_entries = torch.zeros(S * top_k, D)
    y_entries[pos] = ybuf
    y_entries = y_entries.view(S, top_k, D)
    if keep_expert_output:
        w = (idx >= 0).float()
    else:
        w = score * (idx >= 0).float()
    moe = torch.einsum("sk,skd->sd", w, y_entries)
    out = ff_scale * moe
    if residual is not None:
        out = residual.float() + out
    return {"idx": idx, "score": score, "logits": logits, "counts": prep["counts"], "offsets": prep["offsets"],
            "pos": pos, "mapping": prep["mapping"].view(S, top_k), "xbuf": xbuf, "ybuf": ybuf, "moe": moe, "out": out}


def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """torch.nn.LayerNorm over the last dimension, written out: biased variance, eps inside the root, fp32."""
    x = x.float()
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * gamma.float() + beta.float()


def moe_block_forward(x, embed, Wr, br, W1, b1, W2, b2, *, norm_ff=None, norm_final=None, eps=1e-12, ff_scale=0.5,
                      round_norm_to=None, **layer_kw) -> Dict[str, torch.Tensor]:
    """The feed-forward part of one Conformer block (fmoe_transformer.py:144-166):
    out = norm_final(x + ff_scale * MoE(norm_ff(x), embed)).  norm_* = (gamma, beta) or None.
    round_norm_to: dtype the normalised input is rounded to before it enters the layer -- the reference hands the
    LayerNorm plugin's output to the next graph layer in the engine's activation dtype."""
    xn = x.float() if norm_ff is None else layer_norm(x, norm_ff[0], norm_ff[1], eps)
    if round_norm_to is not None:
        xn = xn.to(round_norm_to).float()
    r = moe_forward(xn, embed, Wr, br, W1, b1, W2, b2, residual=x.float(), ff_scale=ff_scale, **layer_kw)
    r["xn"] = xn
    r["pre_norm_out"] = r["out"]
    if norm_final is not None:
        r["out"] = layer_norm(r["out"], norm_final[0], norm_final[1], eps)
    return r


def _expert_ffn_bf16(xbuf, counts, W1, b1, W2, b2, act_type):
    out = torch.empty(xbuf.shape[0], W2.shape[1])
    base = 0
    for e, c in enumerate(counts.tolist()):
        if c == 0:
            continue
        xe = xbuf[base:base + c]
        h = xe @ W1[e].float().t()
        if b1 is not None:
            h = h + b1[e]
        h = activation(h, act_type).bfloat16().float()
        y = h @ W2[e].float().t()
        if b2 is not None:
            y = y + b2[e]
        out[base:base + c] = y
        base += c
    return out


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.double()
    b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
