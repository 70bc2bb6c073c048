"""Seeded synthetic inputs and random-init weights for the fast_moe layer.

Reference initialisation being mirrored (paths relative to the upstream tree):
  expert weights  xavier_uniform_(gain=0.5), biases 0      trainer_3m_fix/fmoe/layers.py:34-38
  router weights  xavier_uniform_(gain=0.5) when rand_init_router, else zeros
                                                          trainer_3m_fix/layer/positionwise_feed_forward.py:134-144,
                                                          trainer_3m_fix/model/dfsmn_base_fmoe_localComm_catEmbed.py:147-148
  the reference's only synthetic-input precedent is np.random.rand(B, S, D) (data/generate_trtexec_inputs.py).

Everything is drawn on the bf16 grid (products of two bf16 values are exact in fp32) and tokens whose top-k / top-(k+1)
router margin is below `margin` are re-drawn deterministically, so that "which expert wins" is a well-posed question
for an fp32 accumulation order different from the oracle's fp64 one.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch


def _bf16_grid(t: torch.Tensor) -> torch.Tensor:
    return t.bfloat16().float()


def xavier_uniform(shape, gain: float, gen: torch.Generator) -> torch.Tensor:
    """torch.nn.init.xavier_uniform_ semantics for a [*, out, in] tensor (fan computed over the last two dims the way
    torch does for >2-D tensors: receptive field = product of dims beyond the first two)."""
    t = torch.empty(shape)
    if t.dim() < 2:
        raise ValueError("xavier needs >= 2 dims")
    rf = 1
    for s in shape[2:]:
        rf *= s
    fan_in = shape[1] * rf
    fan_out = shape[0] * rf
    bound = gain * math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen) * 2.0 - 1.0) * bound


@dataclass
class LayerWeights:
    Wr: torch.Tensor              # [Demb + D, E] fp32 (router_weights)
    br: Optional[torch.Tensor]    # [E] or None
    W1: torch.Tensor              # [E, H, D]
    b1: torch.Tensor              # [E, H]
    W2: torch.Tensor              # [E, D, H]
    b2: torch.Tensor              # [E, D]


def make_weights(seed: int, E: int, D: int, H: int, Demb: int, *, router_bias: bool = False,
                 random_bias: bool = False, router_gain: float = 0.5) -> LayerWeights:
    g = torch.Generator().manual_seed(seed)
    Wr = _bf16_grid(xavier_uniform((Demb + D, E), router_gain, g))
    br = _bf16_grid((torch.rand(E, generator=g) - 0.5) * 0.2) if router_bias else None
    W1 = _bf16_grid(xavier_uniform((E, H, D), 0.5, g))
    W2 = _bf16_grid(xavier_uniform((E, D, H), 0.5, g))
    if random_bias:
        b1 = _bf16_grid((torch.rand(E, H, generator=g) - 0.5) * 0.2)
        b2 = _bf16_grid((torch.rand(E, D, generator=g) - 0.5) * 0.2)
    else:
        b1 = torch.zeros(E, H)
        b2 = torch.zeros(E, D)
    return LayerWeights(Wr, br, W1, b1, W2, b2)


def make_activations(seed: int, S: int, D: int, Demb: int, w: LayerWeights, *, top_k: int = 1,
                     margin: float = 1e-4, max_redraw: int = 64):
    """x [S, D], embed [S, Demb] (None when Demb == 0) ~ N(0, 1) on the bf16 grid, with near-tie tokens re-drawn."""
    g = torch.Generator().manual_seed(seed)
    x = _bf16_grid(torch.randn(S, D, generator=g))
    embed = _bf16_grid(torch.randn(S, Demb, generator=g)) if Demb > 0 else None
    E = w.Wr.shape[1]
    if S == 0 or E <= top_k:
        return x, embed
    for _ in range(max_redraw):
        r = x if embed is None else torch.cat([embed, x], dim=-1)
        logits = r.double() @ w.Wr.double()
        if w.br is not None:
            logits = logits + w.br.double()
        top = torch.topk(logits, top_k + 1, dim=-1).values
        gaps = top[:, :-1] - top[:, 1:]          # every adjacent gap among the top (k+1) must be clear
        bad = (gaps.min(dim=-1).values < margin).nonzero().flatten()
        if bad.numel() == 0:
            break
        x[bad] = _bf16_grid(torch.randn(bad.numel(), D, generator=g))
        if embed is not None:
            embed[bad] = _bf16_grid(torch.randn(bad.numel(), Demb, generator=g))
    return x, embed
