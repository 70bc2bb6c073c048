"""Mirror of trainer_3m_fix/layer/positionwise_feed_forward.py: Expert (:94-112) and LocalFmoeCatEmbedFeedForward
(:115-265), the module that replaces a Conformer block's feed_forward with gated FFN experts.

Same constructor arguments, same parameter names and shapes (`experts.w_1.{weight,bias}`, `experts.w_2.{weight,bias}`,
`router_weights [idim + embed_dim, num_experts * world_size]`, optional `router_bias`), so a 3M-ASR state_dict loads
unchanged.  The reference's forward takes a TensorRT `network_helper` and emits graph layers; here forward takes the
tensors and runs the layer: concat + router MatMul + softmax/top-1 + dispatch + 32-expert FFN + gather + x gate_value,
optionally with the x ff_scale and + residual that the Conformer block applies next (layer/fmoe_transformer.py:145-158).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .fmoe.layers import FMoELinear, PackedExpertCache, activation_code


class Swish(torch.nn.Module):
    """trainer_3m_fix/utils/common.py:24-28"""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x * torch.sigmoid(x)


class Expert(torch.nn.Module):
    """Parameter holder for the two expert linears (positionwise_feed_forward.py:94-112)."""

    def __init__(self, num_experts, idim, hidden_units, dropout_rate, activation=torch.nn.ReLU(), rank=0):
        super().__init__()
        self.w_1 = FMoELinear(num_experts, idim, hidden_units, bias=True, rank=rank)
        self.activation = activation
        self.dropout = torch.nn.Dropout(dropout_rate)
        self.w_2 = FMoELinear(num_experts, hidden_units, idim, bias=True, rank=rank)


class LocalFmoeCatEmbedFeedForward(torch.nn.Module):
    def __init__(self, idim, embed_dim, num_experts=4, rank=0, world_size=1, hidden_units=1024, dropout_rate=0.0,
                 activation=torch.nn.ReLU(), capacity_factor=-1.0, router_regularization="l1_plus_importance",
                 router_with_bias=False, keep_expert_output=False, rand_init_router=False, comm=None):
        super().__init__()
        self.rank = rank
        self.world_size = world_size
        self.comm = comm
        self.num_experts = num_experts
        self.capacity_factor = capacity_factor
        self.router_regularization = router_regularization
        self.idim = idim
        self.embed_dim = embed_dim
        self.hidden_units = hidden_units
        self.experts = Expert(num_experts, idim, hidden_units, dropout_rate, activation=activation, rank=rank)
        router_input_dim = idim + embed_dim
        self.router_weights = torch.nn.Parameter(torch.zeros(router_input_dim, num_experts * world_size))
        if router_with_bias:
            self.router_bias = torch.nn.Parameter(torch.zeros(num_experts * world_size))
        else:
            self.router_bias = None
        if rand_init_router:
            torch.nn.init.xavier_uniform_(self.router_weights, gain=0.5)
        self.keep_expert_output = keep_expert_output
        self._cache = PackedExpertCache()
        if capacity_factor is not None and capacity_factor > 0:
            raise NotImplementedError("token dropping by capacity is a training-time feature (cf = -1 at inference)")

    def invalidate_packed(self) -> None:
        """Drop the packed copies of experts and router (weights changed through `.data`: no version bump to key on)."""
        self._cache.invalidate()
        self._wrp_key = None
        self._wrln_key = None

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.invalidate_packed()

    ep_capacity = 8192   # tokens per rank per call the expert-parallel receive buffers are sized for

    def _ep_context(self, n_tokens: int):
        from . import ep_p2p
        ctx = getattr(self, "_ep_ctx", None)
        if ctx is None:
            ctx = ep_p2p.EpContext.from_process_group(self.num_experts, self.idim, max(self.ep_capacity, n_tokens),
                                                      group=self.comm)
            self._ep_ctx = ctx
        if n_tokens > ctx.cap:
            raise RuntimeError(f"{n_tokens} tokens exceed ep_capacity={ctx.cap}; set a larger `ep_capacity` on every "
                               f"rank before the first forward")
        return ctx

    def _router(self):
        w = self.router_weights
        wr = w.detach()
        if wr.dtype != torch.float32:
            wr = wr.float()
        br = None if self.router_bias is None else self.router_bias.detach().float().contiguous()
        return wr.contiguous(), br

    def _router_packed(self, wr):
        """bf16 hi/lo packing of the router for the tensor-core gate, refreshed when the parameter changes."""
        if wr.shape[1] > 32 or not wr.is_cuda:
            return None
        w = self.router_weights
        key = (w.data_ptr(), w._version, str(w.device))
        if getattr(self, "_wrp_key", None) is None or self._wrp_key != key:
            self._wrp = ops.pack_router(wr)
            self._wrp_key = key
        return self._wrp

    def _router_packed_ln(self, wr, norm_ff):
        """Pre-scaled router for the folded norm_ff, refreshed when the router or the norm's parameters change."""
        if norm_ff is None or wr.shape[1] > 32 or not wr.is_cuda:
            return None
        w = self.router_weights
        key = (w.data_ptr(), w._version, norm_ff.weight.data_ptr(), norm_ff.weight._version, norm_ff.bias.data_ptr(),
               norm_ff.bias._version, str(w.device))
        if getattr(self, "_wrln_key", None) is None or self._wrln_key != key:
            self._wrln = ops.pack_router_ln(wr, norm_ff.weight.detach().float().contiguous(),
                                            norm_ff.bias.detach().float().contiguous())
            self._wrln_key = key
        return self._wrln

    def forward(self, inputs: torch.Tensor, embed: Optional[torch.Tensor], mask: Optional[torch.Tensor] = None, *,
                residual: Optional[torch.Tensor] = None, ff_scale: float = 1.0, return_routing: bool = False,
                norm_ff: Optional[torch.nn.LayerNorm] = None, norm_final: Optional[torch.nn.LayerNorm] = None):
        """inputs [B, T, idim]; embed [B, T, embed_dim]; mask [B] int32 valid lengths (the plugin's `mask` input,
        softmax_topk_plugin.cpp:88-127).  Returns the gate-weighted expert output [B, T, idim]; with `residual` /
        `ff_scale` the Conformer block's  residual + ff_scale * out  is fused in as well, and with `norm_ff` /
        `norm_final` (the block's nn.LayerNorm modules, fmoe_transformer.py:54-65) the LayerNorms either side of it:
        norm_final(residual + ff_scale * MoE(norm_ff(inputs), embed))  -- see `feed_forward_block`."""
        assert inputs.dim() == 3
        B, T, D = inputs.shape
        norms = {}
        for name, ln in (("norm_ff", norm_ff), ("norm_final", norm_final)):
            if ln is not None:
                if tuple(ln.normalized_shape) != (D,) or ln.weight is None or ln.bias is None:
                    raise ValueError(f"{name} must be an affine LayerNorm over the {D} features")
                norms[name] = (ln.weight.detach().float().contiguous(), ln.bias.detach().float().contiguous())
        if norms:
            eps = {float(ln.eps) for ln in (norm_ff, norm_final) if ln is not None}
            if len(eps) != 1:
                raise ValueError("norm_ff and norm_final must share eps (the reference uses 1e-12 for both)")
            norms["eps"] = eps.pop()
            if norm_ff is not None and inputs.dtype == torch.bfloat16:
                norms["Wr_packed_ln"] = self._router_packed_ln(self._router()[0], norm_ff)
        x = inputs.contiguous()
        e = None if embed is None else embed.contiguous()
        Wr, br = self._router()
        packed = self._cache.get(self.experts.w_1, self.experts.w_2)
        act = activation_code(self.experts.activation)
        x_len = None if mask is None else mask.reshape(-1).to(torch.int32).contiguous()
        if self.world_size > 1 and x.dtype == torch.bfloat16:
            # expert parallelism over peer-mapped memory (the product path); the context is created collectively on
            # the first call, `ep_capacity` (tokens per rank per call) sizes its receive buffers
            ctx = self._ep_context(B * T)
            out = ctx.forward(x.view(B * T, D), None if e is None else e.view(B * T, -1), Wr, br, packed,
                              residual=None if residual is None else residual.contiguous().view(B * T, D), x_len=x_len,
                              seq_len=T, top_k=1, gate_mode=ops.GATE_3M, act_type=act, ff_scale=ff_scale,
                              keep_expert_output=self.keep_expert_output, Wr_packed=self._router_packed(Wr), **norms)
            return out.view(B, T, D)
        if self.world_size > 1:
            if norms:
                raise NotImplementedError("LayerNorm fusion under expert parallelism needs bf16 activations")
            from . import ep   # fp32 / fp16 activations: the NCCL all-to-all formulation
            out = ep.ep_moe_layer(x.view(B * T, D), None if e is None else e.view(B * T, -1), Wr, br, packed,
                                  num_local_expert=self.num_experts, group=self.comm, top_k=1,
                                  gate_mode=ops.GATE_3M, act_type=act, ff_scale=ff_scale,
                                  residual=None if residual is None else residual.contiguous().view(B * T, D),
                                  keep_expert_output=self.keep_expert_output, x_len=x_len, seq_len=T)
            return out.view(B, T, D)
        res = ops.moe_layer(x, e, Wr, br, packed, residual=None if residual is None else residual.contiguous(),
                            x_len=x_len, top_k=1, gate_mode=ops.GATE_3M, act_type=act, ff_scale=ff_scale,
                            keep_expert_output=self.keep_expert_output, return_routing=return_routing,
                            Wr_packed=self._router_packed(Wr), **norms)
        return res if return_routing else res.out


def feed_forward_block(block: torch.nn.Module, x: torch.Tensor, embed: Optional[torch.Tensor],
                       mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The feed-forward part of FmoeConformerLayer.forward (trainer_3m_fix/layer/fmoe_transformer.py:144-166) as ONE
    call, for any module `block` that carries the reference's members: `feed_forward` (a LocalFmoeCatEmbedFeedForward),
    `norm_ff`, `ff_scale`, `normalize_before`, and -- for blocks with a convolution module -- `norm_final`:

        residual = x;  x = norm_ff(x);  x = feed_forward(x, embed, x_len);  x = residual + ff_scale * x;  x = norm_final(x)

    Post-norm blocks (normalize_before = False, :160-162) apply norm_ff behind the residual add instead; then norm_ff
    takes the place of the output norm and, if the block also has norm_final, that one runs as a second LayerNorm."""
    ff = block.feed_forward
    if getattr(block, "normalize_before", True):
        final = block.norm_final if getattr(block, "conv_module", None) is not None else None
        return ff(x, embed, mask, residual=x, ff_scale=block.ff_scale, norm_ff=block.norm_ff, norm_final=final)
    y = ff(x, embed, mask, residual=x, ff_scale=block.ff_scale, norm_final=block.norm_ff)
    if getattr(block, "conv_module", None) is not None:
        ln = block.norm_final
        y = ops.layernorm(y, ln.weight.detach().float().contiguous(), ln.bias.detach().float().contiguous(), ln.eps)
    return y
