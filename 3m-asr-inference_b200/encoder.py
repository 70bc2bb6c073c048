"""The Conformer-MoE encoder around the hot path: PyTorch operators for everything that is NOT
the fast_moe layer, the B200 kernels for what is.

Mirror of trainer_3m_fix/model/conformer_fmoe_localComm_catEmbed_domain_acc_hier.py:31-234 (`Net`) and of the modules
it is built from, with the reference's module tree and parameter names so that a 3M-ASR state_dict loads unchanged:

    embed          conformer_embed_domain_acc.py:26-181   6-block Conformer whose after_norm output is the router's `embed`
    subsampling    layer/subsampling.py:77-146            Conv2dSubsampling4 (lengths: MaskConv2dSamplePlugin twice)
    pos_enc        layer/positional_encoding.py:18-130    x * sqrt(d), pos_emb = pe[:, :T]  (no dropout at inference)
    blocks[i]      layer/fmoe_transformer.py:14-166       macaron FFN -> rel-pos MHA -> conv module -> fast_moe FFN -> norm_final
      .self_attn             layer/attention.py:114-384   scores = ((q + u) k^T + (q + v) p^T) / sqrt(d_k), masked softmax
      .conv_module           layer/convolution.py:18-167  masked fill, pointwise conv, GLU, depthwise conv, norm, swish, pointwise conv
      .feed_forward_macaron  layer/positionwise_feed_forward.py:17-52   w_2(act(w_1(x)))
      .feed_forward          this package's LocalFmoeCatEmbedFeedForward; norm_ff, x ff_scale, + residual and norm_final run
                             inside the same C-ABI call (b200moe_block_forward), see layer.feed_forward_block
    after_norm, out_linear   ...hier.py:150-152,197,222-226

The reference's forward methods emit TensorRT layers through a `network_helper`; here they take tensors.  Everything
outside `feed_forward` is library code (cuDNN / cuBLAS / SDPA through torch): it exists so that the encoder-level metric
of BASELINE.json (utterances per second through the 12 / 18-layer encoder) can be measured around the hot path.
`ffn_impl` lets the tests swap the fast_moe call for the CPU oracle; the product default has no CPU path.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .layer import LocalFmoeCatEmbedFeedForward, Swish, feed_forward_block


def _ops():
    from . import ops   # (needs libb200moe.so: imported where a CUDA tensor is in hand)
    return ops


def subsampled_lengths(x_len: torch.Tensor) -> torch.Tensor:
    """Valid frames after Conv2dSubsampling4.  The reference applies MaskConv2dSamplePluginDynamic(left_padding = 2,
    stride = 2) to the lengths twice (subsampling.py:117-133): out = (in - left_padding - 1) / stride + 1
    (TRTAPI++/plugin/mask_conv2d_sample_plugin/mask_conv2d_sample_kernel.cu:27-35) -- 206 frames -> 102 -> 50 tokens."""
    def once(n):
        return torch.div(n - 3, 2, rounding_mode="floor") + 1
    return torch.clamp(once(once(x_len)), min=0)


class Conv2dSubsampling4(nn.Module):
    def __init__(self, idim: int, odim: int, in_ch: int = 1):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_ch, odim, 3, 2), nn.ReLU(), nn.Conv2d(odim, odim, 3, 2), nn.ReLU())
        self.out = nn.Sequential(nn.Linear(odim * (((idim - 1) // 2 - 1) // 2), odim))
        self.subsampling_rate = 4
        self.right_context = 6
        self.in_ch = in_ch

    def forward(self, x: torch.Tensor, x_len: Optional[torch.Tensor]):
        B, T, Fdim = x.shape
        x = x.view(B, T, self.in_ch, Fdim // self.in_ch).transpose(1, 2)   # (B, in_ch, T, idim)
        x = self.conv(x)
        b, c, t, f = x.shape
        x = self.out(x.transpose(1, 2).contiguous().view(b, t, c * f))
        return x, (None if x_len is None else subsampled_lengths(x_len))


class RelPositionalEncoding(nn.Module):
    """positional_encoding.py:18-130; `pe` is a plain attribute upstream (not in the state_dict), a non-persistent buffer here."""

    def __init__(self, d_model: int, dropout_rate: float = 0.0, max_len: int = 5000):
        super().__init__()
        self.d_model = d_model
        self.xscale = math.sqrt(d_model)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0), persistent=False)

    def forward(self, x: torch.Tensor):
        if x.is_cuda and self.pe.dtype == x.dtype:   # RelPositionalEncodingPluginDynamic's job
            out, pos = _ops().rel_pos_encoding(x.contiguous(), self.pe[0], self.xscale)
            return out, pos.unsqueeze(0)
        return x * self.xscale, self.pe[:, : x.size(1)].to(x.dtype)


class RelPositionMultiHeadedAttention(nn.Module):
    def __init__(self, n_head: int, n_feat: int, dropout_rate: float = 0.0):
        super().__init__()
        assert n_feat % n_head == 0
        self.d_k = n_feat // n_head
        self.h = n_head
        self.linear_q = nn.Linear(n_feat, n_feat)
        self.linear_k = nn.Linear(n_feat, n_feat)
        self.linear_v = nn.Linear(n_feat, n_feat)
        self.linear_out = nn.Linear(n_feat, n_feat)
        self.linear_pos = nn.Linear(n_feat, n_feat, bias=False)
        self.pos_bias_u = nn.Parameter(torch.empty(self.h, self.d_k))
        self.pos_bias_v = nn.Parameter(torch.empty(self.h, self.d_k))
        nn.init.xavier_uniform_(self.pos_bias_u)
        nn.init.xavier_uniform_(self.pos_bias_v)

    def forward(self, x: torch.Tensor, x_len: Optional[torch.Tensor], pos_emb: torch.Tensor) -> torch.Tensor:
        B, T, _ = x.shape
        q = self.linear_q(x).view(B, T, self.h, self.d_k)
        k = self.linear_k(x).view(B, T, self.h, self.d_k).transpose(1, 2)
        v = self.linear_v(x).view(B, T, self.h, self.d_k).transpose(1, 2)
        p = self.linear_pos(pos_emb).view(1, -1, self.h, self.d_k).transpose(1, 2)        # (1, h, T, d_k)
        q_u = (q + self.pos_bias_u).transpose(1, 2)
        q_v = (q + self.pos_bias_v).transpose(1, 2)
        # matrix_bd (no rel_shift in this model: attention.py:352-365) goes in as the additive attention bias, the key
        # padding mask with it; SDPA then computes softmax(((q + u) k^T) / sqrt(d_k) + bias) v
        bias = torch.matmul(q_v, p.transpose(-2, -1)) / math.sqrt(self.d_k)               # (B, h, T, T)
        if x_len is not None:
            pad = torch.arange(T, device=x.device)[None, :] >= x_len[:, None]              # (B, T) True = padding key
            bias = bias.masked_fill(pad[:, None, None, :], float("-inf"))
        out = F.scaled_dot_product_attention(q_u, k, v, attn_mask=bias)
        if x_len is not None:   # att_masked_softmax_plugin zeroes the probabilities of padded keys; all-padding rows -> 0
            out = torch.nan_to_num(out, nan=0.0)
        return self.linear_out(out.transpose(1, 2).reshape(B, T, self.h * self.d_k))


class PositionwiseFeedForward(nn.Module):
    def __init__(self, idim: int, hidden_units: int, dropout_rate: float = 0.0, activation: nn.Module = nn.ReLU()):
        super().__init__()
        self.w_1 = nn.Linear(idim, hidden_units)
        self.activation = activation
        self.dropout = nn.Dropout(dropout_rate)
        self.w_2 = nn.Linear(hidden_units, idim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.w_2(self.activation(self.w_1(x)))


class ConvolutionModule(nn.Module):
    def __init__(self, channels: int, kernel_size: int = 15, activation: nn.Module = nn.ReLU(), norm: str = "batch_norm",
                 causal: bool = False, bias: bool = True):
        super().__init__()
        self.pointwise_conv1 = nn.Conv1d(channels, 2 * channels, 1, 1, 0, bias=bias)
        if causal:
            padding, self.lorder = 0, kernel_size - 1
        else:
            assert (kernel_size - 1) % 2 == 0
            padding, self.lorder = (kernel_size - 1) // 2, 0
        self.depthwise_conv = nn.Conv1d(channels, channels, kernel_size, 1, padding, groups=channels, bias=bias)
        assert norm in ("batch_norm", "layer_norm")
        self.use_layer_norm = norm == "layer_norm"
        self.norm = nn.LayerNorm(channels) if self.use_layer_norm else nn.BatchNorm1d(channels)
        self.pointwise_conv2 = nn.Conv1d(channels, channels, 1, 1, 0, bias=bias)
        self.activation = activation

    def forward(self, x: torch.Tensor, x_len: Optional[torch.Tensor]) -> torch.Tensor:
        x = x.transpose(1, 2)                                                            # (B, C, T)
        plugin = x.is_cuda                       # MaskedFillPluginDynamic / GluPluginDynamic as this repository's kernels
        len32 = None if x_len is None else x_len.to(torch.int32)

        def fill(t):
            if len32 is None:
                return t
            if plugin:
                return _ops().masked_fill(t.contiguous(), len32, 0.0)
            keep = (torch.arange(t.size(2), device=t.device)[None, :] < len32[:, None])[:, None, :]
            return t * keep

        x = fill(x)
        if self.lorder > 0:
            x = F.pad(x, (self.lorder, 0))
        x = self.pointwise_conv1(x)
        x = _ops().glu(x.contiguous()) if plugin else F.glu(x, dim=1)
        x = self.depthwise_conv(x)
        x = self.norm(x.transpose(1, 2)).transpose(1, 2) if self.use_layer_norm else self.norm(x)
        x = fill(self.pointwise_conv2(self.activation(x)))
        return x.transpose(1, 2)


class _ConformerLayerBase(nn.Module):
    """Members shared by ConformerEncoderLayer (embed net, dense FFN) and FmoeConformerLayer (fmoe_transformer.py:37-70)."""

    def __init__(self, size, self_attn, feed_forward, feed_forward_macaron, conv_module, dropout_rate=0.1,
                 normalize_before=True, concat_after=False):
        super().__init__()
        self.self_attn = self_attn
        self.feed_forward = feed_forward
        self.feed_forward_macaron = feed_forward_macaron
        self.conv_module = conv_module
        self.norm_ff = nn.LayerNorm(size, eps=1e-12)
        self.norm_mha = nn.LayerNorm(size, eps=1e-12)
        if feed_forward_macaron is not None:
            self.norm_ff_macaron = nn.LayerNorm(size, eps=1e-12)
            self.ff_scale = 0.5
        else:
            self.ff_scale = 1.0
        if conv_module is not None:
            self.norm_conv = nn.LayerNorm(size, eps=1e-12)
            self.norm_final = nn.LayerNorm(size, eps=1e-12)
        self.dropout = nn.Dropout(dropout_rate)
        self.size = size
        self.normalize_before = normalize_before
        self.concat_after = concat_after
        self.concat_linear = nn.Linear(size + size, size)
        if not normalize_before or concat_after:
            raise NotImplementedError("3M-ASR uses pre-norm blocks without concat_after (fmoe_transformer.py:109-111 asserts)")

    def _front(self, x, x_len, pos_emb):
        """Everything in front of the block's last feed-forward: macaron FFN, attention, convolution (:75-141)."""
        if self.feed_forward_macaron is not None:
            x = x + self.ff_scale * self.feed_forward_macaron(self.norm_ff_macaron(x))
        x = x + self.self_attn(self.norm_mha(x), x_len, pos_emb)
        if self.conv_module is not None:
            x = x + self.conv_module(self.norm_conv(x), x_len)
        return x


class ConformerEncoderLayer(_ConformerLayerBase):
    def forward(self, x, x_len, pos_emb):
        x = self._front(x, x_len, pos_emb)
        x = x + self.ff_scale * self.feed_forward(self.norm_ff(x))
        return self.norm_final(x) if self.conv_module is not None else x


class FmoeConformerLayer(_ConformerLayerBase):
    def forward(self, x, embed, x_len, pos_emb, ffn_impl: Optional[Callable] = None):
        x = self._front(x, x_len, pos_emb)
        # norm_ff -> fast_moe -> x ff_scale -> + residual -> norm_final: ONE call into the C ABI (:144-166)
        return (ffn_impl or feed_forward_block)(self, x.contiguous(), embed, x_len)


def _conformer_parts(attention_dim, attention_heads, cnn_module_kernel, activation, cnn_module_norm, causal, use_cnn_module):
    attn = RelPositionMultiHeadedAttention(attention_heads, attention_dim, 0.0)
    conv = ConvolutionModule(attention_dim, cnn_module_kernel, activation, cnn_module_norm, causal) if use_cnn_module else None
    return attn, conv


class ConformerEmbed(nn.Module):
    """conformer_embed_domain_acc.py `Net`: returns (out, lengths, after_norm output = the router's `embed`)."""

    def __init__(self, input_dim, output_dim, attention_heads=4, attention_dim=512, linear_units=1024, num_blocks=6,
                 macaron_style=True, use_cnn_module=True, cnn_module_kernel=15, causal=False, cnn_module_norm="batch_norm",
                 conv_subsample_in_ch=1, **_unused):
        super().__init__()
        act = Swish()
        self.subsampling = Conv2dSubsampling4(input_dim // conv_subsample_in_ch, attention_dim, conv_subsample_in_ch)
        self.pos_enc = RelPositionalEncoding(attention_dim)
        self.after_norm = nn.LayerNorm(attention_dim, eps=1e-12)
        blocks = []
        for _ in range(num_blocks):
            attn, conv = _conformer_parts(attention_dim, attention_heads, cnn_module_kernel, act, cnn_module_norm, causal,
                                          use_cnn_module)
            blocks.append(ConformerEncoderLayer(
                attention_dim, attn, PositionwiseFeedForward(attention_dim, linear_units, 0.0, act),
                PositionwiseFeedForward(attention_dim, linear_units, 0.0, act) if macaron_style else None, conv))
        self.blocks = nn.ModuleList(blocks)
        self.out_linear = nn.Linear(attention_dim, output_dim)

    def forward(self, xs, xs_len):
        xs, xs_len = self.subsampling(xs, xs_len)
        xs, pos_emb = self.pos_enc(xs)
        for layer in self.blocks:
            xs = layer(xs, xs_len, pos_emb)
        xs = self.after_norm(xs)
        return self.out_linear(xs), xs_len, xs


class ConformerMoEEncoder(nn.Module):
    """`Net` of conformer_fmoe_localComm_catEmbed_domain_acc_hier.py (constructor :32-196, forward :198-234)."""

    def __init__(self, input_dim: int, output_dim: int, attention_heads: int = 4, attention_dim: int = 256,
                 num_blocks: int = 6, macaron_style: bool = True, use_cnn_module: bool = True, cnn_module_kernel: int = 15,
                 causal: bool = False, cnn_module_norm: str = "batch_norm", conv_subsample_in_ch: int = 1,
                 embed_conf: Optional[dict] = None, moe_conf: Optional[dict] = None, **_unused):
        super().__init__()
        act = Swish()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.embed_conf = {"attention_heads": 4, "attention_dim": 512, "linear_units": 1024, "num_blocks": 6,
                           "macaron_style": True, "use_cnn_module": True, "cnn_module_kernel": 15, "causal": False,
                           "cnn_module_norm": "batch_norm", "conv_subsample_in_ch": 1}
        if isinstance(embed_conf, dict):
            self.embed_conf.update(embed_conf)
        self.embed = ConformerEmbed(input_dim, output_dim, **self.embed_conf)
        embed_dim = self.embed_conf["attention_dim"]
        self.moe_conf = {"rank": 0, "world_size": 1, "comm": None, "num_experts": 4, "hidden_units": 1024,
                         "dropout_rate": 0.0, "activation": act, "capacity_factor": -1.0,
                         "router_regularization": "l1_plus_importance", "router_with_bias": False,
                         "keep_expert_output": False, "rand_init_router": False}
        if moe_conf is not None:
            self.moe_conf.update(moe_conf)
        self.subsampling = Conv2dSubsampling4(input_dim // conv_subsample_in_ch, attention_dim, conv_subsample_in_ch)
        self.pos_enc = RelPositionalEncoding(attention_dim)
        self.normalize_before = True
        self.after_norm = nn.LayerNorm(attention_dim, eps=1e-12)
        self.after_norm_6 = nn.LayerNorm(attention_dim, eps=1e-12)     # (in the reference's state_dict, unused by forward)
        self.after_norm_12 = nn.LayerNorm(attention_dim, eps=1e-12)
        blocks = []
        for _ in range(num_blocks):
            attn, conv = _conformer_parts(attention_dim, attention_heads, cnn_module_kernel, act, cnn_module_norm, causal,
                                          use_cnn_module)
            dense_args = (attention_dim, self.moe_conf["hidden_units"], self.moe_conf["dropout_rate"], act)
            blocks.append(FmoeConformerLayer(
                attention_dim, attn, LocalFmoeCatEmbedFeedForward(attention_dim, embed_dim, **self.moe_conf),
                PositionwiseFeedForward(*dense_args) if macaron_style else None, conv))
        self.blocks = nn.ModuleList(blocks)
        self.out_linear = nn.Linear(attention_dim, output_dim)

    def to_inference(self, dtype: torch.dtype = torch.bfloat16) -> "ConformerMoEEncoder":
        """eval(); everything outside the fast_moe layers in `dtype`.  The fast_moe modules keep their fp32 masters (the
        router is used in fp32, the experts are packed to bf16 once); LayerNorms that feed the fused block call keep fp32
        gamma / beta, which is what the C ABI takes."""
        self.eval()
        moe = {id(p) for b in self.blocks for p in b.feed_forward.parameters()}
        fused_ln = {id(p) for b in self.blocks for n in ("norm_ff", "norm_final") if hasattr(b, n)
                    for p in getattr(b, n).parameters()}
        for p in self.parameters():
            if id(p) not in moe and id(p) not in fused_ln:
                p.data = p.data.to(dtype)
        for buf_owner in self.modules():
            for name, buf in list(buf_owner.named_buffers(recurse=False)):
                if buf.is_floating_point():
                    setattr(buf_owner, name, buf.to(dtype))
        return self

    def forward(self, xs: torch.Tensor, xs_len: Optional[torch.Tensor] = None, output_embed: bool = False,
                ffn_impl: Optional[Callable] = None):
        """xs [B, T, input_dim], xs_len [B] valid frames (int) or None -> logits [B, T', output_dim] (and the embed net's)."""
        embed_out, _, embed = self.embed(xs, xs_len)
        xs, sub_len = self.subsampling(xs, xs_len)
        xs, pos_emb = self.pos_enc(xs)
        len32 = None if sub_len is None else sub_len.to(torch.int32)
        embed = embed.contiguous()
        for layer in self.blocks:
            xs = layer(xs, embed, len32, pos_emb, ffn_impl)
        out = self.out_linear(self.after_norm(xs))
        return (out, embed_out) if output_embed else out


# The state_dict keys of one FmoeConformerLayer as the reference builds it (fmoe_transformer.py:37-70, attention.py:130-133,
# 286-290, convolution.py:35-80, positionwise_feed_forward.py:17-52,115-149), for the key-parity test.
REFERENCE_BLOCK_KEYS = (
    "self_attn.pos_bias_u", "self_attn.pos_bias_v", "self_attn.linear_q.weight", "self_attn.linear_q.bias",
    "self_attn.linear_k.weight", "self_attn.linear_k.bias", "self_attn.linear_v.weight", "self_attn.linear_v.bias",
    "self_attn.linear_out.weight", "self_attn.linear_out.bias", "self_attn.linear_pos.weight",
    "feed_forward.router_weights", "feed_forward.experts.w_1.weight", "feed_forward.experts.w_1.bias",
    "feed_forward.experts.w_2.weight", "feed_forward.experts.w_2.bias",
    "feed_forward_macaron.w_1.weight", "feed_forward_macaron.w_1.bias", "feed_forward_macaron.w_2.weight",
    "feed_forward_macaron.w_2.bias",
    "conv_module.pointwise_conv1.weight", "conv_module.pointwise_conv1.bias", "conv_module.depthwise_conv.weight",
    "conv_module.depthwise_conv.bias", "conv_module.norm.weight", "conv_module.norm.bias", "conv_module.norm.running_mean",
    "conv_module.norm.running_var", "conv_module.norm.num_batches_tracked", "conv_module.pointwise_conv2.weight",
    "conv_module.pointwise_conv2.bias",
    "norm_ff.weight", "norm_ff.bias", "norm_mha.weight", "norm_mha.bias", "norm_ff_macaron.weight", "norm_ff_macaron.bias",
    "norm_conv.weight", "norm_conv.bias", "norm_final.weight", "norm_final.bias", "concat_linear.weight", "concat_linear.bias",
)
