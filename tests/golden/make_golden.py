#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the REFERENCE's own Python code for the fast_moe path on CPU.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

What is executed from the reference tree (unmodified, imported from /root/reference/trainer_3m_fix):
  fmoe.gates.NaiveGate.forward                                   fmoe/gates.py:51-66
  fmoe.functions.moe_prepare_forward / MOEScatter / MOEbiasLinear / MOEGather   fmoe/functions.py:13-216
  model.dfsmn_base_fmoe_localComm_catEmbed._fmoe_general_global_forward          :30-58   (the only live copy)
  model.dfsmn_base_fmoe_localComm_catEmbed.cFSMN_layer.gate                      :166-192 (the live 3M router)
  fmoe.layers.FMoELinear (parameter shapes + init)                               fmoe/layers.py:21-40
  utils.common.Swish                                                             utils/common.py:24-28
  layer.fmoe_transformer.FmoeConformerLayer (constructor: norm_ff, norm_final, ff_scale)  layer/fmoe_transformer.py:33-70

What is NOT in the reference tree: `fmoe_cuda`, the CUDA extension of laekov/fastmoe (Tencent-modified, un-vendored,
un-pinned).  Its four primitives on this path are stubbed below with their published semantics
(fastmoe cuda/local_exchange.cuh, cuda/parallel_linear.cuh):
  local_scatter(inp, pos)        -> buf[i] = inp[pos[i]]
  local_gather(buf, pos)         -> out[pos[i]] = buf[i]
  forward(inp, weight, cnt, ..)  -> per expert e, rows of e:  out = inp_e @ weight[e].T
The FMoE.forward wiring that is commented out in fmoe/layers.py:186-210 is re-stated line by line in `fmoe_forward`.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/trainer_3m_fix"


def install_fmoe_cuda_stub():
    stub = types.ModuleType("fmoe_cuda")

    def local_scatter(inp, pos):
        return (inp[pos],)

    def local_gather(buf, pos):
        out = torch.empty_like(buf)
        out[pos] = buf
        return (out,)

    def forward(inp, weight, fwd_expert_count, capacity=-1, training=False):
        outs = []
        base = 0
        for e, c in enumerate(fwd_expert_count.tolist()):
            outs.append(inp[base:base + c] @ weight[e].t())
            base += c
        return (torch.cat(outs, dim=0) if outs else inp.new_zeros(0, weight.shape[1]),)

    def ensure_nccl(*a, **k):
        return None

    stub.local_scatter = local_scatter
    stub.local_gather = local_gather
    stub.forward = forward
    stub.ensure_nccl = ensure_nccl
    sys.modules["fmoe_cuda"] = stub


def bits(t):
    """bf16-grid fp32 tensor -> its upper 16 bits (exact); halves the fixture size. Loader: tests/conftest.py."""
    a = t.detach().contiguous().numpy().view(np.uint32)
    assert (a & 0xFFFF).max() == 0, "tensor is not on the bf16 grid"
    return (a >> 16).astype(np.uint16)


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not present; golden vectors can only be regenerated in the build container")
    install_fmoe_cuda_stub()
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    F = importlib.import_module("fmoe.functions")
    gates = importlib.import_module("fmoe.gates")
    layers = importlib.import_module("fmoe.layers")
    dfsmn = importlib.import_module("model.dfsmn_base_fmoe_localComm_catEmbed")
    common = importlib.import_module("utils.common")
    synth = importlib.import_module("3m-asr-inference_b200.synth")

    torch.set_num_threads(1)
    swish = common.Swish()

    def expert_fn_factory(W1, b1, W2, b2, act):
        # Expert.forward (layer/positionwise_feed_forward.py:105-112) with FMoELinear.forward restored from its
        # commented body (fmoe/layers.py:46-51): MOEbiasLinear.apply(inp, weight, bias, fwd_expert_count, ...)
        def expert_fn(inp, fwd_expert_count, capacity=-1):
            h = F.MOEbiasLinear.apply(inp, W1, b1, fwd_expert_count, capacity, False)
            h = act(h)
            return F.MOEbiasLinear.apply(h, W2, b2, fwd_expert_count, capacity, False)
        return expert_fn

    # ---------------- case A: 3M-ASR router (top-1, cat-embed), Swish experts, macaron residual ----------------
    E, D, Demb, H, S = 4, 128, 128, 128, 37
    w = synth.make_weights(1001, E, D, H, Demb, random_bias=True)
    x, embed = synth.make_activations(2001, S, D, Demb, w, top_k=1)
    # parameter holders built by the reference's own class: checks [E, out, in] / [E, out] shapes
    l1 = layers.FMoELinear(E, D, H, bias=True)
    l2 = layers.FMoELinear(E, H, D, bias=True)
    assert tuple(l1.weight.shape) == (E, H, D) and tuple(l1.bias.shape) == (E, H)
    assert tuple(l2.weight.shape) == (E, D, H) and tuple(l2.bias.shape) == (E, D)
    layer = dfsmn.cFSMN_layer(D, Demb, hid_dim=H, mem_dim=D, num_experts=E, rank=0, world_size=1,
                              capacity_factor=-1, skip_connect=True, rand_init_router=True)
    with torch.no_grad():
        layer.rooter_weights.copy_(w.Wr)
        embed_inputs = torch.cat([embed, x], dim=-1)          # dfsmn_base_...py:210
        gate_idx, gate_value, _aux, _n = layer.gate(embed_inputs)
        expert_fn = expert_fn_factory(w.W1, w.b1, w.W2, w.b2, swish)
        expert_outputs = dfsmn._fmoe_general_global_forward(x, gate_idx, expert_fn, E, 1, capacity=-1)
        _pos, local_cnt, _g, fwd_cnt, fwd_bs = F.moe_prepare_forward(gate_idx, E, 1)
        weighted = expert_outputs * gate_value.unsqueeze(1)  # :217-218
        final = x + 0.5 * weighted                            # residual + ff_scale * y (fmoeExMarc_transformer.py:153-154)
    np.savez_compressed(
        os.path.join(HERE, "case_3m_top1.npz"),
        x_bf16=bits(x), embed_bf16=bits(embed), Wr_bf16=bits(w.Wr), W1_bf16=bits(w.W1), b1_bf16=bits(w.b1),
        W2_bf16=bits(w.W2), b2_bf16=bits(w.b2), gate_idx=gate_idx.numpy().astype(np.int64), gate_value=gate_value.numpy(),
        expert_count=local_cnt.numpy().astype(np.int64), fwd_batch_size=np.int64(fwd_bs),
        expert_outputs=expert_outputs.numpy(), weighted=weighted.numpy(), final=final.numpy(),
        ff_scale=np.float32(0.5))
    print("case_3m_top1: counts", local_cnt.tolist())

    # ---------------- case B: FastMoE NaiveGate top-2, GELU experts (FMoETransformerMLP defaults) ----------------
    E, D, H, S, K = 8, 128, 128, 29, 2
    w = synth.make_weights(1002, E, D, H, 0, router_bias=True, random_bias=True)
    x, _ = synth.make_activations(2002, S, D, 0, w, top_k=K)
    gate = gates.NaiveGate(D, E, 1, top_k=K)
    with torch.no_grad():
        gate.gate.weight.copy_(w.Wr.t())   # nn.Linear stores [E, d]
        gate.gate.bias.copy_(w.br)
        gelu = torch.nn.GELU()

        def fmoe_forward(inp):
            # fmoe/layers.py:186-210 (commented in the reference), restated
            gate_top_k_idx, gate_score, _ = gate(inp)
            inp_rep = inp.repeat_interleave(repeats=K, dim=0)
            expert_fn = expert_fn_factory(w.W1, w.b1, w.W2, w.b2, gelu)
            y = dfsmn._fmoe_general_global_forward(inp_rep, gate_top_k_idx, expert_fn, E, 1)
            y = y.view(-1, K, D)
            out = torch.bmm(gate_score, y).reshape(-1, D)
            return gate_top_k_idx, gate_score, y, out

        idx, score, y_entries, out = fmoe_forward(x)
        _pos, local_cnt, _g, _f, _bs = F.moe_prepare_forward(idx, E, 1)
    np.savez_compressed(
        os.path.join(HERE, "case_naive_top2.npz"),
        x_bf16=bits(x), Wr_bf16=bits(w.Wr), br_bf16=bits(w.br), W1_bf16=bits(w.W1), b1_bf16=bits(w.b1),
        W2_bf16=bits(w.W2), b2_bf16=bits(w.b2), gate_idx=idx.view(S, K).numpy().astype(np.int64), gate_score=score.view(S, K).numpy(),
        expert_count=local_cnt.numpy().astype(np.int64), y_entries=y_entries.numpy(), out=out.numpy())
    print("case_naive_top2: counts", local_cnt.tolist())

    # ---------------- case C: the Conformer block's feed-forward part: norm_ff -> 3M MoE -> x ff_scale + residual -> norm_final
    # The LayerNorms and ff_scale come from the reference's own FmoeConformerLayer constructor; its forward emits
    # TensorRT graph layers, so the wiring (fmoe_transformer.py:144-166) is restated with the module's members.
    fmoe_tr = importlib.import_module("layer.fmoe_transformer")
    E, D, Demb, H, S = 4, 128, 128, 128, 41
    w = synth.make_weights(1003, E, D, H, Demb, random_bias=True)
    x, embed = synth.make_activations(2003, S, D, Demb, w, top_k=1)
    x = (x * 3.0 + 0.25).bfloat16().float()   # rows that are not already zero-mean / unit-variance
    block = fmoe_tr.FmoeConformerLayer(D, self_attn=None, feed_forward=None, feed_forward_macaron=torch.nn.Identity(),
                                       conv_module=torch.nn.Identity())
    assert block.ff_scale == 0.5 and block.norm_ff.eps == 1e-12 and block.norm_final.eps == 1e-12
    g = torch.Generator().manual_seed(3003)
    with torch.no_grad():
        for ln in (block.norm_ff, block.norm_final):
            ln.weight.copy_((1.0 + 0.2 * torch.randn(D, generator=g)).bfloat16().float())
            ln.bias.copy_((0.1 * torch.randn(D, generator=g)).bfloat16().float())
        layer = dfsmn.cFSMN_layer(D, Demb, hid_dim=H, mem_dim=D, num_experts=E, rank=0, world_size=1,
                                  capacity_factor=-1, skip_connect=True, rand_init_router=True)
        layer.rooter_weights.copy_(w.Wr)
        residual = x
        xn = block.norm_ff(x)                                                 # :145-148
        gate_idx, gate_value, _aux, _n = layer.gate(torch.cat([embed, xn], dim=-1))
        expert_fn = expert_fn_factory(w.W1, w.b1, w.W2, w.b2, swish)
        y = dfsmn._fmoe_general_global_forward(xn, gate_idx, expert_fn, E, 1, capacity=-1) * gate_value.unsqueeze(1)
        pre = residual + block.ff_scale * y                                   # :155-158
        out = block.norm_final(pre)                                           # :164-166
        _pos, local_cnt, _g, _f, _bs = F.moe_prepare_forward(gate_idx, E, 1)
    np.savez_compressed(
        os.path.join(HERE, "case_block_3m.npz"),
        x_bf16=bits(x), embed_bf16=bits(embed), Wr_bf16=bits(w.Wr), W1_bf16=bits(w.W1), b1_bf16=bits(w.b1),
        W2_bf16=bits(w.W2), b2_bf16=bits(w.b2), ff_gamma_bf16=bits(block.norm_ff.weight.detach()),
        ff_beta_bf16=bits(block.norm_ff.bias.detach()), final_gamma_bf16=bits(block.norm_final.weight.detach()),
        final_beta_bf16=bits(block.norm_final.bias.detach()), xn=xn.numpy(), gate_idx=gate_idx.numpy().astype(np.int64),
        gate_value=gate_value.numpy(), expert_count=local_cnt.numpy().astype(np.int64), pre_norm=pre.numpy(),
        out=out.numpy(), ff_scale=np.float32(block.ff_scale), eps=np.float64(block.norm_ff.eps))
    print("case_block_3m: counts", local_cnt.tolist())

    # ---------------- case D: the repo's own dimensions (E=32, D=512, embed 512, H=1024), 3M router, Swish experts ----------------
    # 128 MiB of fp32 expert weights do not belong in a fixture: they are re-generated from the seed by the same
    # synth.make_weights call in the tests (CPU mt19937 stream), and pinned here by checksums and sample values.
    E, D, Demb, H, S = 32, 512, 512, 1024, 64
    w = synth.make_weights(1004, E, D, H, Demb, random_bias=True)
    x, embed = synth.make_activations(2004, S, D, Demb, w, top_k=1)
    l1 = layers.FMoELinear(E, D, H, bias=True)
    l2 = layers.FMoELinear(E, H, D, bias=True)
    assert tuple(l1.weight.shape) == (E, H, D) and tuple(l2.weight.shape) == (E, D, H)
    layer = dfsmn.cFSMN_layer(D, Demb, hid_dim=H, mem_dim=D, num_experts=E, rank=0, world_size=1,
                              capacity_factor=-1, skip_connect=True, rand_init_router=True)
    torch.set_num_threads(8)
    with torch.no_grad():
        layer.rooter_weights.copy_(w.Wr)
        gate_idx, gate_value, _aux, _n = layer.gate(torch.cat([embed, x], dim=-1))
        expert_fn = expert_fn_factory(w.W1, w.b1, w.W2, w.b2, swish)
        expert_outputs = dfsmn._fmoe_general_global_forward(x, gate_idx, expert_fn, E, 1, capacity=-1)
        _pos, local_cnt, _g, fwd_cnt, fwd_bs = F.moe_prepare_forward(gate_idx, E, 1)
        final = x + 0.5 * expert_outputs * gate_value.unsqueeze(1)
    torch.set_num_threads(1)
    probe = torch.arange(0, E * H * D, 1048573)[:16]
    np.savez_compressed(
        os.path.join(HERE, "case_3m_repo_dims.npz"),
        x_bf16=bits(x), embed_bf16=bits(embed), Wr_bf16=bits(w.Wr), weight_seed=np.int64(1004),
        W1_sum=np.float64(w.W1.double().sum()), W2_sum=np.float64(w.W2.double().sum()),
        b1_sum=np.float64(w.b1.double().sum()), b2_sum=np.float64(w.b2.double().sum()),
        W1_probe=w.W1.reshape(-1)[probe].numpy(), W2_probe=w.W2.reshape(-1)[probe].numpy(),
        gate_idx=gate_idx.numpy().astype(np.int64), gate_value=gate_value.numpy(),
        expert_count=local_cnt.numpy().astype(np.int64), fwd_batch_size=np.int64(fwd_bs),
        expert_outputs=expert_outputs.numpy().astype(np.float32), final=final.numpy().astype(np.float32),
        ff_scale=np.float32(0.5))
    print("case_3m_repo_dims: counts", local_cnt.tolist())


if __name__ == "__main__":
    main()
