"""Host-side logic that needs no GPU: synthetic generator, module surfaces (parameter names / shapes as in the reference)."""
import torch

from conftest import pkg


def test_synth_is_deterministic_and_on_bf16_grid(synth):
    w1 = synth.make_weights(5, 8, 64, 128, 64, random_bias=True)
    w2 = synth.make_weights(5, 8, 64, 128, 64, random_bias=True)
    assert torch.equal(w1.W1, w2.W1) and torch.equal(w1.Wr, w2.Wr)
    for t in (w1.Wr, w1.W1, w1.W2, w1.b1):
        assert torch.equal(t, t.bfloat16().float())
    assert tuple(w1.W1.shape) == (8, 128, 64) and tuple(w1.W2.shape) == (8, 64, 128)
    x, e = synth.make_activations(6, 100, 64, 64, w1, top_k=2, margin=1e-3)
    logits = torch.cat([e, x], -1).double() @ w1.Wr.double()
    top = torch.topk(logits, 3, dim=-1).values
    assert float((top[:, :-1] - top[:, 1:]).min()) >= 1e-3
    # xavier_uniform(gain=0.5) bound for [E, out, in] as torch computes fans
    bound = 0.5 * (6.0 / (128 * 64 + 8 * 64)) ** 0.5
    assert float(w1.W1.abs().max()) <= bound * 1.01


def test_local_fmoe_cat_embed_state_dict_matches_reference_names():
    layer = pkg("layer")
    m = layer.LocalFmoeCatEmbedFeedForward(512, 512, num_experts=32, hidden_units=1024, activation=layer.Swish(),
                                           router_with_bias=True, rand_init_router=True)
    sd = m.state_dict()
    assert set(sd) == {"router_weights", "router_bias", "experts.w_1.weight", "experts.w_1.bias",
                       "experts.w_2.weight", "experts.w_2.bias"}
    assert tuple(sd["router_weights"].shape) == (1024, 32)
    assert tuple(sd["experts.w_1.weight"].shape) == (32, 1024, 512)
    assert tuple(sd["experts.w_1.bias"].shape) == (32, 1024)
    assert tuple(sd["experts.w_2.weight"].shape) == (32, 512, 1024)
    assert tuple(sd["experts.w_2.bias"].shape) == (32, 512)
    assert float(sd["experts.w_1.bias"].abs().max()) == 0.0  # reference init: bias = 0 (fmoe/layers.py:38)
    m2 = layer.LocalFmoeCatEmbedFeedForward(512, 512, num_experts=32, hidden_units=1024)
    assert float(m2.router_weights.abs().max()) == 0.0       # reference default router init: zeros (:135)


def test_fmoe_transformer_mlp_state_dict_matches_reference_names():
    fmoe = pkg("fmoe")
    m = fmoe.FMoETransformerMLP(num_expert=4, d_model=128, d_hidden=256, top_k=2)
    assert set(m.state_dict()) == {"gate.gate.weight", "gate.gate.bias", "experts.htoh4.weight", "experts.htoh4.bias",
                                   "experts.h4toh.weight", "experts.h4toh.bias"}
    assert tuple(m.experts.htoh4.weight.shape) == (4, 256, 128)
    assert tuple(m.gate.gate.weight.shape) == (4, 128)


def test_activation_codes():
    L = pkg("fmoe.layers")
    layer = pkg("layer")
    assert L.activation_code(layer.Swish()) == 0
    assert L.activation_code(torch.nn.SiLU()) == 0
    assert L.activation_code(torch.nn.ReLU()) == 1
    assert L.activation_code(torch.nn.GELU()) == 2


def test_plugin_registry_names_and_fields():
    plugin = pkg("plugin")
    reg = plugin.PluginRegistry()
    c = reg.get_plugin_creator("FMoEExpertPluginDynamic", "1", "")
    assert c is not None and reg.get_plugin_creator("FMoEExpertPluginDynamic", "2", "") is None
    p = c.create_plugin("plugin", {"data_type": 0, "num_expert": 32, "idim": 512, "hidden_units": 1024})
    assert p.fields == {"data_type": 0, "num_expert": 32, "idim": 512, "hidden_units": 1024, "act_type": 0}
    assert len(p.serialize()) == 32
    q = c.deserialize_plugin("plugin", p.serialize())
    assert q.fields == p.fields and q.get_plugin_type() == "FMoEExpertPluginDynamic"
    assert c.create_plugin("plugin", {"data_type": 9, "num_expert": 32, "idim": 512, "hidden_units": 1024}) is None
    assert reg.get_plugin_creator("SoftmaxTopKPluginDynamic", "1", "") is not None


def test_registered_torch_op_schema_and_shape_inference():
    """torch.ops.b200moe.fmoe_forward exists with the expected schema; its fake (meta) kernel infers the output shape
    without touching a GPU.  (The CUDA kernel itself is exercised in tests/test_gpu_layer.py.)"""
    pkg("ops")
    op = torch.ops.b200moe.fmoe_forward
    schema = str(op.default._schema)
    for name in ("x", "embed", "router_weight", "w1", "w2", "residual", "top_k", "gate_mode", "act_type", "ff_scale"):
        assert name in schema
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        x = torch.empty(50, 512, dtype=torch.bfloat16)
        out = op(x, None, torch.empty(512, 32), None, torch.empty(32, 1024, 512, dtype=torch.bfloat16), None,
                 torch.empty(32, 512, 1024, dtype=torch.bfloat16), None, None, 1, 0, 0, 1.0)
        assert out.shape == x.shape and out.dtype == x.dtype
