// TensorRT plugin shim over the C ABI: the three plugins on and next to the MoE path, under the
// names / versions / creator fields / input order / serialisation the reference's builder.py and infer.py expect, with
// their bodies replaced by calls into libb200moe.so.
//
//   FMoEExpertPluginDynamic  v1   TRTAPI++/plugin/fmoe_expert_plugin/fmoe_expert_plugin.{h,cpp}
//       fields data_type, num_expert, idim, hidden_units, act_type (int32 each); six inputs: input [.., idim], gate_idx
//       int32, w1_weight [E, H, D], w1_bias [E, H], w2_weight [E, D, H], w2_bias [E, D]; one un-weighted output shaped
//       like the input; 32-byte serialisation (the five fields + three zero ints)       -> b200moe_plugin_*
//   SoftmaxTopKPluginDynamic v1   TRTAPI++/plugin/softmax_topk_plugin/softmax_topk_plugin.{h,cpp}
//       field data_type; inputs logits [B, T, E], mask [B] int32; outputs value [B, T, 1], idx [B, T, 1] int32;
//       24-byte serialisation (data_type, axis_dim, k + three zero ints)                 -> b200moe_softmax_topk_enqueue
//   LayerNormPluginDynamic   v1   TRTAPI++/plugin/layer_norm_plugin/layer_norm_plugin.{h,cpp}
//       fields data_type, dim, eps; inputs x [.., dim], gamma [dim], beta [dim]; one output; 16-byte serialisation
//       (int32 data_type, size_t dim, float eps)                                         -> b200moe_layernorm
//
// Built only where TensorRT's headers exist:  make -C 3m-asr-inference_b200/csrc trt TRT_INCLUDE=/path/to/TensorRT/include
// (produces libb200moe_trt.so; load it next to libnvinfer like the reference's libtrtplugin++.so).  TensorRT is not in
// the development image, so the file is compile- and behaviour-checked against a stand-in for the TensorRT 8 plugin API
// (tests/trt_stub/NvInfer.h, tests/test_trt_shim.py): creator -> plugin -> serialise -> deserialise -> clone on the host,
// and enqueue on the GPU against the oracle, all through the same virtual calls TensorRT makes.
#include <NvInfer.h>

#include <cstring>
#include <string>
#include <vector>

#include "b200moe.h"

namespace b200moe_trt {

using namespace nvinfer1;

namespace {

int64_t volume(const Dims& d) {
  int64_t v = 1;
  for (int i = 0; i < d.nbDims; ++i) v *= d.d[i];
  return v;
}

// TensorRT DataType -> b200moe dtype code (0 fp32, 1 fp16, 2 bf16); -1 = not an activation type of these plugins
int dtype_code(DataType t) {
  if (t == DataType::kFLOAT) return B200MOE_F32;
  if (t == DataType::kHALF) return B200MOE_F16;
#if NV_TENSORRT_MAJOR >= 9
  if (t == DataType::kBF16) return B200MOE_BF16;
#endif
  return -1;
}

// the reference's "data_type" field: 0 = kFLOAT, 1 = kHALF; 2 = bf16 is accepted in addition where TensorRT has it
bool field_type_to_trt(int id, DataType* out) {
  if (id == 0) { *out = DataType::kFLOAT; return true; }
  if (id == 1) { *out = DataType::kHALF; return true; }
#if NV_TENSORRT_MAJOR >= 9
  if (id == 2) { *out = DataType::kBF16; return true; }
#endif
  return false;
}

int field_int(const PluginFieldCollection* fc, const char* name, int dflt) {
  for (int i = 0; fc != nullptr && i < fc->nbFields; ++i)
    if (fc->fields[i].name != nullptr && fc->fields[i].data != nullptr && std::strcmp(fc->fields[i].name, name) == 0)
      return *static_cast<const int32_t*>(fc->fields[i].data);
  return dflt;
}

float field_float(const PluginFieldCollection* fc, const char* name, float dflt) {
  for (int i = 0; fc != nullptr && i < fc->nbFields; ++i)
    if (fc->fields[i].name != nullptr && fc->fields[i].data != nullptr && std::strcmp(fc->fields[i].name, name) == 0)
      return *static_cast<const float*>(fc->fields[i].data);
  return dflt;
}

// what every plugin of this file shares: namespace bookkeeping and the trivial life-cycle calls
class PluginCommon : public IPluginV2DynamicExt {
 public:
  int32_t initialize() noexcept override { return 0; }
  void terminate() noexcept override {}
  void destroy() noexcept override { delete this; }
  void setPluginNamespace(const AsciiChar* ns) noexcept override { ns_ = ns ? ns : ""; }
  const AsciiChar* getPluginNamespace() const noexcept override { return ns_.c_str(); }
  const AsciiChar* getPluginVersion() const noexcept override { return "1"; }
  void configurePlugin(const DynamicPluginTensorDesc*, int32_t, const DynamicPluginTensorDesc*, int32_t) noexcept override {}

 protected:
  std::string ns_;
};

class CreatorCommon : public IPluginCreator {
 public:
  const AsciiChar* getPluginVersion() const noexcept override { return "1"; }
  const PluginFieldCollection* getFieldNames() noexcept override { return &fc_; }
  void setPluginNamespace(const AsciiChar* ns) noexcept override { ns_ = ns ? ns : ""; }
  const AsciiChar* getPluginNamespace() const noexcept override { return ns_.c_str(); }

 protected:
  void declare(std::initializer_list<PluginField> fields) {
    attrs_.assign(fields);
    fc_.nbFields = static_cast<int32_t>(attrs_.size());
    fc_.fields = attrs_.data();
  }
  std::vector<PluginField> attrs_;
  PluginFieldCollection fc_{0, nullptr};
  std::string ns_;
};

}  // namespace

// ======================================================================================================================
// FMoEExpertPluginDynamic
// ======================================================================================================================
class FMoEExpertPlugin final : public PluginCommon {
 public:
  // takes ownership of the handle
  explicit FMoEExpertPlugin(b200moe_plugin* h) : h_(h) {
    std::memset(cfg_, 0, sizeof(cfg_));
    if (h_ != nullptr && b200moe_plugin_serialization_size(h_) == sizeof(cfg_)) b200moe_plugin_serialize(h_, cfg_);
  }
  ~FMoEExpertPlugin() override { b200moe_plugin_destroy(h_); }

  bool ok() const { return h_ != nullptr; }
  int data_type() const { return cfg_[0]; }
  int num_expert() const { return cfg_[1]; }
  int idim() const { return cfg_[2]; }
  int hidden_units() const { return cfg_[3]; }

  IPluginV2DynamicExt* clone() const noexcept override {
    auto* p = new FMoEExpertPlugin(b200moe_plugin_clone(h_));
    p->setPluginNamespace(ns_.c_str());
    return p;
  }
  const AsciiChar* getPluginType() const noexcept override { return "FMoEExpertPluginDynamic"; }
  int32_t getNbOutputs() const noexcept override { return 1; }
  size_t getSerializationSize() const noexcept override { return b200moe_plugin_serialization_size(h_); }
  void serialize(void* buffer) const noexcept override { b200moe_plugin_serialize(h_, buffer); }

  DataType getOutputDataType(int32_t, const DataType* inputTypes, int32_t) const noexcept override { return inputTypes[0]; }
  DimsExprs getOutputDimensions(int32_t, const DimsExprs* inputs, int32_t, IExprBuilder&) noexcept override {
    return inputs[0];  // the un-weighted expert output has the input's shape
  }
  bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t nbInputs,
                                 int32_t nbOutputs) noexcept override {
    if (nbInputs != 6 || nbOutputs != 1 || pos < 0 || pos > 6) return false;
    if (io[pos].format != TensorFormat::kLINEAR) return false;
    if (pos == 1) return io[pos].type == DataType::kINT32;  // gate_idx
    return dtype_code(io[pos].type) == data_type();          // input, the four weight tensors, the output
  }
  size_t getWorkspaceSize(const PluginTensorDesc* in, int32_t, const PluginTensorDesc*, int32_t) const noexcept override {
    return b200moe_plugin_workspace_bytes(h_, tokens(in[0]));
  }
  int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs, void* const* outputs,
                  void* workspace, cudaStream_t stream) noexcept override {
    const int S = tokens(in[0]);
    return b200moe_plugin_enqueue(h_, inputs[0], static_cast<const int*>(inputs[1]), inputs[2], inputs[3], inputs[4],
                                  inputs[5], S, outputs[0], workspace, b200moe_plugin_workspace_bytes(h_, S), stream);
  }

 private:
  int tokens(const PluginTensorDesc& d) const { return idim() > 0 ? static_cast<int>(volume(d.dims) / idim()) : 0; }
  b200moe_plugin* h_;
  int32_t cfg_[8];  // the 32 serialised bytes: data_type, num_expert, idim, hidden_units, act_type, 0, 0, 0
};

class FMoEExpertPluginCreator final : public CreatorCommon {
 public:
  FMoEExpertPluginCreator() {
    declare({{"data_type", nullptr, PluginFieldType::kINT32, 1}, {"num_expert", nullptr, PluginFieldType::kINT32, 1},
             {"idim", nullptr, PluginFieldType::kINT32, 1}, {"hidden_units", nullptr, PluginFieldType::kINT32, 1},
             {"act_type", nullptr, PluginFieldType::kINT32, 1}});
  }
  const AsciiChar* getPluginName() const noexcept override { return "FMoEExpertPluginDynamic"; }
  IPluginV2* createPlugin(const AsciiChar*, const PluginFieldCollection* fc) noexcept override {
    const int type_id = field_int(fc, "data_type", -1);
    DataType t;
    if (!field_type_to_trt(type_id, &t)) return nullptr;  // (the reference: "invalid type_id")
    return adopt(b200moe_plugin_create(type_id, field_int(fc, "num_expert", 0), field_int(fc, "idim", 0),
                                       field_int(fc, "hidden_units", 0), field_int(fc, "act_type", 0)));
  }
  IPluginV2* deserializePlugin(const AsciiChar*, const void* data, size_t length) noexcept override {
    return adopt(b200moe_plugin_deserialize(data, length));
  }

 private:
  IPluginV2* adopt(b200moe_plugin* h) {
    if (h == nullptr) return nullptr;
    auto* p = new FMoEExpertPlugin(h);
    p->setPluginNamespace(ns_.c_str());
    return p;
  }
};

// ======================================================================================================================
// SoftmaxTopKPluginDynamic (top-1: value = max softmax probability, idx = its expert)
// ======================================================================================================================
class SoftmaxTopKPlugin final : public PluginCommon {
 public:
  SoftmaxTopKPlugin(int data_type, int axis_dim, int k) : data_type_(data_type), axis_dim_(axis_dim), k_(k) {}
  SoftmaxTopKPlugin(const void* data, size_t length) {
    int32_t w[6] = {0, -1, 1, 0, 0, 0};
    if (data != nullptr && length >= sizeof(w)) std::memcpy(w, data, sizeof(w));
    data_type_ = w[0];
    axis_dim_ = w[1];
    k_ = w[2];
  }
  IPluginV2DynamicExt* clone() const noexcept override {
    auto* p = new SoftmaxTopKPlugin(data_type_, axis_dim_, k_);
    p->setPluginNamespace(ns_.c_str());
    return p;
  }
  const AsciiChar* getPluginType() const noexcept override { return "SoftmaxTopKPluginDynamic"; }
  int32_t getNbOutputs() const noexcept override { return 2; }
  size_t getSerializationSize() const noexcept override { return 6 * sizeof(int32_t); }
  void serialize(void* buffer) const noexcept override {
    const int32_t w[6] = {data_type_, axis_dim_, k_, 0, 0, 0};
    std::memcpy(buffer, w, sizeof(w));
  }
  DataType getOutputDataType(int32_t index, const DataType* inputTypes, int32_t) const noexcept override {
    return index == 0 ? inputTypes[0] : DataType::kINT32;  // value in the logits' type, idx int32
  }
  DimsExprs getOutputDimensions(int32_t, const DimsExprs* inputs, int32_t, IExprBuilder& eb) noexcept override {
    DimsExprs out = inputs[0];
    out.d[out.nbDims - 1] = eb.constant(1);  // [B, T, E] -> [B, T, 1] for both outputs
    return out;
  }
  bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t nbInputs,
                                 int32_t nbOutputs) noexcept override {
    if (nbInputs != 2 || nbOutputs != 2 || pos < 0 || pos > 3) return false;
    if (io[pos].format != TensorFormat::kLINEAR) return false;
    if (pos == 1 || pos == 3) return io[pos].type == DataType::kINT32;  // mask, idx
    return dtype_code(io[pos].type) == data_type_;                       // logits, value
  }
  size_t getWorkspaceSize(const PluginTensorDesc*, int32_t, const PluginTensorDesc*, int32_t) const noexcept override {
    return 0;
  }
  int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs, void* const* outputs,
                  void*, cudaStream_t stream) noexcept override {
    if (in[0].dims.nbDims != 3) return -1;  // the reference asserts [B, T, E]
    return b200moe_softmax_topk_enqueue(inputs[0], static_cast<const int*>(inputs[1]), in[0].dims.d[0], in[0].dims.d[1],
                                        in[0].dims.d[2], data_type_, outputs[0], static_cast<int*>(outputs[1]), stream);
  }

 private:
  int data_type_ = 0, axis_dim_ = -1, k_ = 1;
};

class SoftmaxTopKPluginCreator final : public CreatorCommon {
 public:
  SoftmaxTopKPluginCreator() { declare({{"data_type", nullptr, PluginFieldType::kINT32, 1}}); }
  const AsciiChar* getPluginName() const noexcept override { return "SoftmaxTopKPluginDynamic"; }
  IPluginV2* createPlugin(const AsciiChar*, const PluginFieldCollection* fc) noexcept override {
    const int type_id = field_int(fc, "data_type", -1);
    DataType t;
    if (!field_type_to_trt(type_id, &t)) return nullptr;
    auto* p = new SoftmaxTopKPlugin(type_id, -1, 1);
    p->setPluginNamespace(ns_.c_str());
    return p;
  }
  IPluginV2* deserializePlugin(const AsciiChar*, const void* data, size_t length) noexcept override {
    auto* p = new SoftmaxTopKPlugin(data, length);
    p->setPluginNamespace(ns_.c_str());
    return p;
  }
};

// ======================================================================================================================
// LayerNormPluginDynamic (gamma / beta arrive as inputs; fp32 activations carry fp32 gamma / beta)
// ======================================================================================================================
class LayerNormPlugin final : public PluginCommon {
 public:
  LayerNormPlugin(int data_type, int dim, float eps) : data_type_(data_type), dim_(dim), eps_(eps) {}
  // the reference's byte stream: int32 data_type, size_t dim, float eps, written back to back (16 bytes)
  LayerNormPlugin(const void* data, size_t length) {
    if (data != nullptr && length >= 16) {
      uint64_t dim = 0;
      std::memcpy(&data_type_, data, 4);
      std::memcpy(&dim, static_cast<const char*>(data) + 4, 8);
      std::memcpy(&eps_, static_cast<const char*>(data) + 12, 4);
      dim_ = static_cast<int>(dim);
    }
  }
  IPluginV2DynamicExt* clone() const noexcept override {
    auto* p = new LayerNormPlugin(data_type_, dim_, eps_);
    p->setPluginNamespace(ns_.c_str());
    return p;
  }
  const AsciiChar* getPluginType() const noexcept override { return "LayerNormPluginDynamic"; }
  int32_t getNbOutputs() const noexcept override { return 1; }
  size_t getSerializationSize() const noexcept override { return 16; }
  void serialize(void* buffer) const noexcept override {
    const uint64_t dim = static_cast<uint64_t>(dim_);
    std::memcpy(buffer, &data_type_, 4);
    std::memcpy(static_cast<char*>(buffer) + 4, &dim, 8);
    std::memcpy(static_cast<char*>(buffer) + 12, &eps_, 4);
  }
  DataType getOutputDataType(int32_t, const DataType* inputTypes, int32_t) const noexcept override { return inputTypes[0]; }
  DimsExprs getOutputDimensions(int32_t, const DimsExprs* inputs, int32_t, IExprBuilder&) noexcept override {
    return inputs[0];
  }
  bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* io, int32_t nbInputs,
                                 int32_t nbOutputs) noexcept override {
    if (nbInputs != 3 || nbOutputs != 1 || pos < 0 || pos > 3) return false;
    if (io[pos].format != TensorFormat::kLINEAR) return false;
    if (pos == 1 || pos == 2) return io[pos].type == DataType::kFLOAT;  // gamma / beta: fp32 (b200moe_layernorm)
    return dtype_code(io[pos].type) == data_type_;
  }
  size_t getWorkspaceSize(const PluginTensorDesc*, int32_t, const PluginTensorDesc*, int32_t) const noexcept override {
    return 0;
  }
  int32_t enqueue(const PluginTensorDesc* in, const PluginTensorDesc*, const void* const* inputs, void* const* outputs,
                  void*, cudaStream_t stream) noexcept override {
    if (dim_ <= 0) return -1;
    const int S = static_cast<int>(volume(in[0].dims) / dim_);
    return b200moe_layernorm(inputs[0], static_cast<const float*>(inputs[1]), static_cast<const float*>(inputs[2]), eps_,
                             S, dim_, data_type_, outputs[0], stream);
  }

 private:
  int data_type_ = 0, dim_ = 0;
  float eps_ = 1e-12f;
};

class LayerNormPluginCreator final : public CreatorCommon {
 public:
  LayerNormPluginCreator() {
    declare({{"data_type", nullptr, PluginFieldType::kINT32, 1}, {"dim", nullptr, PluginFieldType::kINT32, 1},
             {"eps", nullptr, PluginFieldType::kFLOAT32, 1}});
  }
  const AsciiChar* getPluginName() const noexcept override { return "LayerNormPluginDynamic"; }
  IPluginV2* createPlugin(const AsciiChar*, const PluginFieldCollection* fc) noexcept override {
    const int type_id = field_int(fc, "data_type", -1);
    DataType t;
    if (!field_type_to_trt(type_id, &t)) return nullptr;
    auto* p = new LayerNormPlugin(type_id, field_int(fc, "dim", 0), field_float(fc, "eps", 1e-12f));
    p->setPluginNamespace(ns_.c_str());
    return p;
  }
  IPluginV2* deserializePlugin(const AsciiChar*, const void* data, size_t length) noexcept override {
    auto* p = new LayerNormPlugin(data, length);
    p->setPluginNamespace(ns_.c_str());
    return p;
  }
};

REGISTER_TENSORRT_PLUGIN(FMoEExpertPluginCreator);
REGISTER_TENSORRT_PLUGIN(SoftmaxTopKPluginCreator);
REGISTER_TENSORRT_PLUGIN(LayerNormPluginCreator);

}  // namespace b200moe_trt
