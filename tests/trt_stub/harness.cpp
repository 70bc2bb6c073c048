// Drives the TensorRT shim through the stand-in API the way a TensorRT builder / runtime would -- TEST INFRASTRUCTURE.
// Exposes a few extern "C" entry points for tests/test_trt_shim.py (ctypes).
#include <NvInfer.h>

#include <cstring>
#include <string>

using namespace nvinfer1;

namespace {

IPluginCreator* find_creator(const char* name) {
  for (IPluginCreator* c : stubRegistry())
    if (std::strcmp(c->getPluginName(), name) == 0 && std::strcmp(c->getPluginVersion(), "1") == 0) return c;
  return nullptr;
}

Dims make_dims(const int* d, int n) {
  Dims out{};
  out.nbDims = n;
  for (int i = 0; i < n; ++i) out.d[i] = d[i];
  return out;
}

struct ConstExpr : IDimensionExpr {
  explicit ConstExpr(int v) : v_(v) {}
  bool isConstant() const noexcept override { return true; }
  int32_t getConstantValue() const noexcept override { return v_; }
  int v_;
};
struct Builder : IExprBuilder {
  const IDimensionExpr* constant(int32_t v) noexcept override {
    pool.emplace_back(new ConstExpr(v));
    return pool.back();
  }
  std::vector<ConstExpr*> pool;
  ~Builder() {
    for (auto* e : pool) delete e;
  }
};

}  // namespace

extern "C" {

int shim_num_creators() { return static_cast<int>(stubRegistry().size()); }

// names of the creator's fields, comma separated, into out (host)
int shim_field_names(const char* plugin, char* out, int cap) {
  IPluginCreator* c = find_creator(plugin);
  if (!c) return -1;
  std::string s;
  const PluginFieldCollection* fc = c->getFieldNames();
  for (int i = 0; i < fc->nbFields; ++i) s += std::string(i ? "," : "") + fc->fields[i].name;
  std::strncpy(out, s.c_str(), cap - 1);
  out[cap - 1] = 0;
  return fc->nbFields;
}

// createPlugin from int fields (names / values) + one optional float field -> opaque IPluginV2DynamicExt*
void* shim_create(const char* plugin, int n, const char** names, const int* values, const char* fname, float fvalue) {
  IPluginCreator* c = find_creator(plugin);
  if (!c) return nullptr;
  std::vector<PluginField> f;
  for (int i = 0; i < n; ++i) f.push_back({names[i], &values[i], PluginFieldType::kINT32, 1});
  if (fname) f.push_back({fname, &fvalue, PluginFieldType::kFLOAT32, 1});
  PluginFieldCollection fc{static_cast<int32_t>(f.size()), f.data()};
  return c->createPlugin("layer", &fc);
}
void* shim_deserialize(const char* plugin, const void* data, size_t n) {
  IPluginCreator* c = find_creator(plugin);
  return c ? c->deserializePlugin("layer", data, n) : nullptr;
}
void* shim_clone(void* p) { return static_cast<IPluginV2DynamicExt*>(p)->clone(); }
void shim_destroy(void* p) { static_cast<IPluginV2DynamicExt*>(p)->destroy(); }
const char* shim_type(void* p) { return static_cast<IPluginV2DynamicExt*>(p)->getPluginType(); }
const char* shim_version(void* p) { return static_cast<IPluginV2DynamicExt*>(p)->getPluginVersion(); }
int shim_nb_outputs(void* p) { return static_cast<IPluginV2DynamicExt*>(p)->getNbOutputs(); }
size_t shim_serialization_size(void* p) { return static_cast<IPluginV2DynamicExt*>(p)->getSerializationSize(); }
void shim_serialize(void* p, void* buf) { static_cast<IPluginV2DynamicExt*>(p)->serialize(buf); }
void shim_set_namespace(void* p, const char* ns) { static_cast<IPluginV2DynamicExt*>(p)->setPluginNamespace(ns); }
const char* shim_get_namespace(void* p) { return static_cast<IPluginV2DynamicExt*>(p)->getPluginNamespace(); }

// supportsFormatCombination for a list of (type, format) over inputs + outputs
int shim_supports(void* p, int pos, const int* types, int nb_in, int nb_out) {
  std::vector<PluginTensorDesc> d(nb_in + nb_out);
  for (int i = 0; i < nb_in + nb_out; ++i) {
    d[i] = PluginTensorDesc{};
    d[i].type = static_cast<DataType>(types[i]);
    d[i].format = TensorFormat::kLINEAR;
  }
  return static_cast<IPluginV2DynamicExt*>(p)->supportsFormatCombination(pos, d.data(), nb_in, nb_out) ? 1 : 0;
}

// getOutputDimensions on constant input dims -> out_dims (returns nbDims)
int shim_output_dims(void* p, int out_index, const int* in_dims, int nb_dims, int nb_inputs, int* out_dims) {
  Builder b;
  std::vector<DimsExprs> in(nb_inputs);
  for (auto& e : in) {
    e.nbDims = nb_dims;
    for (int i = 0; i < nb_dims; ++i) e.d[i] = b.constant(in_dims[i]);
  }
  DimsExprs o = static_cast<IPluginV2DynamicExt*>(p)->getOutputDimensions(out_index, in.data(), nb_inputs, b);
  for (int i = 0; i < o.nbDims; ++i) out_dims[i] = o.d[i]->getConstantValue();
  return o.nbDims;
}

size_t shim_workspace(void* p, const int* in0_dims, int nb_dims, int nb_in, int nb_out) {
  std::vector<PluginTensorDesc> in(nb_in), out(nb_out);
  in[0].dims = make_dims(in0_dims, nb_dims);
  return static_cast<IPluginV2DynamicExt*>(p)->getWorkspaceSize(in.data(), nb_in, out.data(), nb_out);
}

// enqueue: device pointers for inputs / outputs, dims of input 0
int shim_enqueue(void* p, const int* in0_dims, int nb_dims, int nb_in, const void* const* inputs, int nb_out,
                 void* const* outputs, void* workspace, cudaStream_t stream) {
  std::vector<PluginTensorDesc> in(nb_in), out(nb_out);
  in[0].dims = make_dims(in0_dims, nb_dims);
  return static_cast<IPluginV2DynamicExt*>(p)->enqueue(in.data(), out.data(), inputs, outputs, workspace, stream);
}

}  // extern "C"
