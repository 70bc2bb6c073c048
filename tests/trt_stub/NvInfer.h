// Stand-in for TensorRT's NvInfer.h -- TEST INFRASTRUCTURE ONLY.
// TensorRT is not in this image, so the plugin shim (3m-asr-inference_b200/csrc/trt/b200moe_trt_plugins.cpp) cannot be
// compiled against the real header here.  This file declares just the part of the TensorRT 8 plugin API the shim uses
// (names and signatures as published in the TensorRT developer guide: IPluginV2DynamicExt, IPluginCreator,
// PluginTensorDesc, PluginFieldCollection, REGISTER_TENSORRT_PLUGIN), so that tests/test_trt_shim.py can compile the
// shim, link it with libb200moe.so and drive it through the same virtual calls a TensorRT builder / runtime makes.
// Where TensorRT exists, build the shim against the real header: make -C 3m-asr-inference_b200/csrc trt TRT_INCLUDE=...
#pragma once
#include <cuda_runtime_api.h>

#include <cstddef>
#include <cstdint>
#include <vector>

#define NV_TENSORRT_MAJOR 8
#define B200MOE_TRT_STUB 1

namespace nvinfer1 {

using AsciiChar = char;
enum class DataType : int32_t { kFLOAT = 0, kHALF = 1, kINT8 = 2, kINT32 = 3, kBOOL = 4 };
enum class TensorFormat : int32_t { kLINEAR = 0 };
enum class PluginFieldType : int32_t { kFLOAT16 = 0, kFLOAT32 = 1, kFLOAT64 = 2, kINT8 = 3, kINT16 = 4, kINT32 = 5,
                                       kCHAR = 6, kDIMS = 7, kUNKNOWN = 8 };

struct Dims {
  static constexpr int32_t MAX_DIMS = 8;
  int32_t nbDims;
  int32_t d[MAX_DIMS];
};

class IDimensionExpr {
 public:
  virtual bool isConstant() const noexcept = 0;
  virtual int32_t getConstantValue() const noexcept = 0;

 protected:
  virtual ~IDimensionExpr() noexcept = default;
};

struct DimsExprs {
  int32_t nbDims;
  const IDimensionExpr* d[Dims::MAX_DIMS];
};

class IExprBuilder {
 public:
  virtual const IDimensionExpr* constant(int32_t value) noexcept = 0;

 protected:
  virtual ~IExprBuilder() noexcept = default;
};

struct PluginTensorDesc {
  Dims dims;
  DataType type;
  TensorFormat format;
  float scale;
};

struct DynamicPluginTensorDesc {
  PluginTensorDesc desc;
  Dims min;
  Dims max;
};

struct PluginField {
  const AsciiChar* name;
  const void* data;
  PluginFieldType type;
  int32_t length;
};

struct PluginFieldCollection {
  int32_t nbFields;
  const PluginField* fields;
};

class IPluginV2 {
 public:
  virtual const AsciiChar* getPluginType() const noexcept = 0;
  virtual const AsciiChar* getPluginVersion() const noexcept = 0;
  virtual int32_t getNbOutputs() const noexcept = 0;
  virtual int32_t initialize() noexcept = 0;
  virtual void terminate() noexcept = 0;
  virtual size_t getSerializationSize() const noexcept = 0;
  virtual void serialize(void* buffer) const noexcept = 0;
  virtual void destroy() noexcept = 0;
  virtual void setPluginNamespace(const AsciiChar* pluginNamespace) noexcept = 0;
  virtual const AsciiChar* getPluginNamespace() const noexcept = 0;

 protected:
  virtual ~IPluginV2() noexcept = default;
};

class IPluginV2Ext : public IPluginV2 {
 public:
  virtual DataType getOutputDataType(int32_t index, const DataType* inputTypes, int32_t nbInputs) const noexcept = 0;
};

class IPluginV2DynamicExt : public IPluginV2Ext {
 public:
  virtual IPluginV2DynamicExt* clone() const noexcept = 0;
  virtual DimsExprs getOutputDimensions(int32_t outputIndex, const DimsExprs* inputs, int32_t nbInputs,
                                        IExprBuilder& exprBuilder) noexcept = 0;
  virtual bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* inOut, int32_t nbInputs,
                                         int32_t nbOutputs) noexcept = 0;
  virtual void configurePlugin(const DynamicPluginTensorDesc* in, int32_t nbInputs, const DynamicPluginTensorDesc* out,
                               int32_t nbOutputs) noexcept = 0;
  virtual size_t getWorkspaceSize(const PluginTensorDesc* inputs, int32_t nbInputs, const PluginTensorDesc* outputs,
                                  int32_t nbOutputs) const noexcept = 0;
  virtual int32_t enqueue(const PluginTensorDesc* inputDesc, const PluginTensorDesc* outputDesc,
                          const void* const* inputs, void* const* outputs, void* workspace,
                          cudaStream_t stream) noexcept = 0;
};

class IPluginCreator {
 public:
  virtual const AsciiChar* getPluginName() const noexcept = 0;
  virtual const AsciiChar* getPluginVersion() const noexcept = 0;
  virtual const PluginFieldCollection* getFieldNames() noexcept = 0;
  virtual IPluginV2* createPlugin(const AsciiChar* name, const PluginFieldCollection* fc) noexcept = 0;
  virtual IPluginV2* deserializePlugin(const AsciiChar* name, const void* serialData, size_t serialLength) noexcept = 0;
  virtual void setPluginNamespace(const AsciiChar* pluginNamespace) noexcept = 0;
  virtual const AsciiChar* getPluginNamespace() const noexcept = 0;
  virtual ~IPluginCreator() = default;
};

// the stub's plugin registry: REGISTER_TENSORRT_PLUGIN appends here (TensorRT: getPluginRegistry()->registerCreator)
inline std::vector<IPluginCreator*>& stubRegistry() {
  static std::vector<IPluginCreator*> r;
  return r;
}
template <typename T>
struct StubRegistrar {
  StubRegistrar() { stubRegistry().push_back(&instance); }
  T instance;
};

}  // namespace nvinfer1

#define REGISTER_TENSORRT_PLUGIN(name) static nvinfer1::StubRegistrar<name> pluginRegistrar##name {}
