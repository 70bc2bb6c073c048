// C-ABI layer: argument validation, workspace carving, stage sequencing. See include/b200moe.h for the contract and the
// reference interfaces each entry point replaces.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "common.cuh"

namespace b200moe {

namespace {
thread_local std::string g_last_error;
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(B200MOE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

bool dtype_ok(int d) { return d == B200MOE_F32 || d == B200MOE_F16 || d == B200MOE_BF16; }


__global__ void __launch_bounds__(256) to_f32_kernel(const void* src, int dtype, float* dst, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v;
    if (dtype == B200MOE_F32)
      v = static_cast<const float*>(src)[i];
    else if (dtype == B200MOE_F16)
      v = __half2float(static_cast<const __half*>(src)[i]);
    else
      v = __bfloat162float(static_cast<const bf16*>(src)[i]);
    dst[i] = v;
  }
}

// ---- optional per-stage timing with CUDA events (eager mode only; used by bench.py for the roofline numbers) ----------
constexpr int kStages = 4;  // 0 gate, 1 dispatch, 2 expert_ffn (+ fused combine), 3 combine
constexpr int kMaxRecords = 8192;
struct StageTimer {
  bool enabled = false;
  cudaEvent_t ev[kMaxRecords][2];
  int stage[kMaxRecords];
  int created = 0;
  int used = 0;
};
StageTimer g_timer;

struct StageScope {
  int slot = -1;
  cudaStream_t stream;
  StageScope(int stage, cudaStream_t s) : stream(s) {
    if (!g_timer.enabled || g_timer.used >= kMaxRecords) return;
    slot = g_timer.used++;
    if (slot >= g_timer.created) {
      cudaEventCreate(&g_timer.ev[slot][0]);
      cudaEventCreate(&g_timer.ev[slot][1]);
      g_timer.created = slot + 1;
    }
    g_timer.stage[slot] = stage;
    cudaEventRecord(g_timer.ev[slot][0], stream);
  }
  ~StageScope() {
    if (slot >= 0) cudaEventRecord(g_timer.ev[slot][1], stream);
  }
};

}  // namespace

void count_launch(int n) { g_launches.fetch_add(static_cast<unsigned long long>(n), std::memory_order_relaxed); }

namespace {
int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}
}  // namespace

// Tunables: environment default, overridable at run time through b200moe_config() (tests flip them per case).
namespace {
std::atomic<int> g_pdl{-1}, g_pdl_trig{-1}, g_prefetch{-1}, g_route{-1}, g_ln_fuse{-1};
int knob(std::atomic<int>& v, const char* env, int dflt) {
  int cur = v.load(std::memory_order_relaxed);
  if (cur < 0) {
    cur = env_int(env, dflt);
    if (cur < 0) cur = dflt;
    v.store(cur, std::memory_order_relaxed);
  }
  return cur;
}
}  // namespace

// Defaults: the gate / route kernel and the expert-FFN kernel are launched with the PDL attribute and the FFN kernel
// releases its dependents at its start, so that the next layer's route kernel does the embed half of the router GEMM
// (constants only) while this layer's FFN kernel drains: measured 32.7 -> 29.8 us per layer on cfg3 under CUDA graphs.
int pdl_mask() { return knob(g_pdl, "B200MOE_PDL", kPdlGate | kPdlFfn | kPdlLn); }
int pdl_trigger() { return knob(g_pdl_trig, "B200MOE_PDL_TRIG", kPdlFfn); }
int route_mode() { return knob(g_route, "B200MOE_ROUTE", 1); }
// 1: when the caller supplies the pre-scaled router (b200moe_block_args.Wr_packed_ln) the block's norm_ff is folded into
// the route kernel; 0: always the row pass in front.
int ln_fuse_mode() { return knob(g_ln_fuse, "B200MOE_LN_FUSE", 1); }

int prefetch_mode() { return knob(g_prefetch, "B200MOE_PREFETCH", 0); }

}  // namespace b200moe

using namespace b200moe;

struct b200moe_plugin {
  int data_type;
  int num_expert;
  int idim;
  int hidden_units;
  int act_type;
  // bf16 / fp32 copies of the weights the plugin OWNS (cudaMalloc on the first enqueue, freed by destroy): the scratch
  // workspace TensorRT hands to enqueue is shared between layers and not preserved between calls, so nothing cached may
  // live there.  packed_src = the four input pointers the copies were made from.
  const void* packed_src[4];
  void* packed;   // [w1 bf16 | w2 bf16 | b1 fp32 | b2 fp32]; w1 / w2 parts absent when data_type is already bf16
  int device;
};

extern "C" {

const char* b200moe_last_error(void) { return g_last_error.c_str(); }

int b200moe_version(void) { return B200MOE_VERSION; }

unsigned long long b200moe_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int b200moe_device_supported(int dev) {
  int major = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
  return major == 10 ? 1 : 0;
}

int b200moe_status(int* host_status, int clear) {
  if (!host_status) return fail(B200MOE_ERR_ARG, "status: null pointer");
  cudaError_t e = read_route_status(host_status, clear != 0);
  if (e != cudaSuccess) return cuda_fail(e, "status");
  return B200MOE_OK;
}

int b200moe_debug_route_trace(void* dev_buf) {
  set_route_trace(dev_buf);
  return B200MOE_OK;
}

int b200moe_debug_timeline(void* dev_buf, int max_launches) {
  b200moe::set_timeline(dev_buf, max_launches);
  return B200MOE_OK;
}

int b200moe_debug_timeline_kind(int slot) { return b200moe::timeline_kind(slot); }

int b200moe_debug_ffn_trace(void* dev_buf, int records_per_cta) {
  set_ffn_trace(dev_buf, records_per_cta);
  return B200MOE_OK;
}

int b200moe_config(const char* key, int value) {
  if (!key || value < 0) return fail(B200MOE_ERR_ARG, "config: bad argument");
  const std::string k(key);
  if (k == "route") g_route.store(value);
  else if (k == "pdl") g_pdl.store(value);
  else if (k == "pdl_trig") g_pdl_trig.store(value);
  else if (k == "prefetch") g_prefetch.store(value);
  else if (k == "ln_fuse") g_ln_fuse.store(value);
  else return fail(B200MOE_ERR_ARG, "config: unknown key '%s'", key);
  return B200MOE_OK;
}

int b200moe_profile_enable(int on) {
  g_timer.enabled = on != 0;
  g_timer.used = 0;
  return B200MOE_OK;
}

int b200moe_profile_read(float* stage_ms, int* stage_calls) {
  if (!stage_ms || !stage_calls) return fail(B200MOE_ERR_ARG, "profile_read: null pointer");
  for (int i = 0; i < kStages; ++i) {
    stage_ms[i] = 0.0f;
    stage_calls[i] = 0;
  }
  for (int r = 0; r < g_timer.used; ++r) {
    cudaError_t e = cudaEventSynchronize(g_timer.ev[r][1]);
    if (e != cudaSuccess) return cuda_fail(e, "profile_read");
    float ms = 0.0f;
    e = cudaEventElapsedTime(&ms, g_timer.ev[r][0], g_timer.ev[r][1]);
    if (e != cudaSuccess) return cuda_fail(e, "profile_read");
    stage_ms[g_timer.stage[r]] += ms;
    stage_calls[g_timer.stage[r]] += 1;
  }
  g_timer.used = 0;
  return B200MOE_OK;
}

int b200moe_pack_bf16(const void* src, int src_dtype, void* dst_bf16, size_t n, cudaStream_t stream) {
  if ((!src || !dst_bf16) && n > 0) return fail(B200MOE_ERR_ARG, "pack_bf16: null pointer");
  if (!dtype_ok(src_dtype)) return fail(B200MOE_ERR_ARG, "pack_bf16: bad dtype %d", src_dtype);
  cudaError_t e = launch_pack_bf16(src, src_dtype, static_cast<bf16*>(dst_bf16), n, stream);
  if (e != cudaSuccess) return cuda_fail(e, "pack_bf16");
  return B200MOE_OK;
}

int b200moe_pack_tf32(const float* src, float* dst, size_t n, cudaStream_t stream) {
  if ((!src || !dst) && n > 0) return fail(B200MOE_ERR_ARG, "pack_tf32: null pointer");
  cudaError_t e = launch_pack_tf32(src, dst, n, stream);
  if (e != cudaSuccess) return cuda_fail(e, "pack_tf32");
  return B200MOE_OK;
}

size_t b200moe_workspace_bytes(int S, int E, int D, int H, int top_k) {
  if (S < 0 || E < 1 || D < 1 || H < 0 || top_k < 1) return 0;
  return carve_workspace(nullptr, S, E, D, H, top_k).bytes;
}

int b200moe_gate(const void* x, const void* embed, const float* Wr, const float* br, const int* x_len, int B, int T,
                 int D, int Demb, int E, int top_k, int gate_mode, int dtype, int* idx, float* score,
                 cudaStream_t stream) {
  if (B < 0 || T < 0 || D < 1) return fail(B200MOE_ERR_ARG, "gate: bad shape B=%d T=%d D=%d", B, T, D);
  if (B * T > 0 && (!x || !Wr || !idx || !score)) return fail(B200MOE_ERR_ARG, "gate: null pointer");
  if (!dtype_ok(dtype)) return fail(B200MOE_ERR_ARG, "gate: bad dtype %d", dtype);
  if (gate_mode != B200MOE_GATE_3M && gate_mode != B200MOE_GATE_NAIVE)
    return fail(B200MOE_ERR_ARG, "gate: bad gate_mode %d", gate_mode);
  if (E < 1 || E > kMaxExperts) return fail(B200MOE_ERR_ARG, "gate: E=%d outside [1, %d]", E, kMaxExperts);
  if (top_k < 1 || top_k > 8 || top_k > E) return fail(B200MOE_ERR_ARG, "gate: top_k=%d unsupported", top_k);
  if (gate_mode == B200MOE_GATE_3M && top_k != 1)
    return fail(B200MOE_ERR_ARG, "gate: the 3M router is top-1 (got top_k=%d)", top_k);
  cudaError_t e = launch_gate(x, embed, Wr, br, x_len, B, T, D, embed ? Demb : 0, E, top_k, gate_mode, dtype, idx,
                              score, stream);
  if (e != cudaSuccess) return cuda_fail(e, "gate");
  return B200MOE_OK;
}

size_t b200moe_router_pack_bytes(int R) { return R > 0 ? router_pack_bytes(R) : 0; }

size_t b200moe_router_ln_pack_bytes(int R) { return R > 0 ? router_ln_pack_bytes(R) : 0; }

int b200moe_pack_router_ln(const float* Wr, int R, int E, int D, const float* gamma, const float* beta, void* packed,
                           cudaStream_t stream) {
  if (!Wr || !packed || !gamma || !beta) return fail(B200MOE_ERR_ARG, "pack_router_ln: null pointer");
  if (E < 1 || E > 32 || R < 1 || D < 1 || D > R)
    return fail(B200MOE_ERR_ARG, "pack_router_ln: needs 1 <= E <= 32 and 1 <= D <= R (got E=%d, R=%d, D=%d)", E, R, D);
  cudaError_t e = launch_pack_router_ln(Wr, R, E, D, gamma, beta, packed, stream);
  if (e != cudaSuccess) return cuda_fail(e, "pack_router_ln");
  return B200MOE_OK;
}

int b200moe_pack_router(const float* Wr, int R, int E, void* packed, cudaStream_t stream) {
  if (!Wr || !packed) return fail(B200MOE_ERR_ARG, "pack_router: null pointer");
  if (E < 1 || E > 32 || R < 1) return fail(B200MOE_ERR_ARG, "pack_router: needs 1 <= E <= 32 (got E=%d, R=%d)", E, R);
  cudaError_t e = launch_pack_router(Wr, R, E, packed, stream);
  if (e != cudaSuccess) return cuda_fail(e, "pack_router");
  return B200MOE_OK;
}

int b200moe_gate_tc(const void* x, const void* embed, const void* Wr_packed, const float* br, const int* x_len, int B,
                    int T, int D, int Demb, int E, int top_k, int gate_mode, int* idx, float* score,
                    cudaStream_t stream) {
  if (B < 0 || T < 0 || D < 1) return fail(B200MOE_ERR_ARG, "gate_tc: bad shape");
  if (B * T > 0 && (!x || !Wr_packed || !idx || !score)) return fail(B200MOE_ERR_ARG, "gate_tc: null pointer");
  if (gate_mode != B200MOE_GATE_3M && gate_mode != B200MOE_GATE_NAIVE)
    return fail(B200MOE_ERR_ARG, "gate_tc: bad gate_mode %d", gate_mode);
  if (top_k < 1 || top_k > 8 || top_k > E) return fail(B200MOE_ERR_ARG, "gate_tc: top_k=%d unsupported", top_k);
  if (gate_mode == B200MOE_GATE_3M && top_k != 1) return fail(B200MOE_ERR_ARG, "gate_tc: the 3M router is top-1");
  if (!gate_tc_supported(D, embed ? Demb : 0, E, top_k, B200MOE_BF16))
    return fail(B200MOE_ERR_ARG, "gate_tc: needs bf16 activations, E <= 32, D and Demb multiples of 64");
  cudaError_t e = launch_gate_tc(x, embed, Wr_packed, br, x_len, B, T, D, embed ? Demb : 0, E, top_k, gate_mode, idx,
                                 score, nullptr, nullptr, 0, nullptr, 0, stream);
  if (e != cudaSuccess) return cuda_fail(e, "gate_tc");
  return B200MOE_OK;
}

int b200moe_softmax_topk_enqueue(const void* logits, const int* mask, int B, int T, int E, int data_type, void* value,
                                 int* idx, cudaStream_t stream) {
  if (B * T > 0 && (!logits || !value || !idx)) return fail(B200MOE_ERR_ARG, "softmax_topk: null pointer");
  if (!dtype_ok(data_type)) return fail(B200MOE_ERR_ARG, "softmax_topk: bad data_type %d", data_type);
  // the reference kernel handles dim <= 128 and returns -1 beyond that (softmax_topk_kernel.cu:96-116)
  if (E < 1 || E > kMaxExperts) return fail(B200MOE_ERR_ARG, "softmax_topk: E=%d outside [1, %d]", E, kMaxExperts);
  cudaError_t e = launch_softmax_topk(logits, mask, B, T, E, data_type, value, idx, stream);
  if (e != cudaSuccess) return cuda_fail(e, "softmax_topk");
  return B200MOE_OK;
}

int b200moe_dispatch(const void* x, const int* idx, int S, int D, int E, int top_k, int dtype, int* counts,
                     int* offsets, int* mapping, void* xbuf, void* ws, cudaStream_t stream) {
  if (S < 0 || D < 1 || top_k < 1) return fail(B200MOE_ERR_ARG, "dispatch: bad shape");
  if (E < 1 || E > kMaxExperts) return fail(B200MOE_ERR_ARG, "dispatch: E=%d outside [1, %d]", E, kMaxExperts);
  if (D % 8 != 0) return fail(B200MOE_ERR_ARG, "dispatch: D=%d must be a multiple of 8", D);
  if (!dtype_ok(dtype)) return fail(B200MOE_ERR_ARG, "dispatch: bad dtype %d", dtype);
  if (!ws || (S > 0 && (!x || !idx || !xbuf))) return fail(B200MOE_ERR_ARG, "dispatch: null pointer");
  RouteWs w = carve_workspace(ws, S, E, D, 0, top_k);
  cudaError_t e = launch_dispatch(x, idx, nullptr, S, D, E, top_k, dtype, choose_bn(S * top_k, E), w, counts, offsets,
                                  mapping, static_cast<bf16*>(xbuf), nullptr, nullptr, nullptr, stream);
  if (e != cudaSuccess) return cuda_fail(e, "dispatch");
  return B200MOE_OK;
}

int b200moe_prepare(const int* idx, int S, int E, int top_k, int* counts, int* offsets, int* mapping, int* pos, void* ws,
                    cudaStream_t stream) {
  if (S < 0 || top_k < 1) return fail(B200MOE_ERR_ARG, "prepare: bad shape");
  if (E < 1 || E > kMaxExperts) return fail(B200MOE_ERR_ARG, "prepare: E=%d outside [1, %d]", E, kMaxExperts);
  if (!ws || (S > 0 && !idx)) return fail(B200MOE_ERR_ARG, "prepare: null pointer");
  RouteWs w = carve_workspace(ws, S, E, 8, 0, top_k);
  cudaError_t e = launch_dispatch(nullptr, idx, nullptr, S, 8, E, top_k, B200MOE_BF16, choose_bn(S * top_k, E), w, counts,
                                  offsets, mapping, nullptr, nullptr, nullptr, nullptr, stream);
  if (e == cudaSuccess && pos != nullptr && S > 0)
    e = cudaMemcpyAsync(pos, w.pos, sizeof(int) * static_cast<size_t>(S) * top_k, cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) return cuda_fail(e, "prepare");
  return B200MOE_OK;
}

int b200moe_scatter_rows(const void* in, const int* index, int n, int n_out, int D, int dtype, void* out,
                         cudaStream_t stream) {
  if (n < 0 || n_out < 0 || D < 1 || !dtype_ok(dtype)) return fail(B200MOE_ERR_ARG, "scatter_rows: bad argument");
  const int row_bytes = D * (dtype == B200MOE_F32 ? 4 : 2);
  if (row_bytes % 16 != 0) return fail(B200MOE_ERR_ARG, "scatter_rows: rows of %d bytes are not a multiple of 16", row_bytes);
  if (n > 0 && (!in || !index || !out)) return fail(B200MOE_ERR_ARG, "scatter_rows: null pointer");
  cudaError_t e = launch_scatter_rows(in, index, n, n_out, row_bytes, out, stream);
  if (e != cudaSuccess) return cuda_fail(e, "scatter_rows");
  return B200MOE_OK;
}

int b200moe_expert_linear(const void* xbuf, const int* offsets, int n_rows, const void* W, const float* bias, int E,
                          int K, int N, int act_type, void* out_bf16, void* ws, cudaStream_t stream) {
  if (n_rows < 0) return fail(B200MOE_ERR_ARG, "expert_linear: n_rows < 0");
  if (E < 1 || E > kMaxExperts) return fail(B200MOE_ERR_ARG, "expert_linear: E=%d outside [1, %d]", E, kMaxExperts);
  if (K % 128 != 0 || N % 128 != 0)
    return fail(B200MOE_ERR_ARG, "expert_linear: in=%d and out=%d features must be multiples of 128", K, N);
  if (act_type < 0 || act_type > 3) return fail(B200MOE_ERR_ARG, "expert_linear: bad act_type %d (0 silu, 1 relu, 2 gelu, 3 none)", act_type);
  if (!ws || !offsets || !W || (n_rows > 0 && (!xbuf || !out_bf16))) return fail(B200MOE_ERR_ARG, "expert_linear: null pointer");
  if (n_rows == 0) return B200MOE_OK;
  RouteWs w = carve_workspace(ws, n_rows, E, K, 0, 1);
  const int bn = choose_bn(n_rows, E);
  const int gmax = max_groups(n_rows, E, bn);
  cudaError_t e = launch_build_groups(offsets, E, bn, w.groups, w.n_groups, w.h_ready, gmax, stream);
  if (e != cudaSuccess) return cuda_fail(e, "expert_linear/build_groups");
  FfnLaunch a{};
  a.xbuf = static_cast<const bf16*>(xbuf);
  a.hbuf = static_cast<bf16*>(out_bf16);   // the first GEMM's output IS the result
  a.W1 = static_cast<const bf16*>(W);
  a.b1 = bias;
  a.groups = w.groups;
  a.n_groups = w.n_groups;
  a.h_ready = w.h_ready;
  a.n_rows = n_rows;
  a.E = E;
  a.D = K;
  a.H = N;
  a.bn = bn;
  a.act = act_type;
  a.gmax = gmax;
  a.out_dtype = B200MOE_BF16;
  a.top_k = 1;
  a.ff_scale = 1.0f;
  a.p1_only = 1;
  e = launch_ffn(a, stream);
  if (e != cudaSuccess) return cuda_fail(e, "expert_linear");
  return B200MOE_OK;
}

int b200moe_expert_ffn(const void* xbuf, const int* offsets, int n_rows, const void* W1, const float* b1,
                       const void* W2, const float* b2, int E, int D, int H, int act_type, int out_dtype, void* ybuf,
                       void* ws, cudaStream_t stream) {
  if (n_rows < 0) return fail(B200MOE_ERR_ARG, "expert_ffn: n_rows < 0");
  if (E < 1 || E > kMaxExperts) return fail(B200MOE_ERR_ARG, "expert_ffn: E=%d outside [1, %d]", E, kMaxExperts);
  if (D % 128 != 0 || H % 128 != 0)
    return fail(B200MOE_ERR_ARG, "expert_ffn: D=%d and H=%d must be multiples of 128", D, H);
  if (act_type < 0 || act_type > 2) return fail(B200MOE_ERR_ARG, "expert_ffn: bad act_type %d", act_type);
  if (!dtype_ok(out_dtype)) return fail(B200MOE_ERR_ARG, "expert_ffn: bad out_dtype %d", out_dtype);
  if (!ws || !offsets || !W1 || !W2 || (n_rows > 0 && (!xbuf || !ybuf)))
    return fail(B200MOE_ERR_ARG, "expert_ffn: null pointer");
  if (n_rows == 0) return B200MOE_OK;
  RouteWs w = carve_workspace(ws, n_rows, E, D, H, 1);
  const int bn = choose_bn(n_rows, E);
  const int gmax = max_groups(n_rows, E, bn);
  cudaError_t e = launch_build_groups(offsets, E, bn, w.groups, w.n_groups, w.h_ready, gmax, stream);
  if (e != cudaSuccess) return cuda_fail(e, "expert_ffn/build_groups");
  FfnLaunch a{};
  a.xbuf = static_cast<const bf16*>(xbuf);
  a.hbuf = static_cast<bf16*>(w.hbuf);
  a.W1 = static_cast<const bf16*>(W1);
  a.W2 = static_cast<const bf16*>(W2);
  a.b1 = b1;
  a.b2 = b2;
  a.groups = w.groups;
  a.n_groups = w.n_groups;
  a.h_ready = w.h_ready;
  a.n_rows = n_rows;
  a.E = E;
  a.D = D;
  a.H = H;
  a.bn = bn;
  a.act = act_type;
  a.gmax = gmax;
  a.fused = 0;
  a.out_dtype = out_dtype;
  a.out = ybuf;
  a.top_k = 1;
  a.ff_scale = 1.0f;
  e = launch_ffn(a, stream);
  if (e != cudaSuccess) return cuda_fail(e, "expert_ffn");
  return B200MOE_OK;
}

int b200moe_combine(const void* ybuf, const int* mapping, const float* score, const void* residual, float ff_scale,
                    int S, int D, int top_k, int dtype, void* out, cudaStream_t stream) {
  if (S < 0 || D < 1 || top_k < 1) return fail(B200MOE_ERR_ARG, "combine: bad shape");
  if (D % 8 != 0) return fail(B200MOE_ERR_ARG, "combine: D=%d must be a multiple of 8", D);
  if (!dtype_ok(dtype)) return fail(B200MOE_ERR_ARG, "combine: bad dtype %d", dtype);
  if (S > 0 && (!ybuf || !mapping || !out)) return fail(B200MOE_ERR_ARG, "combine: null pointer");
  cudaError_t e = launch_combine(ybuf, mapping, score, residual, ff_scale, S, D, top_k, dtype, out, stream);
  if (e != cudaSuccess) return cuda_fail(e, "combine");
  return B200MOE_OK;
}

namespace {
// norm_ff fused into the route kernel (block call); null gamma = the layer alone
struct LnFuse {
  const float* gamma;
  const float* beta;
  float eps;
  const void* packed;  // norm_ff only: the pre-scaled router of b200moe_pack_router_ln (replaces Wr_packed)
};

const float* ln_consts(const LnFuse* ln, int R) {
  return reinterpret_cast<const float*>(static_cast<const uint8_t*>(ln->packed) + static_cast<size_t>(64) * R * sizeof(bf16));
}

bool takes_route_kernel(const b200moe_layer_args* a) {
  const int Demb = a->embed ? a->Demb : 0;
  return a->Wr_packed != nullptr && gate_tc_supported(a->D, Demb, a->E, a->top_k, a->dtype) &&
         route_supported(a->B * a->T, a->D, Demb, a->E, a->top_k, a->dtype);
}
}  // namespace

static int forward_impl(const b200moe_layer_args* a, void* ws, size_t ws_bytes, cudaStream_t stream, const LnFuse* ln,
                        const LnFuse* ln_out);

int b200moe_forward(const b200moe_layer_args* a, void* ws, size_t ws_bytes, cudaStream_t stream) {
  return forward_impl(a, ws, ws_bytes, stream, nullptr, nullptr);
}

// ln: norm_ff fused into the route kernel (the caller has checked takes_route_kernel); ln_out: norm_final behind the
// residual add -- inside the combine kernel for top-k > 1, as a row pass over `out` behind the fused top-1 epilogue.
// Measured and dropped, both for norm_final inside the expert kernel (cfg3, us per block, same box as the row pass):
//   per group -- every second-GEMM tile bumps its group's flag, waits for the D / 128 sibling tiles, normalises its
//     share of the rows from L2: 45.0 vs 39.8 (sibling skew; the extra live state spilled registers in the epilogue and
//     cost the bare layer 1.4 us);
//   at the kernel's end behind a grid-wide barrier, all warps normalising from L2: 42.2 vs 41.2-42.7.  The row pass is
//     not the cost; what any of these lose is the overlap of the next layer's route prologue with this kernel's tail.
static int forward_impl(const b200moe_layer_args* a, void* ws, size_t ws_bytes, cudaStream_t stream, const LnFuse* ln,
                        const LnFuse* ln_out) {
  if (!a) return fail(B200MOE_ERR_ARG, "forward: null args");
  const int S = a->B * a->T;
  if (a->B < 0 || a->T < 0) return fail(B200MOE_ERR_ARG, "forward: bad B/T");
  if (a->E < 1 || a->E > kMaxExperts) return fail(B200MOE_ERR_ARG, "forward: E=%d outside [1, %d]", a->E, kMaxExperts);
  if (a->D % 128 != 0 || a->H % 128 != 0)
    return fail(B200MOE_ERR_ARG, "forward: D=%d and H=%d must be multiples of 128", a->D, a->H);
  if (!dtype_ok(a->dtype)) return fail(B200MOE_ERR_ARG, "forward: bad dtype %d", a->dtype);
  if (a->top_k < 1 || a->top_k > 8 || a->top_k > a->E) return fail(B200MOE_ERR_ARG, "forward: bad top_k %d", a->top_k);
  if (a->gate_mode == B200MOE_GATE_3M && a->top_k != 1)
    return fail(B200MOE_ERR_ARG, "forward: the 3M router is top-1 (got top_k=%d)", a->top_k);
  if (a->act_type < 0 || a->act_type > 2) return fail(B200MOE_ERR_ARG, "forward: bad act_type %d", a->act_type);
  if (S == 0) return B200MOE_OK;
  if (!a->x || !a->out || (!a->Wr && !a->Wr_packed) || !a->W1 || !a->W2 || !ws)
    return fail(B200MOE_ERR_ARG, "forward: null pointer");
  if (!a->Wr && !gate_tc_supported(a->D, a->embed ? a->Demb : 0, a->E, a->top_k, a->dtype))
    return fail(B200MOE_ERR_ARG, "forward: this shape/dtype needs the fp32 router (Wr), the packed one is not enough");
  const int Demb = a->embed ? a->Demb : 0;
  RouteWs w = carve_workspace(ws, S, a->E, a->D, a->H, a->top_k);
  if (ws_bytes < w.bytes)
    return fail(B200MOE_ERR_WORKSPACE, "forward: workspace %zu B < required %zu B", ws_bytes, w.bytes);
  const int Sk = S * a->top_k;
  int* idx = a->idx_out ? a->idx_out : w.idx;
  float* score = a->score_out ? a->score_out : w.score;

  cudaError_t e;
  const bool tf32 = a->compute == B200MOE_COMPUTE_TF32;
  if (a->compute != B200MOE_COMPUTE_BF16 && !tf32) return fail(B200MOE_ERR_ARG, "forward: bad compute mode %d", a->compute);
  if (tf32 && a->dtype != B200MOE_F32)
    return fail(B200MOE_ERR_ARG, "forward: TF32 compute takes fp32 activations and fp32 weights");
  if (tf32 && !a->Wr) return fail(B200MOE_ERR_ARG, "forward: TF32 compute needs the fp32 router (Wr)");
  const bool tc_gate = a->Wr_packed != nullptr && gate_tc_supported(a->D, Demb, a->E, a->top_k, a->dtype);
  const int bn = choose_bn(Sk, a->E);
  const int gmax = max_groups(Sk, a->E, bn);
  const bool fused = a->top_k == 1;
  const bool route = tc_gate && route_supported(S, a->D, Demb, a->E, a->top_k, a->dtype);
  if (route) {
    // small batch: gate and dispatch as one kernel (grid barrier instead of a kernel boundary)
    StageScope t(0, stream);
    e = launch_route(a->x, a->embed, ln ? ln->packed : a->Wr_packed, a->br, a->x_len, a->B, a->T, a->D, Demb, a->E,
                     a->gate_mode, a->keep_expert_output, idx, score, bn, w, a->counts_out, nullptr, a->mapping_out,
                     w.xbuf, fused ? a->out : nullptr, a->residual, stream, nullptr, false, ln ? ln->gamma : nullptr,
                     ln ? ln->beta : nullptr, ln ? ln->eps : 0.0f, ln ? ln_consts(ln, a->D + Demb) : nullptr);
    if (e != cudaSuccess) return cuda_fail(e, "forward/route");
  } else {
  {
    StageScope t(0, stream);
    if (tc_gate) {
      // small batches are bound by streaming the 2 * E * D * H weights: start pulling them into L2 now
      const size_t wbytes = static_cast<size_t>(a->E) * a->H * a->D * sizeof(bf16);
      const bool pf = Sk <= 32768;
      e = launch_gate_tc(a->x, a->embed, a->Wr_packed, a->br, a->x_len, a->B, a->T, a->D, Demb, a->E, a->top_k,
                         a->gate_mode, idx, score, w.hist32, pf ? a->W1 : nullptr, wbytes, pf ? a->W2 : nullptr, wbytes,
                         stream);
    }
    else
      e = launch_gate(a->x, a->embed, a->Wr, a->br, a->x_len, a->B, a->T, a->D, Demb, a->E, a->top_k, a->gate_mode,
                      a->dtype, idx, score, stream);
  }
  if (e != cudaSuccess) return cuda_fail(e, "forward/gate");
  {
    StageScope t(1, stream);
    e = launch_dispatch(a->x, idx, a->keep_expert_output ? nullptr : score, S, a->D, a->E, a->top_k, a->dtype, bn, w,
                        a->counts_out, nullptr, a->mapping_out, w.xbuf, fused ? a->out : nullptr, a->residual,
                        tc_gate ? w.hist32 : nullptr, stream, nullptr, false, tf32);
  }
  if (e != cudaSuccess) return cuda_fail(e, "forward/dispatch");
  }

  FfnLaunch f{};
  f.xbuf = w.xbuf;
  f.hbuf = static_cast<bf16*>(w.hbuf);
  f.W1 = static_cast<const bf16*>(a->W1);
  f.W2 = static_cast<const bf16*>(a->W2);
  f.b1 = a->b1;
  f.b2 = a->b2;
  f.groups = w.groups;
  f.n_groups = w.n_groups;
  f.h_ready = w.h_ready;
  f.n_rows = Sk;
  f.E = a->E;
  f.D = a->D;
  f.H = a->H;
  f.bn = bn;
  f.act = a->act_type;
  f.gmax = gmax;
  f.out_dtype = a->dtype;
  f.top_k = a->top_k;
  f.ff_scale = a->ff_scale;
  f.tf32 = tf32 ? 1 : 0;
  if (route) {
    f.clear_ptr = w.hist32;
    f.clear_ints = 2 * ((S + 31) / 32) * a->E;  // 64-bit words
  }
  if (fused) {
    f.fused = 1;
    f.out = a->out;
    f.residual = a->residual;
    f.pos = w.pos;
    f.row_score = w.row_score;
  } else {
    f.fused = 0;
    f.out = w.ybuf;
  }
  {
    StageScope t(2, stream);
    e = launch_ffn(f, stream);
  }
  if (e != cudaSuccess) return cuda_fail(e, "forward/expert_ffn");
  if (!fused) {
    StageScope t(3, stream);
    e = launch_combine(w.ybuf, w.mapping, a->keep_expert_output ? nullptr : score, a->residual, a->ff_scale, S, a->D,
                       a->top_k, a->dtype, a->out, stream, ln_out ? ln_out->gamma : nullptr,
                       ln_out ? ln_out->beta : nullptr, ln_out ? ln_out->eps : 0.0f);
    if (e != cudaSuccess) return cuda_fail(e, "forward/combine");
  } else if (ln_out) {
    StageScope t(3, stream);
    e = launch_layernorm(a->out, ln_out->gamma, ln_out->beta, ln_out->eps, S, a->D, a->dtype, a->out, stream);
    if (e != cudaSuccess) return cuda_fail(e, "forward/norm_final");
  }
  return B200MOE_OK;
}

// ---- expert parallelism over peer-mapped memory ------------------------------------------------------------------

struct b200moe_ep_ctx {
  EpPeers peers;
  // folded path: the owners are still writing this rank's `out` rows of the last layer call; whatever touches `out`
  // next has to be ordered behind launch_ep_wait_done (the next b200moe_ep_forward* does it itself)
  bool pending = false;
  void* pending_out = nullptr;
  int pending_S = 0;
};

size_t b200moe_ep_buffer_bytes(int world, int E_local, int D, int cap) {
  if (world < 1 || world > kMaxEpWorld || E_local < 1 || D < 1 || cap < 1) return 0;
  return ep_layout(world, E_local, D, cap).bytes;
}

int b200moe_ep_alloc(size_t bytes, void** dev_ptr) {
  if (!dev_ptr || bytes == 0) return fail(B200MOE_ERR_ARG, "ep_alloc: bad argument");
  // cudaMalloc (not a pool or VMM allocation) so that the buffer can be exported with cudaIpcGetMemHandle
  cudaError_t e = cudaMalloc(dev_ptr, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "ep_alloc/cudaMalloc");
  e = cudaMemset(*dev_ptr, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return cuda_fail(e, "ep_alloc/cudaMemset");
  return B200MOE_OK;
}

int b200moe_ep_free(void* dev_ptr) {
  cudaError_t e = cudaFree(dev_ptr);
  if (e != cudaSuccess) return cuda_fail(e, "ep_free");
  return B200MOE_OK;
}

int b200moe_ep_ipc_export(void* dev_ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  if (!dev_ptr || !handle64) return fail(B200MOE_ERR_ARG, "ep_ipc_export: null pointer");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, dev_ptr);
  if (e != cudaSuccess) return cuda_fail(e, "ep_ipc_export");
  std::memcpy(handle64, &h, sizeof(h));
  return B200MOE_OK;
}

int b200moe_ep_ipc_open(const void* handle64, void** peer_ptr) {
  if (!handle64 || !peer_ptr) return fail(B200MOE_ERR_ARG, "ep_ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return cuda_fail(e, "ep_ipc_open");
  return B200MOE_OK;
}

int b200moe_ep_ipc_close(void* peer_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(peer_ptr);
  if (e != cudaSuccess) return cuda_fail(e, "ep_ipc_close");
  return B200MOE_OK;
}

b200moe_ep_ctx* b200moe_ep_create(int rank, int world, int E_local, int D, int cap, void* const* bufs,
                                  int timeout_ms) {
  if (world < 1 || world > kMaxEpWorld || rank < 0 || rank >= world || E_local < 1 ||
      E_local * world > kMaxExperts || D < 8 || D % 8 != 0 || cap < 1 || !bufs) {
    fail(B200MOE_ERR_ARG, "ep_create: bad argument (world <= %d, E_local * world <= %d)", kMaxEpWorld, kMaxExperts);
    return nullptr;
  }
  for (int r = 0; r < world; ++r)
    if (!bufs[r]) {
      fail(B200MOE_ERR_ARG, "ep_create: null buffer for rank %d", r);
      return nullptr;
    }
  b200moe_ep_ctx* c = new (std::nothrow) b200moe_ep_ctx();
  if (!c) return nullptr;
  std::memset(&c->peers, 0, sizeof(c->peers));
  c->peers.rank = rank;
  c->peers.world = world;
  c->peers.E_local = E_local;
  c->peers.cap = cap;
  c->peers.D = D;
  c->peers.timeout_ms = timeout_ms > 0 ? timeout_ms : 2000;
  c->peers.lay = ep_layout(world, E_local, D, cap);
  for (int r = 0; r < world; ++r) c->peers.base[r] = static_cast<uint8_t*>(bufs[r]);
  return c;
}

void b200moe_ep_destroy(b200moe_ep_ctx* c) { delete c; }

size_t b200moe_ep_workspace_bytes(const b200moe_ep_ctx* c, int H) {
  if (!c || H < 0) return 0;
  const EpPeers& p = c->peers;
  return carve_workspace(nullptr, p.world * p.cap, p.E_local * p.world, p.D, H, 1).bytes;
}

int b200moe_ep_status(const b200moe_ep_ctx* c, int* host_status) {
  if (!c || !host_status) return fail(B200MOE_ERR_ARG, "ep_status: null pointer");
  const EpPeers& p = c->peers;
  int ctrl[4];
  cudaError_t e = cudaMemcpy(ctrl, p.base[p.rank] + p.lay.ctrl, sizeof(ctrl), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return cuda_fail(e, "ep_status");
  *host_status = ctrl[3];
  return B200MOE_OK;
}

namespace {
constexpr int kEpAllStages = 15;
constexpr int kEpFlagDeferWait = 1;

int ep_flush_pending(b200moe_ep_ctx* c, cudaStream_t stream) {
  if (!c->pending) return B200MOE_OK;
  cudaError_t e = launch_ep_wait_done(c->peers, c->pending_out, c->pending_S, c->peers.D, stream);
  c->pending = false;
  c->pending_out = nullptr;
  if (e != cudaSuccess) return cuda_fail(e, "ep_wait");
  return B200MOE_OK;
}
}  // namespace

static int ep_forward_impl(b200moe_ep_ctx* c, const b200moe_layer_args* a, void* ws, size_t ws_bytes, int stages,
                           int flags, cudaStream_t stream, const LnFuse* ln_in, const LnFuse* ln_out);

int b200moe_ep_forward(b200moe_ep_ctx* c, const b200moe_layer_args* a, void* ws, size_t ws_bytes,
                       cudaStream_t stream) {
  return ep_forward_impl(c, a, ws, ws_bytes, kEpAllStages, 0, stream, nullptr, nullptr);
}

int b200moe_ep_forward_deferred(b200moe_ep_ctx* c, const b200moe_layer_args* a, void* ws, size_t ws_bytes,
                                cudaStream_t stream) {
  return ep_forward_impl(c, a, ws, ws_bytes, kEpAllStages, kEpFlagDeferWait, stream, nullptr, nullptr);
}

int b200moe_ep_wait(b200moe_ep_ctx* c, cudaStream_t stream) {
  if (!c) return fail(B200MOE_ERR_ARG, "ep_wait: null context");
  return ep_flush_pending(c, stream);
}

int b200moe_ep_out_buffer(const b200moe_ep_ctx* c, int slot, void** dev_ptr, size_t* bytes) {
  if (!c || !dev_ptr || slot < 0 || slot > 1) return fail(B200MOE_ERR_ARG, "ep_out_buffer: bad argument");
  const EpPeers& p = c->peers;
  const size_t one = sizeof(bf16) * static_cast<size_t>(p.cap) * p.D;
  *dev_ptr = p.base[p.rank] + p.lay.out_heap + slot * one;
  if (bytes) *bytes = one;
  return B200MOE_OK;
}

int b200moe_ep_forward_stages(b200moe_ep_ctx* c, const b200moe_layer_args* a, void* ws, size_t ws_bytes, int stages,
                              cudaStream_t stream) {
  return ep_forward_impl(c, a, ws, ws_bytes, stages, 0, stream, nullptr, nullptr);
}

// stages (bit mask): 1 gate + counts to every rank, 8 scatter (wait for the counts, push rows), 2 wait for the rows +
// expert FFN + push back, 4 wait for the results (+ combine); 15 = the whole layer, with the fused gate + dispatch
// kernel where it applies.
// ln_in: norm_ff fused into the route kernel (the caller has checked that the route kernel is taken);
// ln_out: norm_final fused into the combine kernel
static int ep_forward_impl(b200moe_ep_ctx* c, const b200moe_layer_args* a, void* ws, size_t ws_bytes, int stages,
                           int flags, cudaStream_t stream, const LnFuse* ln_in, const LnFuse* ln_out) {
  if (!c || !a) return fail(B200MOE_ERR_ARG, "ep_forward: null argument");
  const EpPeers& ep = c->peers;
  const int S = a->B * a->T;
  const int E_total = ep.E_local * ep.world;
  if (a->B < 0 || a->T < 0) return fail(B200MOE_ERR_ARG, "ep_forward: bad B/T");
  if (a->E != E_total) return fail(B200MOE_ERR_ARG, "ep_forward: router has E=%d, context has %d x %d", a->E, ep.E_local, ep.world);
  if (a->D != ep.D) return fail(B200MOE_ERR_ARG, "ep_forward: D=%d, context has %d", a->D, ep.D);
  if (a->D % 128 != 0 || a->H % 128 != 0)
    return fail(B200MOE_ERR_ARG, "ep_forward: D=%d and H=%d must be multiples of 128", a->D, a->H);
  if (a->dtype != B200MOE_BF16) return fail(B200MOE_ERR_ARG, "ep_forward: activations must be bf16");
  if (a->compute != B200MOE_COMPUTE_BF16) return fail(B200MOE_ERR_ARG, "ep_forward: bf16 compute only");
  if (a->top_k < 1 || a->top_k > 8 || a->top_k > a->E) return fail(B200MOE_ERR_ARG, "ep_forward: bad top_k %d", a->top_k);
  if (a->gate_mode == B200MOE_GATE_3M && a->top_k != 1) return fail(B200MOE_ERR_ARG, "ep_forward: the 3M router is top-1");
  if (a->act_type < 0 || a->act_type > 2) return fail(B200MOE_ERR_ARG, "ep_forward: bad act_type %d", a->act_type);
  if (stages < 1 || stages > kEpAllStages) return fail(B200MOE_ERR_ARG, "ep_forward: bad stage mask %d", stages);
  const int Sk = S * a->top_k;
  if (Sk > ep.cap) return fail(B200MOE_ERR_ARG, "ep_forward: %d entries exceed the context capacity %d", Sk, ep.cap);
  if ((S > 0 && (!a->x || !a->out)) || (!a->Wr && !a->Wr_packed) || !a->W1 || !a->W2 || !ws)
    return fail(B200MOE_ERR_ARG, "ep_forward: null pointer");
  const int Demb = a->embed ? a->Demb : 0;
  if (!a->Wr && !gate_tc_supported(a->D, Demb, a->E, a->top_k, a->dtype))
    return fail(B200MOE_ERR_ARG, "ep_forward: this shape needs the fp32 router (Wr)");
  const int rows_cap = ep.world * ep.cap;
  RouteWs w = carve_workspace(ws, rows_cap, E_total, a->D, a->H, 1);
  if (ws_bytes < w.bytes)
    return fail(B200MOE_ERR_WORKSPACE, "ep_forward: workspace %zu B < required %zu B", ws_bytes, w.bytes);
  int* idx = a->idx_out ? a->idx_out : w.idx;
  float* score = a->score_out ? a->score_out : w.score;

  // Folded combine: the owners' second-GEMM epilogue writes  residual + ff_scale * score * y  straight into this rank's
  // `out` rows.  Needs: top-1; the residual (if any) to be the very rows that were dispatched (the owner holds them); no
  // norm_final (the owner produces a row in D / 128 pieces); and `out` inside the symmetric buffer, at the same offset
  // on every rank, where the peers can reach it (b200moe_ep_out_buffer).  Every rank must decide the same way: the
  // decision travels with the counts and a mismatch raises the status word.
  const uint8_t* heap0 = ep.base[ep.rank] + ep.lay.out_heap;
  const size_t out_bytes = sizeof(bf16) * static_cast<size_t>(S) * a->D;
  const uint8_t* outp = static_cast<const uint8_t*>(a->out);
  // (decided from the pointers alone, also for a rank without tokens: it still serves the others as an owner)
  const bool out_in_heap =
      outp != nullptr && outp >= heap0 && outp + out_bytes <= heap0 + 2 * sizeof(bf16) * static_cast<size_t>(ep.cap) * a->D;
  static const int fold_env = env_int("B200MOE_EP_FOLD", 1);
  const bool fold = fold_env != 0 && a->top_k == 1 && ln_in == nullptr && ln_out == nullptr && out_in_heap &&
                    (a->residual == nullptr || a->residual == a->x);
  const int mode = (fold ? kEpModeFold : 0) | ((fold && a->residual) ? kEpModeResidual : 0);

  cudaError_t e = cudaSuccess;
  // the previous call's results may still be on their way into the buffer this call reads (or overwrites)
  if (stages & 1) {
    int rc = ep_flush_pending(c, stream);
    if (rc != B200MOE_OK) return rc;
  }
  const bool tc_gate = a->Wr_packed != nullptr && gate_tc_supported(a->D, Demb, a->E, a->top_k, a->dtype);
  // merged layout: every local expert's rows are contiguous whatever rank they came from, about Sk / E_local each
  const int bn = choose_bn(Sk, ep.E_local);
  const int gmax = max_groups(rows_cap, ep.world * ep.E_local, bn);  // one run of tiles per (local expert, source rank)
  const bool one_call = (stages & 11) == 11;   // gate, scatter and the expert kernel in this call
  const bool route = one_call && tc_gate && route_supported(S, a->D, Demb, a->E, a->top_k, a->dtype);
  void* drop_out = fold ? a->out : nullptr;     // padded / dropped tokens: output row = residual row, written locally
  const int ffn_ctas = ffn_grid_ctas(bn, gmax, a->D, a->H, false);  // announced with the counts (see EpLayout)
  if (route) {
    StageScope t(0, stream);
    e = launch_route(a->x, a->embed, ln_in ? ln_in->packed : a->Wr_packed, a->br, a->x_len, a->B, a->T, a->D, Demb,
                     a->E, a->gate_mode, a->keep_expert_output, idx, score, bn, w, a->counts_out, nullptr, a->mapping_out,
                     w.xbuf, drop_out, a->residual, stream, &ep, true, ln_in ? ln_in->gamma : nullptr,
                     ln_in ? ln_in->beta : nullptr, ln_in ? ln_in->eps : 0.0f,
                     ln_in ? ln_consts(ln_in, a->D + Demb) : nullptr, mode, ffn_ctas);
    if (e != cudaSuccess) return cuda_fail(e, "ep_forward/route");
  } else {
    if (S > 0 && (stages & 1)) {
      StageScope t(0, stream);
      if (tc_gate)
        e = launch_gate_tc(a->x, a->embed, a->Wr_packed, a->br, a->x_len, a->B, a->T, a->D, Demb, a->E, a->top_k,
                           a->gate_mode, idx, score, w.hist32, nullptr, 0, nullptr, 0, stream);
      else
        e = launch_gate(a->x, a->embed, a->Wr, a->br, a->x_len, a->B, a->T, a->D, Demb, a->E, a->top_k, a->gate_mode,
                        a->dtype, idx, score, stream);
      if (e != cudaSuccess) return cuda_fail(e, "ep_forward/gate");
    }
    const float* dscore = a->keep_expert_output ? nullptr : score;
    const int* hist = tc_gate && S > 0 ? w.hist32 : nullptr;
    if ((stages & 9) == 9) {          // counts + scatter in one kernel
      StageScope t(1, stream);
      e = launch_dispatch(a->x, idx, dscore, S, a->D, E_total, a->top_k, a->dtype, bn, w, a->counts_out, nullptr,
                          a->mapping_out, w.xbuf, drop_out, a->residual, hist, stream, &ep, one_call, false, mode, 0,
                          ffn_ctas);
    } else if (stages & 1) {          // staged: the counts alone ...
      e = launch_dispatch(a->x, idx, dscore, S, a->D, E_total, a->top_k, a->dtype, bn, w, nullptr, nullptr, nullptr,
                          w.xbuf, nullptr, nullptr, hist, stream, &ep, false, false, mode, 1, ffn_ctas);
    } else if (stages & 8) {          // ... then the scatter, once every rank's counts are on their way
      e = launch_dispatch(a->x, idx, dscore, S, a->D, E_total, a->top_k, a->dtype, bn, w, a->counts_out, nullptr,
                          a->mapping_out, w.xbuf, drop_out, a->residual, hist, stream, &ep, false, false, mode, 2,
                          ffn_ctas);
    }
    if (e != cudaSuccess) return cuda_fail(e, "ep_forward/dispatch");
  }
  if ((stages & 2) && !one_call) {
    StageScope t(1, stream);
    e = launch_ep_wait_rows(ep, stream);
    if (e != cudaSuccess) return cuda_fail(e, "ep_forward/wait_rows");
  }

  if (stages & 2) {
    FfnLaunch f{};
    f.xbuf = reinterpret_cast<const bf16*>(ep.base[ep.rank] + ep.lay.recv_x);
    f.hbuf = static_cast<bf16*>(w.hbuf);
    f.W1 = static_cast<const bf16*>(a->W1);
    f.W2 = static_cast<const bf16*>(a->W2);
    f.b1 = a->b1;
    f.b2 = a->b2;
    f.groups = w.groups;
    f.n_groups = w.n_groups;
    f.h_ready = w.h_ready;
    f.n_rows = rows_cap;
    f.E = ep.E_local;
    f.D = a->D;
    f.H = a->H;
    f.bn = bn;
    f.act = a->act_type;
    f.gmax = gmax;
    f.fused = 0;
    f.out_dtype = B200MOE_BF16;
    f.out = ep.base[ep.rank] + ep.lay.ret_y;
    f.top_k = 1;
    f.ff_scale = a->ff_scale;
    f.ep = &ep;
    f.ep_fold = fold ? 1 : 0;
    f.ep_out_off = fold ? static_cast<size_t>(outp - ep.base[ep.rank]) : 0;
    f.residual = fold ? ep.base[ep.rank] + ep.lay.recv_x : nullptr;   // (per row: only where the source had one)
    if (route) {
      f.clear_ptr = w.hist32;
      f.clear_ints = 2 * ((S + 31) / 32) * a->E;
    }
    StageScope t(2, stream);
    e = launch_ffn(f, stream);
    if (e != cudaSuccess) return cuda_fail(e, "ep_forward/expert_ffn");
  }
  if (stages & 4) {
    StageScope t(3, stream);
    if (fold) {
      c->pending = true;
      c->pending_out = a->out;
      c->pending_S = S;
      if (!(flags & kEpFlagDeferWait)) {
        int rc = ep_flush_pending(c, stream);
        if (rc != B200MOE_OK) return rc;
      }
    } else {
      e = launch_ep_combine(ep, w.mapping, a->keep_expert_output ? nullptr : score, a->residual, a->ff_scale, S, a->D,
                            a->top_k, a->out, stream, ln_out ? ln_out->gamma : nullptr, ln_out ? ln_out->beta : nullptr,
                            ln_out ? ln_out->eps : 0.0f);
      if (e != cudaSuccess) return cuda_fail(e, "ep_forward/combine");
    }
  }
  return B200MOE_OK;
}

// ---- plugin mirror ----------------------------------------------------------------------------------------------

b200moe_plugin* b200moe_plugin_create(int data_type, int num_expert, int idim, int hidden_units, int act_type) {
  // the reference creator rejects type_id outside {0, 1} (fmoe_expert_plugin.cpp:360-363); bf16 (2) is new here
  if (!dtype_ok(data_type)) {
    fail(B200MOE_ERR_ARG, "fmoe: invalid type_id %d", data_type);
    return nullptr;
  }
  if (num_expert < 1 || num_expert > kMaxExperts || idim < 1 || hidden_units < 1 || idim % 128 != 0 ||
      hidden_units % 128 != 0) {
    fail(B200MOE_ERR_ARG, "fmoe: unsupported num_expert=%d idim=%d hidden_units=%d", num_expert, idim, hidden_units);
    return nullptr;
  }
  b200moe_plugin* p = new (std::nothrow) b200moe_plugin();
  if (!p) return nullptr;
  p->data_type = data_type;
  p->num_expert = num_expert;
  p->idim = idim;
  p->hidden_units = hidden_units;
  p->act_type = act_type;
  std::memset(p->packed_src, 0, sizeof(p->packed_src));
  p->packed = nullptr;
  p->device = -1;
  return p;
}

b200moe_plugin* b200moe_plugin_clone(const b200moe_plugin* p) {
  if (!p) return nullptr;
  b200moe_plugin* q = b200moe_plugin_create(p->data_type, p->num_expert, p->idim, p->hidden_units, p->act_type);
  return q;
}

size_t b200moe_plugin_serialization_size(const b200moe_plugin*) { return 8 * sizeof(int); }

int b200moe_plugin_serialize(const b200moe_plugin* p, void* host_buffer) {
  if (!p || !host_buffer) return fail(B200MOE_ERR_ARG, "serialize: null pointer");
  int v[8] = {p->data_type, p->num_expert, p->idim, p->hidden_units, p->act_type, 0, 0, 0};
  std::memcpy(host_buffer, v, sizeof(v));
  return B200MOE_OK;
}

b200moe_plugin* b200moe_plugin_deserialize(const void* host_data, size_t length) {
  if (!host_data || length < 8 * sizeof(int)) {
    fail(B200MOE_ERR_ARG, "deserialize: need %zu bytes, got %zu", 8 * sizeof(int), length);
    return nullptr;
  }
  int v[8];
  std::memcpy(v, host_data, sizeof(v));
  return b200moe_plugin_create(v[0], v[1], v[2], v[3], v[4]);
}

namespace {
void plugin_release(b200moe_plugin* p) {
  if (p->packed != nullptr) {
    int cur = -1;
    cudaGetDevice(&cur);
    if (p->device >= 0 && p->device != cur) cudaSetDevice(p->device);
    cudaFree(p->packed);  // (synchronises with the device: no enqueue of this plugin is still reading it afterwards)
    if (p->device >= 0 && p->device != cur) cudaSetDevice(cur);
    p->packed = nullptr;
  }
  std::memset(p->packed_src, 0, sizeof(p->packed_src));
}

// Layout of the plugin-owned weight copies (independent of S).
struct PluginPack {
  size_t w1_off, w2_off, b1_off, b2_off, total;
};
PluginPack plugin_pack_layout(const b200moe_plugin* p) {
  PluginPack l;
  const size_t E = p->num_expert, D = p->idim, H = p->hidden_units;
  const size_t wbytes = p->data_type == B200MOE_BF16 ? 0 : E * H * D * sizeof(bf16);  // bf16 inputs are used in place
  size_t off = 0;
  l.w1_off = off;
  off = align_up(off + wbytes, 1024);
  l.w2_off = off;
  off = align_up(off + wbytes, 1024);
  l.b1_off = off;
  off = align_up(off + E * H * sizeof(float), 1024);
  l.b2_off = off;
  off = align_up(off + E * D * sizeof(float), 1024);
  l.total = off;
  return l;
}
}  // namespace

void b200moe_plugin_destroy(b200moe_plugin* p) {
  if (!p) return;
  plugin_release(p);
  delete p;
}

int b200moe_plugin_invalidate(b200moe_plugin* p) {
  if (!p) return fail(B200MOE_ERR_ARG, "invalidate: null plugin");
  std::memset(p->packed_src, 0, sizeof(p->packed_src));  // the buffer itself is kept and re-filled by the next enqueue
  return B200MOE_OK;
}

size_t b200moe_plugin_workspace_bytes(const b200moe_plugin* p, int S) {
  if (!p || S < 0) return 0;
  return carve_workspace(nullptr, S, p->num_expert, p->idim, p->hidden_units, 1).bytes;
}

int b200moe_plugin_enqueue(b200moe_plugin* p, const void* input, const int* gate_idx, const void* w1_weight,
                           const void* w1_bias, const void* w2_weight, const void* w2_bias, int S, void* output,
                           void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (!p) return fail(B200MOE_ERR_ARG, "enqueue: null plugin");
  if (S < 0) return fail(B200MOE_ERR_ARG, "enqueue: S < 0");
  if (S == 0) return B200MOE_OK;
  if (!input || !gate_idx || !w1_weight || !w1_bias || !w2_weight || !w2_bias || !output || !workspace)
    return fail(B200MOE_ERR_ARG, "enqueue: null pointer");
  const size_t need = carve_workspace(nullptr, S, p->num_expert, p->idim, p->hidden_units, 1).bytes;
  if (workspace_bytes < need)
    return fail(B200MOE_ERR_WORKSPACE, "enqueue: workspace %zu B < required %zu B", workspace_bytes, need);
  const int E = p->num_expert, D = p->idim, H = p->hidden_units;
  // The reference receives its weights as plugin inputs on every enqueue (README.md:225) and streams them as fp32.
  // Here they are cast once into memory the PLUGIN owns and reused while the four pointers stay the same -- never into
  // the scratch workspace: TensorRT shares that between the layers of an engine, does not preserve it between
  // enqueues, and its layout here depends on S (b200moe_plugin_invalidate after an in-place weight update).
  const void* src[4] = {w1_weight, w1_bias, w2_weight, w2_bias};
  const PluginPack l = plugin_pack_layout(p);
  if (p->packed == nullptr || std::memcmp(p->packed_src, src, sizeof(src)) != 0) {
    if (p->packed == nullptr) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(stream, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone)
        return fail(B200MOE_ERR_ARG, "enqueue: the first enqueue of a plugin allocates its weight copies and cannot be "
                                      "captured into a CUDA graph; run it once eagerly first");
      cudaError_t e = cudaGetDevice(&p->device);
      if (e == cudaSuccess) e = cudaMalloc(&p->packed, l.total);
      if (e != cudaSuccess) {
        p->packed = nullptr;
        return cuda_fail(e, "enqueue/cudaMalloc of the packed weights");
      }
    }
    char* pk = static_cast<char*>(p->packed);
    cudaError_t e = cudaSuccess;
    if (p->data_type != B200MOE_BF16) {
      const size_t nw = static_cast<size_t>(E) * H * D;
      e = launch_pack_bf16(w1_weight, p->data_type, reinterpret_cast<bf16*>(pk + l.w1_off), nw, stream);
      if (e == cudaSuccess)
        e = launch_pack_bf16(w2_weight, p->data_type, reinterpret_cast<bf16*>(pk + l.w2_off), nw, stream);
      if (e != cudaSuccess) return cuda_fail(e, "enqueue/pack");
    }
    to_f32_kernel<<<64, 256, 0, stream>>>(w1_bias, p->data_type, reinterpret_cast<float*>(pk + l.b1_off),
                                          static_cast<size_t>(E) * H);
    to_f32_kernel<<<64, 256, 0, stream>>>(w2_bias, p->data_type, reinterpret_cast<float*>(pk + l.b2_off),
                                          static_cast<size_t>(E) * D);
    count_launch(2);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "enqueue/bias");
    std::memcpy(p->packed_src, src, sizeof(src));
  }
  char* pk = static_cast<char*>(p->packed);
  const bool in_place = p->data_type == B200MOE_BF16;
  const bf16* w1p = in_place ? static_cast<const bf16*>(w1_weight) : reinterpret_cast<const bf16*>(pk + l.w1_off);
  const bf16* w2p = in_place ? static_cast<const bf16*>(w2_weight) : reinterpret_cast<const bf16*>(pk + l.w2_off);
  const float* b1p = reinterpret_cast<const float*>(pk + l.b1_off);
  const float* b2p = reinterpret_cast<const float*>(pk + l.b2_off);
  RouteWs w = carve_workspace(workspace, S, E, D, H, 1);
  const int bn = choose_bn(S, E);
  const int gmax = max_groups(S, E, bn);
  cudaError_t e = launch_dispatch(input, gate_idx, nullptr, S, D, E, 1, p->data_type, bn, w, nullptr, nullptr, nullptr,
                                  w.xbuf, output, nullptr, nullptr, stream);
  if (e != cudaSuccess) return cuda_fail(e, "enqueue/dispatch");
  FfnLaunch f{};
  f.xbuf = w.xbuf;
  f.hbuf = static_cast<bf16*>(w.hbuf);
  f.W1 = w1p;
  f.W2 = w2p;
  f.b1 = b1p;
  f.b2 = b2p;
  f.groups = w.groups;
  f.n_groups = w.n_groups;
  f.h_ready = w.h_ready;
  f.n_rows = S;
  f.E = E;
  f.D = D;
  f.H = H;
  f.bn = bn;
  // the reference stores act_type but always applies SiLU (fmoe_expert_plugin.cpp:106); honour the field here,
  // with the reference's default 0 = SiLU
  f.act = (p->act_type >= 0 && p->act_type <= 2) ? p->act_type : B200MOE_ACT_SILU;
  f.gmax = gmax;
  f.fused = 1;  // un-weighted output in token order: score = 1, ff_scale = 1, no residual
  f.out_dtype = p->data_type;
  f.out = output;
  f.residual = nullptr;
  f.pos = w.pos;
  f.row_score = nullptr;
  f.ff_scale = 1.0f;
  f.top_k = 1;
  e = launch_ffn(f, stream);
  if (e != cudaSuccess) return cuda_fail(e, "enqueue/expert_ffn");
  return B200MOE_OK;
}

// ---- the block around the layer: norm_ff in front, norm_final behind (fmoe_transformer.py:144-166) -----------------
namespace {
inline size_t align256(size_t n) { return (n + 255) & ~static_cast<size_t>(255); }
inline size_t dtype_bytes(int dtype) { return dtype == B200MOE_F32 ? 4 : 2; }

// The normalised input lives behind the layer's own workspace.
struct BlockPlan {
  b200moe_layer_args layer;
  size_t layer_ws;
  void* xn;
};

int plan_block(const b200moe_block_args* b, size_t layer_ws, void* ws, size_t ws_bytes, BlockPlan* plan, const char* who) {
  const b200moe_layer_args& a = b->layer;
  const int S = a.B * a.T;
  const bool ln_in = b->norm_ff_gamma != nullptr, ln_out = b->norm_final_gamma != nullptr;
  if ((ln_in && !b->norm_ff_beta) || (ln_out && !b->norm_final_beta))
    return fail(B200MOE_ERR_ARG, "%s: a LayerNorm needs both gamma and beta", who);
  if ((ln_in || ln_out) && !layernorm_supported(a.D))
    return fail(B200MOE_ERR_ARG, "%s: LayerNorm over D=%d is not supported (multiple of 8, at most 1024)", who, a.D);
  if (!(b->eps >= 0.0f)) return fail(B200MOE_ERR_ARG, "%s: bad eps", who);
  plan->layer = a;
  plan->layer_ws = layer_ws;
  plan->xn = nullptr;
  if (ln_in && S > 0) {
    const size_t need = align256(layer_ws) + static_cast<size_t>(S) * a.D * dtype_bytes(a.dtype);
    if (!ws || ws_bytes < need)
      return fail(B200MOE_ERR_WORKSPACE, "%s: workspace %zu B < required %zu B", who, ws_bytes, need);
    plan->xn = static_cast<uint8_t*>(ws) + align256(layer_ws);
    plan->layer.x = plan->xn;
  }
  return B200MOE_OK;
}
}  // namespace

size_t b200moe_block_workspace_bytes(int S, int E, int D, int H, int top_k) {
  const size_t layer = b200moe_workspace_bytes(S, E, D, H, top_k);
  if (layer == 0) return 0;
  return align256(layer) + static_cast<size_t>(S) * D * 4;
}

size_t b200moe_ep_block_workspace_bytes(const b200moe_ep_ctx* c, int H) {
  const size_t layer = b200moe_ep_workspace_bytes(c, H);
  if (layer == 0) return 0;
  return align256(layer) + static_cast<size_t>(c->peers.cap) * c->peers.D * 4;
}

int b200moe_layernorm(const void* in, const float* gamma, const float* beta, float eps, int S, int D, int dtype,
                      void* out, cudaStream_t stream) {
  if (S < 0 || !dtype_ok(dtype)) return fail(B200MOE_ERR_ARG, "layernorm: bad S=%d / dtype=%d", S, dtype);
  if (!layernorm_supported(D)) return fail(B200MOE_ERR_ARG, "layernorm: D=%d is not supported (multiple of 8, at most 1024)", D);
  if (S > 0 && (!in || !out || !gamma || !beta)) return fail(B200MOE_ERR_ARG, "layernorm: null pointer");
  cudaError_t e = launch_layernorm(in, gamma, beta, eps, S, D, dtype, out, stream);
  if (e != cudaSuccess) return cuda_fail(e, "layernorm");
  return B200MOE_OK;
}

int b200moe_att_masked_softmax(const void* in, const int* mask, float scale, int B, int N, int S, int ld, int dtype,
                               void* out, cudaStream_t stream) {
  if (B < 0 || N < 0 || S < 0 || !dtype_ok(dtype)) return fail(B200MOE_ERR_ARG, "att_masked_softmax: bad shape / dtype");
  if (static_cast<long long>(B) * N * S == 0) return B200MOE_OK;
  if (!masked_softmax_supported(ld)) return fail(B200MOE_ERR_ARG, "att_masked_softmax: ld=%d outside [1, 1024]", ld);
  if (!in || !out) return fail(B200MOE_ERR_ARG, "att_masked_softmax: null pointer");
  cudaError_t e = launch_att_masked_softmax(in, mask, scale, B, N, S, ld, dtype, out, stream);
  if (e != cudaSuccess) return cuda_fail(e, "att_masked_softmax");
  return B200MOE_OK;
}

int b200moe_glu(const void* x, int M, int C, int N, int dtype, void* y, cudaStream_t stream) {
  if (M < 0 || C < 0 || N < 0 || !dtype_ok(dtype)) return fail(B200MOE_ERR_ARG, "glu: bad shape / dtype");
  if (static_cast<long long>(M) * C * N > 0 && (!x || !y)) return fail(B200MOE_ERR_ARG, "glu: null pointer");
  cudaError_t e = launch_glu(x, M, C, N, dtype, y, stream);
  if (e != cudaSuccess) return cuda_fail(e, "glu");
  return B200MOE_OK;
}

int b200moe_masked_fill(const void* in, const int* mask, float fill, int B, int dim, int T, int dtype, void* out,
                        cudaStream_t stream) {
  if (B < 0 || dim < 0 || T < 0 || !dtype_ok(dtype)) return fail(B200MOE_ERR_ARG, "masked_fill: bad shape / dtype");
  if (static_cast<long long>(B) * dim * T > 0 && (!in || !mask || !out)) return fail(B200MOE_ERR_ARG, "masked_fill: null pointer");
  cudaError_t e = launch_masked_fill(in, mask, fill, B, dim, T, dtype, out, stream);
  if (e != cudaSuccess) return cuda_fail(e, "masked_fill");
  return B200MOE_OK;
}

int b200moe_rel_pos_encoding(const void* in, const void* pe, float scale, int B, int T, int D, int dtype, void* out,
                             void* pos_emb, cudaStream_t stream) {
  if (B < 0 || T < 0 || D < 0 || !dtype_ok(dtype)) return fail(B200MOE_ERR_ARG, "rel_pos_encoding: bad shape / dtype");
  if (static_cast<long long>(B) * T * D > 0 && (!in || !pe || !out || !pos_emb))
    return fail(B200MOE_ERR_ARG, "rel_pos_encoding: null pointer");
  cudaError_t e = launch_rel_pos_encoding(in, pe, scale, B, T, D, dtype, out, pos_emb, stream);
  if (e != cudaSuccess) return cuda_fail(e, "rel_pos_encoding");
  return B200MOE_OK;
}

int b200moe_block_forward(const b200moe_block_args* b, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (!b) return fail(B200MOE_ERR_ARG, "block_forward: null args");
  const b200moe_layer_args& a = b->layer;
  if (a.B < 0 || a.T < 0 || a.E < 1 || a.D < 1 || a.H < 0 || a.top_k < 1)
    return fail(B200MOE_ERR_ARG, "block_forward: bad shape");
  const int S = a.B * a.T;
  BlockPlan plan;
  const size_t layer_ws = b200moe_workspace_bytes(S, a.E, a.D, a.H, a.top_k);
  int rc = plan_block(b, layer_ws, ws, ws_bytes, &plan, "block_forward");
  if (rc != B200MOE_OK) return rc;
  const LnFuse lin{b->norm_ff_gamma, b->norm_ff_beta, b->eps, b->Wr_packed_ln};
  const LnFuse lout{b->norm_final_gamma, b->norm_final_beta, b->eps, nullptr};
  // small bf16 batches: the route kernel normalises the rows it has just fetched; otherwise a row pass in front
  const bool fuse_in = plan.xn != nullptr && ln_fuse_mode() != 0 && b->Wr_packed_ln != nullptr && takes_route_kernel(&a);
  if (fuse_in) plan.layer.x = a.x;
  if (plan.xn && !fuse_in) {
    if (!a.x) return fail(B200MOE_ERR_ARG, "block_forward: null pointer");
    StageScope t(0, stream);
    cudaError_t e = launch_layernorm(a.x, b->norm_ff_gamma, b->norm_ff_beta, b->eps, S, a.D, a.dtype, plan.xn, stream);
    if (e != cudaSuccess) return cuda_fail(e, "block_forward/norm_ff");
  }
  return forward_impl(&plan.layer, ws, ws_bytes < layer_ws ? ws_bytes : layer_ws, stream, fuse_in ? &lin : nullptr,
                      (b->norm_final_gamma && S > 0) ? &lout : nullptr);
}

int b200moe_ep_block_forward(b200moe_ep_ctx* c, const b200moe_block_args* b, void* ws, size_t ws_bytes,
                             cudaStream_t stream) {
  if (!c || !b) return fail(B200MOE_ERR_ARG, "ep_block_forward: null argument");
  const b200moe_layer_args& a = b->layer;
  if (a.B < 0 || a.T < 0) return fail(B200MOE_ERR_ARG, "ep_block_forward: bad B/T");
  const int S = a.B * a.T;
  BlockPlan plan;
  const size_t layer_ws = b200moe_ep_workspace_bytes(c, a.H);
  int rc = plan_block(b, layer_ws, ws, ws_bytes, &plan, "ep_block_forward");
  if (rc != B200MOE_OK) return rc;
  const LnFuse lin{b->norm_ff_gamma, b->norm_ff_beta, b->eps, b->Wr_packed_ln};
  const LnFuse lout{b->norm_final_gamma, b->norm_final_beta, b->eps, nullptr};
  const bool fuse_in = plan.xn != nullptr && ln_fuse_mode() != 0 && b->Wr_packed_ln != nullptr && takes_route_kernel(&a);
  if (fuse_in) plan.layer.x = a.x;
  if (plan.xn && !fuse_in) {
    if (!a.x) return fail(B200MOE_ERR_ARG, "ep_block_forward: null pointer");
    cudaError_t e = launch_layernorm(a.x, b->norm_ff_gamma, b->norm_ff_beta, b->eps, S, a.D, a.dtype, plan.xn, stream);
    if (e != cudaSuccess) return cuda_fail(e, "ep_block_forward/norm_ff");
  }
  // norm_final rides in the combine kernel (it runs even for a rank without tokens: the flags must be consumed)
  return ep_forward_impl(c, &plan.layer, ws, ws_bytes < layer_ws ? ws_bytes : layer_ws, kEpAllStages, 0, stream,
                         fuse_in ? &lin : nullptr, b->norm_final_gamma ? &lout : nullptr);
}

}  // extern "C"
