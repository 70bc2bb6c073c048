// Internal declarations shared by the kernels and the C-ABI layer (api.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <utility>

#include "../../include/b200moe.h"

namespace b200moe {

using bf16 = __nv_bfloat16;

constexpr int kMaxExperts = 256;      // smem tables in gate/dispatch are sized for this
constexpr int kDispatchThreads = 256;  // 8 warps per dispatch CTA
constexpr int kMaxChunks = 296;        // 2 x 148 dispatch chunks at most (make_chunking divides by this)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// One GEMM "group" = one expert x one tile of <= BN of its tokens.
struct GroupRec {
  int expert;  // (local) expert whose weights the group uses
  int row0;    // first row of the group in xbuf / hbuf
  int nrows;   // 1 .. BN
  int src;     // unused (0)
  int orow0;   // first output row of the group in ybuf (== row0)
  int pad[3];
};

// ---- expert parallelism over peer-mapped memory (ep.cu) -------------------------------------------------------
constexpr int kMaxEpWorld = 8;

// Byte offsets inside the symmetric buffer every rank allocates (identical on all ranks).
//
// Protocol of one layer call `seq` on rank r (W ranks, E_local experts per rank, E = W * E_local):
//   1. gate -> r's per-expert counts for ALL E experts.  One CTA stores them into cnt_all[r][:] of EVERY rank, each as a
//      self-validating word (seq << 32 | count).
//   2. every owner keeps one segment of `cap` receive rows per SOURCE rank; a source lays its rows for an owner out in its
//      own expert order, so where r's rows belong follows from r's own offsets: nothing is waited for before the rows
//      leave.  (An earlier layout merged the sources' rows of an expert into contiguous rows: full token tiles, but every
//      rank's counts had to arrive before the first row could leave -- a system-scope round trip in front of the pushes.)
//      The counts are needed by the owner only, for its expert kernel's group table: one run of token tiles per (local
//      expert, source rank), built by one CTA at the end of the dispatch / route kernel (functions.py:37-50's count
//      exchange, without the host and off the senders' path).
//   3. rows -> recv_x[row] and 8 bytes of routing data -> meta[row] on the owner; every CTA: fence.acq_rel.sys, then one
//      remote increment of arrive[r] on every rank ("one more CTA of r has delivered").  How many CTAs there will have been
//      after this call travels with the counts (cumulative, never reset), so nobody has to be the last.
//   4. owner: expert FFN over recv_x.  Its producer streams the first weight tiles at once and waits for the arrival
//      counters only before it touches the rows.  The second GEMM's epilogue sends each row where meta says: either
//      (fold) the finished layer output  residual + ff_scale * score * y  straight into the source rank's `out` row, or
//      the bare y row into the source's ret_y for ep_combine.  Every CTA: fence.acq_rel.sys, then one remote increment of
//      done[owner] on every rank.
//   5. source: waits until every owner's done counter has reached what that owner announced (ep_wait_done_kernel in front
//      of whatever consumes `out`, or ep_combine).
// Counters and sequence numbers only ever grow.  One receive buffer suffices: r pushes layer L+1 only after it has seen
// every owner finish layer L, i.e. after the owner's last read of recv_x / meta / cnt_all.
struct EpLayout {
  size_t ctrl;       // int32[64], local bookkeeping: [0] seq (layer calls so far), [1] dispatch CTAs done (this call),
                     //   [3] error, [5] / [6] cumulative dispatch / FFN CTAs this rank has announced,
                     //   [16 + s] arrive[s] to wait for in this call, [32 + o] done[o] to wait for in this call
  size_t arrive;     // int32[kMaxEpWorld]: arrive[s] = CTAs of rank s's dispatch kernels that have delivered (cumulative)
  size_t done;       // int32[kMaxEpWorld]: done[o] = CTAs of owner o's expert kernels that have delivered (cumulative)
  size_t cnt_all;    // uint64[kMaxEpWorld][E + 3]: [s][e] = (seq << 32) | rows rank s routes to global expert e in call seq;
                     //   [s][E] mode bits, [s][E + 1] / [s][E + 2] the arrive / done counts rank s will have reached after
                     //   this call.  Self-validating words: no flag, no fence behind them
  size_t meta;       // int2 [world * cap]: per received row {residual? << 31 | source rank << 27 | index at the source,
                     //   gate score bits}; index = the token (fold) or the source's expert-order row (ret_y + ep_combine)
  size_t recv_x;     // bf16 [world][cap][D]: received rows, one segment per source rank, in the source's expert order
  size_t ret_y;      // bf16 [cap][D]: expert outputs for this rank's own entries, in its expert order (un-folded path)
  size_t out_heap;   // bf16 [2][cap][D]: two layer-output buffers inside the symmetric heap (the folded path writes the
                     //   source rank's output rows remotely, so `out` has to live where the peers can reach it)
  size_t bytes;
};

constexpr int kEpModeFold = 1, kEpModeResidual = 2;
constexpr int kEpCntExtra = 3;                       // words behind the E counts of a count message
constexpr int kEpCtrlArrive = 16, kEpCtrlDone = 32;  // ctrl[] slots of the counter values to wait for
constexpr int kEpMetaRankShift = 27;
constexpr int kEpMetaIndexMask = (1 << kEpMetaRankShift) - 1;
constexpr int kEpMetaRankMask = 0xF;
__host__ __device__ inline int ep_meta_word(int rank, int index, bool residual) {
  return static_cast<int>((residual ? 0x80000000u : 0u) | (static_cast<unsigned>(rank) << kEpMetaRankShift) |
                          static_cast<unsigned>(index));
}

inline EpLayout ep_layout(int world, int E_local, int D, int cap) {
  EpLayout l;
  size_t off = 0;
  auto take = [&](size_t b) {
    size_t o = off;
    off = align_up(off + b, 1024);
    return o;
  };
  const size_t E = static_cast<size_t>(world) * E_local;
  l.ctrl = take(sizeof(int) * 64);
  l.arrive = take(sizeof(int) * kMaxEpWorld);
  l.done = take(sizeof(int) * kMaxEpWorld);
  l.cnt_all = take(sizeof(unsigned long long) * kMaxEpWorld * (E + kEpCntExtra));
  l.meta = take(sizeof(int) * 2 * static_cast<size_t>(world) * cap);
  l.recv_x = take(sizeof(bf16) * static_cast<size_t>(world) * cap * D);
  l.ret_y = take(sizeof(bf16) * static_cast<size_t>(cap) * D);
  l.out_heap = take(sizeof(bf16) * 2 * static_cast<size_t>(cap) * D);
  l.bytes = off;
  return l;
}

// Passed by value to the kernels: where every rank's symmetric buffer is mapped in THIS process.
struct EpPeers {
  int rank, world, E_local, cap, D;
  int timeout_ms;            // spins on remote flags give up after this long and raise ctrl[3]
  uint8_t* base[kMaxEpWorld];  // base[rank] is the local buffer
  EpLayout lay;
};

// Device-side routing state produced by dispatch and consumed by the FFN kernel. Lives in the workspace.
struct RouteWs {
  int* chunk_hist;   // [kMaxChunks, E] per-dispatch-chunk expert counts (count kernel)
  int* hist32;       // [ceil(S/32), E] per-32-token expert counts (tensor-core gate: plain ints; route kernel: the same
                     //   region as 64-bit self-validating words, hence sized for 8 bytes per entry)
  int* counts;       // [E]
  int* offsets;      // [E + 1]
  int* mapping;      // [Sk]   entry -> expert-order row (-1 = dropped)
  int* pos;          // [Sk]   expert-order row -> entry (fastmoe's `pos`)
  float* row_score;  // [Sk]   gate score of the entry that landed in this row
  GroupRec* groups;  // [Gmax]
  int* n_groups;     // [1]
  int* h_ready;      // [Gmax] G1 tiles finished per group (dependency flags for the second GEMM)
  int* idx;          // [Sk]   gate output when the caller does not want it
  float* score;      // [Sk]
  bf16* xbuf;        // [Sk, D]
  void* hbuf;        // [Sk, H]  bf16 (or fp32 for the tf32 path)
  void* ybuf;        // [Sk, D]  staging for the un-fused combine (top_k > 1)
  size_t bytes;
};

inline int max_groups(int Sk, int E, int bn) { return (Sk + bn - 1) / bn + E; }

// Carves `base` (may be null: then only sizes are computed). Worst case group count uses BN = 32.
inline RouteWs carve_workspace(void* base, int S, int E, int D, int H, int top_k) {
  RouteWs w;
  const size_t Sk = static_cast<size_t>(S) * top_k;
  const size_t gmax = static_cast<size_t>(max_groups(static_cast<int>(Sk), E, 32));
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return base ? static_cast<char*>(base) + o : static_cast<char*>(nullptr);
  };
  w.chunk_hist = reinterpret_cast<int*>(take(sizeof(int) * kMaxChunks * E));
  w.hist32 = reinterpret_cast<int*>(take(sizeof(long long) * ((static_cast<size_t>(S) + 31) / 32) * E));
  w.counts = reinterpret_cast<int*>(take(sizeof(int) * E));
  w.offsets = reinterpret_cast<int*>(take(sizeof(int) * (E + 1)));
  w.mapping = reinterpret_cast<int*>(take(sizeof(int) * Sk));
  w.pos = reinterpret_cast<int*>(take(sizeof(int) * Sk));
  w.row_score = reinterpret_cast<float*>(take(sizeof(float) * Sk));
  w.groups = reinterpret_cast<GroupRec*>(take(sizeof(GroupRec) * gmax));
  w.n_groups = reinterpret_cast<int*>(take(sizeof(int) * 8));  // [0] groups, [2..5] the route kernel's barrier words
  w.h_ready = reinterpret_cast<int*>(take(sizeof(int) * gmax));
  w.idx = reinterpret_cast<int*>(take(sizeof(int) * Sk));
  w.score = reinterpret_cast<float*>(take(sizeof(float) * Sk));
  // xbuf / hbuf are sized for fp32 rows so that the same workspace serves the TF32 flavour
  w.xbuf = reinterpret_cast<bf16*>(take(sizeof(float) * Sk * D));
  w.ybuf = take(sizeof(float) * Sk * D);
  w.hbuf = take(sizeof(float) * Sk * H);  // last: the only field whose size depends on H
  w.bytes = off;
  return w;
}

// ---- kernel launchers (each returns a cudaError_t from the launch; none synchronises) ----------------------

// gate.cu
cudaError_t launch_gate(const void* x, const void* embed, const float* Wr, const float* br, const int* x_len, int B,
                        int T, int D, int Demb, int E, int top_k, int gate_mode, int dtype, int* idx, float* score,
                        cudaStream_t stream);
// gate_tc.cu: tensor-core gate for bf16 activations and E <= 32 (router packed by launch_pack_router)
bool gate_tc_supported(int D, int Demb, int E, int top_k, int dtype);
size_t router_pack_bytes(int R);
cudaError_t launch_pack_router(const float* Wr, int R, int E, void* packed, cudaStream_t stream);
cudaError_t launch_gate_tc(const void* x, const void* embed, const void* wr_packed, const float* br, const int* x_len,
                           int B, int T, int D, int Demb, int E, int top_k, int gate_mode, int* idx, float* score,
                           int* hist32, const void* pf0, size_t pf0_bytes, const void* pf1, size_t pf1_bytes,
                           cudaStream_t stream);
cudaError_t launch_softmax_topk(const void* logits, const int* mask, int B, int T, int E, int dtype, void* value,
                                int* idx, cudaStream_t stream);

// dispatch.cu
// bn: token-tile width the FFN kernel will use (group table is built for it); score may be null (row_score = 1).
// drop_out (optional, `dtype`, [S, D], top_k == 1 only): rows of dropped tokens (idx < 0) are written here as
// drop_residual[row] (or zeros), so that a fused FFN epilogue -- which only touches routed tokens -- leaves a
// fully defined output.
// hist32 (optional): per-32-token expert counts already produced by the tensor-core gate; the count kernel is skipped.
cudaError_t launch_dispatch(const void* x, const int* idx, const float* score, int S, int D, int E, int top_k,
                            int dtype, int bn, const RouteWs& ws, int* counts_out, int* offsets_out,
                            int* mapping_out, bf16* xbuf, void* drop_out, const void* drop_residual,
                            const int* hist32, cudaStream_t stream, const EpPeers* ep = nullptr,
                            bool ep_fold_wait = false, bool xbuf_f32 = false, int ep_mode = 0, int ep_phase = 0,
                            int ep_ffn_ctas = 0);
constexpr int kMaxHistRows = 512;  // above this many 32-token rows the scatter CTAs would re-read too much
// Builds only the group table (+ zeroes the flags) from an existing offsets array.
cudaError_t launch_build_groups(const int* offsets, int E, int bn, GroupRec* groups, int* n_groups, int* h_ready,
                                int gmax, cudaStream_t stream);
int choose_bn(int Sk, int E);

// route.cu: gate + dispatch in one launch for small batches (bf16, E <= 32, top-1); same outputs as launch_gate_tc
// followed by launch_dispatch.
bool route_supported(int S, int D, int Demb, int E, int top_k, int dtype);
cudaError_t read_route_status(int* host_status, bool clear);  // synchronous read of the device status word
void set_route_trace(void* dev_buf);  // debug: 16 records of 16 B per CTA, see tools/route_trace.py
cudaError_t launch_route(const void* x, const void* embed, const void* wr_packed, const float* br, const int* x_len,
                         int B, int T, int D, int Demb, int E, int gate_mode, int keep_expert_output, int* idx,
                         float* score, int bn, const RouteWs& ws, int* counts_out, int* offsets_out, int* mapping_out,
                         bf16* xbuf, void* drop_out, const void* drop_residual, cudaStream_t stream,
                         const EpPeers* ep = nullptr, bool ep_fold_wait = false, const float* ln_gamma = nullptr,
                         const float* ln_beta = nullptr, float ln_eps = 0.0f, const float* ln_c = nullptr,
                         int ep_mode = 0, int ep_ffn_ctas = 0);
// Router packed for the route kernel's fused norm_ff: like launch_pack_router, with the x rows (k >= R - D) scaled by
// gamma, followed by c1[32] = gamma^T Wr_x and c0[32] = beta^T Wr_x (fp32).  router_ln_pack_bytes(R) bytes.
size_t router_ln_pack_bytes(int R);
cudaError_t launch_pack_router_ln(const float* Wr, int R, int E, int D, const float* gamma, const float* beta,
                                  void* packed, cudaStream_t stream);

// ffn.cu
struct FfnLaunch {
  const bf16* xbuf;   // [n_rows, D]
  bf16* hbuf;         // [n_rows, H]
  const bf16* W1;     // [E, H, D]
  const bf16* W2;     // [E, D, H]
  const float* b1;    // [E, H] or null
  const float* b2;    // [E, D] or null
  const GroupRec* groups;
  const int* n_groups;
  int* h_ready;
  int n_rows, E, D, H, bn, act;
  int gmax;           // upper bound on the number of groups (sizes the grid)
  // epilogue of the second GEMM
  int fused;          // 0: ybuf[row] = y (expert order). 1: out[pos[row]/top_k] = residual + ff_scale*score*y
  int out_dtype;
  void* out;          // ybuf (fused = 0) or the layer output (fused = 1)
  const void* residual;
  const int* pos;
  const float* row_score;  // null => 1
  float ff_scale;
  int top_k;
  const EpPeers* ep;  // expert parallelism: every row goes to the rank its routing data (EpLayout::meta) names
  int ep_fold;        // 1: finished output rows (residual = the received row, x ff_scale x score) into the source's `out`
  size_t ep_out_off;  //    at this byte offset of every rank's symmetric buffer; 0: bare y rows into the source's ret_y
  int* clear_ptr;     // optional: `clear_ints` ints zeroed at kernel start (the route kernel's tagged histogram)
  int clear_ints;
  int tf32;           // 1: xbuf, W1, W2 and hbuf hold fp32 (the pointers above are reinterpreted), TF32 tensor-core math
  int p1_only;        // 1: first GEMM only -- hbuf [n_rows, H] = act(xbuf . W1^T + b1) is the result (one grouped linear:
                      //    MOELinear / MOEbiasLinear, trainer_3m_fix/fmoe/functions.py:107-152); W2 / b2 / out unused
};
cudaError_t launch_ffn(const FfnLaunch& a, cudaStream_t stream);
// Grid the expert kernel will be launched with (expert parallelism announces it ahead of the launch).
int ffn_grid_ctas(int bn, int gmax, int D, int H, bool tf32);
// Debug timeline: every following ffn launch records per-CTA events into dev_buf (16 B records); null disables.
void set_ffn_trace(void* dev_buf, int records_per_cta);
// Debug, across kernels: every launch of the route and expert kernels takes the next slot of 148 CTAs x 8 marks
// (%globaltimer ns; 0 = not reached) while a buffer is set -- when its CTAs started, passed their dependency wait,
// saw their first data, finished.  Works under CUDA-graph capture (the slot is frozen into the captured launch).
// kind 1 = route, 2 = expert kernel; the slot header is written by the host side into `kinds`.
constexpr int kTimelineMarks = 8;
void set_timeline(void* dev_buf, int max_launches);
unsigned long long* next_timeline_slot(int kind);  // nullptr when off or full
int timeline_kind(int slot);

// ep.cu
// Waits until every rank's rows of the current layer call have landed (staged drivers; the one-call path folds the wait
// into the dispatch / route kernel).
cudaError_t launch_ep_wait_rows(const EpPeers& ep, cudaStream_t stream);
// Folded path: returns (on the stream) once every owner has written this rank's output rows of the current layer call;
// poisons out [S, D] bf16 with NaNs when a peer failed to deliver.
cudaError_t launch_ep_wait_done(const EpPeers& ep, void* out, int S, int D, cudaStream_t stream);
// Waits until every expert rank has returned this rank's rows, then out = residual + ff_scale * sum_k score * ret_y[mapping].
// ln_gamma / ln_beta non-null: LayerNorm over the D features of every output row on top (norm_final).
cudaError_t launch_ep_combine(const EpPeers& ep, const int* mapping, const float* score, const void* residual,
                              float ff_scale, int S, int D, int top_k, void* out, cudaStream_t stream,
                              const float* ln_gamma = nullptr, const float* ln_beta = nullptr, float ln_eps = 0.0f);

// layernorm.cu
// out[s, :] = LayerNorm(in[s, :]) * gamma + beta over D features (biased variance, eps inside the root); in == out allowed.
bool layernorm_supported(int D);
cudaError_t launch_layernorm(const void* in, const float* gamma, const float* beta, float eps, int S, int D, int dtype,
                             void* out, cudaStream_t stream);

// combine.cu
cudaError_t launch_combine(const void* ybuf, const int* mapping, const float* score, const void* residual,
                           float ff_scale, int S, int D, int top_k, int dtype, void* out, cudaStream_t stream,
                           const float* ln_gamma = nullptr, const float* ln_beta = nullptr, float ln_eps = 0.0f);
// The other encoder plugins (encoder_ops.cu)
bool masked_softmax_supported(int ld);
cudaError_t launch_att_masked_softmax(const void* in, const int* mask, float scale, int B, int N, int S, int ld, int dtype,
                                      void* out, cudaStream_t stream);
cudaError_t launch_glu(const void* x, long long M, int C, int N, int dtype, void* y, cudaStream_t stream);
cudaError_t launch_masked_fill(const void* in, const int* mask, float fill, int B, int dim, int T_len, int dtype, void* out,
                               cudaStream_t stream);
cudaError_t launch_rel_pos_encoding(const void* in, const void* pe, float scale, int B, int T_len, int D, int dtype,
                                    void* out, void* pos_emb, cudaStream_t stream);
cudaError_t launch_scatter_rows(const void* in, const int* index, int n, int n_out, int row_bytes, void* out,
                                cudaStream_t stream);
cudaError_t launch_pack_bf16(const void* src, int src_dtype, bf16* dst, size_t n, cudaStream_t stream);
cudaError_t launch_pack_tf32(const float* src, float* dst, size_t n, cudaStream_t stream);  // round to nearest TF32

void count_launch(int n = 1);

// Tunables read once from the environment (api.cu):
//   B200MOE_PDL=m       bit mask of the kernels launched with programmatic dependent launch (1 gate / route, 2 dispatch,
//                       4 expert FFN, 8 LayerNorm; default 13); 0 = ordinary stream ordering.  B200MOE_PDL_TRIG=m: which of them
//                       release their dependents at their start instead of at exit (default 4)
//   B200MOE_PREFETCH=0  no L2 prefetch of the layer's expert weights from the gate kernel; 1 (default) = issued before
//                       the gate waits for the previous kernel; 2 = issued after that wait
//   B200MOE_2PROD=0     expert kernel with ONE TMA producer thread for both operands (default 1: weights from warp 0, token
//                       rows from warp 2; ffn.cu); B200MOE_KPS=1: one k-block per TMA instruction instead of two
//   B200MOE_ROUTE_TCTA=0  no extra table-writing CTA in the fused gate + dispatch kernel (default 1; route.cu: launch_route)
int pdl_mask();      // bit 0 gate, bit 1 dispatch, bit 2 expert FFN, bit 3 LayerNorm: kernel launched with the PDL attribute
int pdl_trigger();   // same bits: kernel executes griddepcontrol.launch_dependents at its start
int prefetch_mode();
int ln_fuse_mode(); // B200MOE_LN_FUSE: 1 (default) = the block's norm_ff is folded into the route kernel when the caller supplies
                    // the pre-scaled router, 0 = always a row pass in front
int route_mode();   // B200MOE_ROUTE: 1 (default) = fused gate + dispatch kernel for small batches, 0 = separate kernels
constexpr int kPdlGate = 1, kPdlDispatch = 2, kPdlFfn = 4, kPdlLn = 8;

// Kernel launch with the programmatic-stream-serialization attribute: the kernel may become resident as soon as every
// CTA of the previous kernel in the stream has executed griddepcontrol.launch_dependents (or exited); everything it reads
// or writes that an earlier kernel produces or still reads must come after its own griddepcontrol.wait.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 int pdl_bit, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_bit & pdl_mask()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- small device helpers ------------------------------------------------------------------------------------
template <typename T>
struct IoType;
template <>
struct IoType<float> {
  static constexpr int code = B200MOE_F32;
};
template <>
struct IoType<__half> {
  static constexpr int code = B200MOE_F16;
};
template <>
struct IoType<bf16> {
  static constexpr int code = B200MOE_BF16;
};

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_float(bf16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_float(float v);
template <>
__device__ __forceinline__ float from_float<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ __half from_float<__half>(float v) {
  return __float2half_rn(v);
}
template <>
__device__ __forceinline__ bf16 from_float<bf16>(float v) {
  return __float2bfloat16_rn(v);
}

}  // namespace b200moe
