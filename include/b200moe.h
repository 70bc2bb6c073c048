/*
 * b200moe.h -- C ABI of the B200-native fast_moe expert layer (gate -> dispatch -> 32-expert FFN -> combine).
 *
 * This is the drop-in boundary for ONE hot path of LitLeo/3m-asr-inference: the work its TensorRT plugins
 * `SoftmaxTopKPluginDynamic` and `FMoEExpertPluginDynamic` do in `enqueue`, and the work trainer_3m_fix/fmoe's
 * FMoE.forward hands to the (un-vendored) `fmoe_cuda` extension.  Reference citations are relative to the
 * upstream tree (/root/reference in the build container):
 *
 *   b200moe_gate        <- TRTAPI++/plugin/softmax_topk_plugin/softmax_topk_plugin.cpp:88-127 (enqueue),
 *                          softmax_topk_kernel.cu:26-120, plus the router MatMul/Add that precede it in
 *                          trainer_3m_fix/layer/positionwise_feed_forward.py:169-207,225; and
 *                          trainer_3m_fix/fmoe/gates.py:51-66 (NaiveGate.forward) for gate_mode 1.
 *   b200moe_dispatch    <- fmoe_expert_kernel.cu:25-128 (ScatterMapping + ScatterMappingCopy) and
 *                          trainer_3m_fix/fmoe/functions.py:13-52,62-86 (moe_prepare_forward, MOEScatter.forward).
 *   b200moe_expert_ffn  <- fmoe_expert_plugin.cpp:80-128 (per-expert cublasSgemm/BiasSilu/cublasSgemm/Bias loop)
 *                          and functions.py:135-152 (MOEbiasLinear.forward), fmoe/transformer.py:22-30.
 *   b200moe_combine     <- fmoe_expert_kernel.cu:191-227 (GatherrMappingCopy), functions.py:175-199 (MOEGather),
 *                          positionwise_feed_forward.py:257-258 (x gate_value), fmoe/layers.py:204-206 (bmm with
 *                          the top-k scores), layer/fmoe_transformer.py:155-158 (x ff_scale, + residual).
 *   b200moe_forward     <- the whole of LocalFmoeCatEmbedFeedForward.forward + the scale/residual that follow it
 *                          (positionwise_feed_forward.py:209-265, fmoe_transformer.py:145-158).
 *   b200moe_plugin_*    <- FMoEExpertPlugin / FMoEExpertPluginCreator (fmoe_expert_plugin.h:31-132,
 *                          fmoe_expert_plugin.cpp:144-377): same creator fields, same six inputs, same
 *                          32-byte serialisation, same int status convention (0 = ok).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its comment says "host";
 *   - memory is caller-owned; no per-layer entry point allocates device memory, none synchronises the host, all
 *     work is enqueued on the caller's stream, so every one of them is CUDA-graph capturable (the one exception:
 *     the FIRST b200moe_plugin_enqueue of a plugin object allocates the plugin's own weight copies);
 *   - workspaces are pure scratch: they need no initialisation and nothing has to survive in them between calls;
 *   - return value: 0 = ok, negative = error (B200MOE_ERR_*); b200moe_last_error() gives the message of the
 *     calling thread's last failure;
 *   - there is no CPU fallback: on a machine without an sm_100 GPU the compute entry points return
 *     B200MOE_ERR_CUDA.
 */
#ifndef B200MOE_H_
#define B200MOE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Opaque CUDA stream (same type the CUDA runtime declares). */
#ifndef __DRIVER_TYPES_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define B200MOE_VERSION 100

/* Element types of activations at the boundary.  0/1 are the reference plugin's `data_type` field values
 * (fmoe_expert_plugin.cpp:325-366: 0 = fp32, 1 = fp16); 2 (bf16) is the native type of this implementation. */
enum { B200MOE_F32 = 0, B200MOE_F16 = 1, B200MOE_BF16 = 2 };

/* Expert activation (reference: act_type is parsed but SiLU is hard-coded, fmoe_expert_plugin.cpp:106). */
enum { B200MOE_ACT_SILU = 0, B200MOE_ACT_RELU = 1, B200MOE_ACT_GELU = 2 };

/* Gate flavours. 0: 3M-ASR router, softmax over all experts then max (top-1), value = max probability.
 * 1: FastMoE NaiveGate, top-k of the logits then softmax over the k selected logits. */
enum { B200MOE_GATE_3M = 0, B200MOE_GATE_NAIVE = 1 };

/* Arithmetic of the expert GEMMs. */
enum { B200MOE_COMPUTE_BF16 = 0, B200MOE_COMPUTE_TF32 = 1 };

enum {
  B200MOE_OK = 0,
  B200MOE_ERR_ARG = -1,      /* bad argument (null pointer, unsupported size or dtype) */
  B200MOE_ERR_CUDA = -2,     /* a CUDA runtime / driver call failed (message has the CUDA error string) */
  B200MOE_ERR_WORKSPACE = -3 /* workspace too small */
};

const char* b200moe_last_error(void);
/* Device-side status of the kernels launched so far on the current device (synchronous 4-byte read; call it at a point
 * where the stream is synchronised anyway): 0 = ok, 1 = a fused gate + dispatch launch gave up waiting for the other
 * CTAs of its grid (SMs held by a foreign kernel for > 2 s): that layer's output is undefined, nothing was written out
 * of range and the context stays usable.  clear != 0 resets the word.  (The reference's enqueue has the same int-status
 * contract but can only ever return 0: fmoe_expert_plugin.cpp:247-268, common/common.h:26-38.) */
int b200moe_status(int* host_status, int clear);
int b200moe_version(void);
/* 1 if device `dev` can run the kernels (compute capability 10.x), 0 if not, negative on CUDA error. */
int b200moe_device_supported(int dev);

/* ---- weight packing (once per checkpoint load) ------------------------------------------------------------
 * Expert weights stay in the reference's layout, FMoELinear.weight = [E, out, in] row-major
 * (trainer_3m_fix/fmoe/layers.py:34), which is already the K-major B operand the tensor cores want; packing
 * only casts them to bf16.  n = number of elements. */
int b200moe_pack_bf16(const void* src, int src_dtype, void* dst_bf16, size_t n, cudaStream_t stream);
/* For B200MOE_COMPUTE_TF32: the fp32 weights rounded to the nearest TF32 value (10-bit mantissa), still fp32 words in
 * the same [E, out, in] layout; dst may alias src.  tcgen05 kind::tf32 truncates whatever fp32 words it reads, which
 * on un-rounded weights leaves a systematic ~1e-3 relative bias. */
int b200moe_pack_tf32(const float* src, float* dst, size_t n, cudaStream_t stream);

/* ---- workspace --------------------------------------------------------------------------------------------
 * Bytes of scratch b200moe_forward / b200moe_plugin_enqueue / the staged entry points need for S tokens.
 * (Reference: FMoEExpertPlugin::getWorkspaceSize, fmoe_expert_plugin.cpp:224-239.) */
size_t b200moe_workspace_bytes(int S, int E, int D, int H, int top_k);

/* ---- stage 1: gate -------------------------------------------------------------------------------------------
 * logits = cat(embed, x) . Wr (+ br); embed may be NULL (then R = D).  x [B*T, D], embed [B*T, Demb] in `dtype`;
 * Wr [Demb + D, E] fp32 row-major (router_weights, positionwise_feed_forward.py:134-135), br [E] fp32 or NULL.
 * x_len [B] int32 or NULL: rows t >= x_len[b] are padding; they get idx = -1, score = 0 (the reference leaves
 * them unwritten, softmax_topk_kernel.cu:40).  Ties resolve to the lowest expert index.
 * Outputs: idx [B*T, top_k] int32, score [B*T, top_k] fp32 (top-k ordered by descending logit). */
int b200moe_gate(const void* x, const void* embed, const float* Wr, const float* br, const int* x_len, int B, int T,
                 int D, int Demb, int E, int top_k, int gate_mode, int dtype, int* idx, float* score,
                 cudaStream_t stream);

/* Tensor-core variant of the gate for bf16 activations and E <= 32 (D, Demb multiples of 64).  The router is packed
 * once per checkpoint load into bf16 [64, Demb + D]: rows 0..31 = bf16(Wr^T), rows 32..63 = bf16(Wr^T - rows 0..31), so
 * that the fp32 router weights survive to ~2^-17 and bf16 x bf16 products stay exact in the fp32 accumulator.
 * Same outputs and tie rule as b200moe_gate. */
size_t b200moe_router_pack_bytes(int R);
int b200moe_pack_router(const float* Wr, int R, int E, void* packed, cudaStream_t stream);
int b200moe_gate_tc(const void* x, const void* embed, const void* Wr_packed, const float* br, const int* x_len, int B,
                    int T, int D, int Demb, int E, int top_k, int gate_mode, int* idx, float* score,
                    cudaStream_t stream);

/* ---- stage 2: dispatch -------------------------------------------------------------------------------------
 * Counts tokens per expert, exclusive-scans, and stably scatters token rows into expert-contiguous order:
 *   mapping[i] = offsets[idx[i]] + #{ j < i : idx[j] == idx[i] }   for entry i = s*top_k + j; -1 if idx[i] < 0
 *   xbuf[mapping[i], :] = bf16(x[i / top_k, :])
 * counts [E], offsets [E+1] (= the reference's acc_histogram), mapping [S*top_k], all int32.
 * xbuf [S*top_k, D] bf16.  ws = scratch of at least b200moe_workspace_bytes(). */
int b200moe_dispatch(const void* x, const int* idx, int S, int D, int E, int top_k, int dtype, int* counts,
                     int* offsets, int* mapping, void* xbuf, void* ws, cudaStream_t stream);

/* ---- stage 3: expert FFN ----------------------------------------------------------------------------------
 * For every expert e and its rows [offsets[e], offsets[e+1]) of xbuf:
 *   ybuf = act(xbuf . W1[e]^T + b1[e]) . W2[e]^T + b2[e]
 * W1 [E, H, D], W2 [E, D, H] bf16 (b200moe_pack_bf16 of the reference tensors), b1 [E, H], b2 [E, D] fp32
 * (NULL = no bias).  n_rows = S*top_k (capacity of xbuf / ybuf).  ybuf [n_rows, D] in `out_dtype`.
 * One persistent tcgen05/TMEM/TMA kernel; no host synchronisation (the reference syncs twice and launches up to
 * 128 kernels on 8 streams, fmoe_expert_plugin.cpp:75-130). */
int b200moe_expert_ffn(const void* xbuf, const int* offsets, int n_rows, const void* W1, const float* b1,
                       const void* W2, const float* b2, int E, int D, int H, int act_type, int out_dtype, void* ybuf,
                       void* ws, cudaStream_t stream);

/* ---- stage 4: combine --------------------------------------------------------------------------------------
 * out[s, :] = (residual ? residual[s, :] : 0) + ff_scale * sum_j (score ? score[s, j] : 1) * ybuf[mapping[s*k+j], :]
 * Entries with mapping < 0 contribute nothing.  score NULL = the reference's keep_expert_output.
 * ybuf / residual / out are in `dtype`. */
int b200moe_combine(const void* ybuf, const int* mapping, const float* score, const void* residual, float ff_scale,
                    int S, int D, int top_k, int dtype, void* out, cudaStream_t stream);

/* ---- the staged operator surface of trainer_3m_fix/fmoe/functions.py (forward halves) --------------------------
 * What the reference's autograd Functions hand to the un-vendored `fmoe_cuda` extension, one entry point each, so that
 * MOEScatter / MOELinear / MOEbiasLinear / MOEGather `.apply` (3m-asr-inference_b200/fmoe/functions.py) are thin.
 *
 * b200moe_prepare      <- moe_prepare_forward, functions.py:13-52: routing tables only, no row copy, no host sync.
 *   idx [S*top_k] int32 target experts (-1 = not routed).  counts [E], offsets [E+1], mapping [S*top_k] as
 *   b200moe_dispatch; pos [S*top_k] or NULL = the reference's `pos` (the STABLE argsort of idx: pos[mapping[i]] = i,
 *   entries past offsets[E] unspecified).  ws >= b200moe_workspace_bytes(S, E, 8, 0, top_k).
 * b200moe_scatter_rows <- fmoe_cuda.local_gather, functions.py:194 (and the backward of local_scatter :102):
 *   out[index[i], :] = in[i, :] for i < n, rows of D elements in `dtype`; index values outside [0, n_out) are skipped.
 *   (fmoe_cuda.local_scatter, functions.py:72, is the gather out[i] = in[pos[i]]: b200moe_combine with score = residual
 *   = NULL, ff_scale = 1, top_k = 1.)
 * b200moe_expert_linear <- fmoe_cuda.forward, functions.py:113-116,141-149 (MOELinear / MOEbiasLinear.forward): ONE
 *   grouped linear, out[r, :] = act(xbuf[r, :] . W[e]^T + bias[e]) for the rows r of expert e.  xbuf [n_rows, K] bf16,
 *   W [E, N, K] bf16, bias [E, N] fp32 or NULL, out [n_rows, N] bf16; act_type 0..2 as B200MOE_ACT_*, 3 = none (the
 *   reference applies its activation between the two Functions, fmoe/transformer.py:27-29).  K, N multiples of 128.
 *   The first GEMM of the expert kernel on its own.  ws >= b200moe_workspace_bytes(n_rows, E, K, 0, 1). */
int b200moe_prepare(const int* idx, int S, int E, int top_k, int* counts, int* offsets, int* mapping, int* pos,
                    void* ws, cudaStream_t stream);
int b200moe_scatter_rows(const void* in, const int* index, int n, int n_out, int D, int dtype, void* out,
                         cudaStream_t stream);
int b200moe_expert_linear(const void* xbuf, const int* offsets, int n_rows, const void* W, const float* bias, int E,
                          int K, int N, int act_type, void* out_bf16, void* ws, cudaStream_t stream);

/* ---- the fused layer --------------------------------------------------------------------------------------- */
typedef struct b200moe_layer_args {
  /* activations, `dtype` */
  const void* x;        /* [B*T, D]  MoE input (already layer-normed by the caller)                  */
  const void* embed;    /* [B*T, Demb] or NULL                                                        */
  const void* residual; /* [B*T, D] or NULL                                                           */
  void* out;            /* [B*T, D]                                                                    */
  const int* x_len;     /* [B] or NULL                                                                 */
  /* router */
  const float* Wr;      /* [Demb + D, E] fp32 (may be NULL when Wr_packed is given and usable)        */
  const void* Wr_packed;/* b200moe_pack_router(Wr) or NULL: enables the tensor-core gate for bf16 / E <= 32 */
  const float* br;      /* [E] fp32 or NULL                                                            */
  /* experts (bf16 packed weights, fp32 biases) */
  const void* W1;       /* [E, H, D] bf16                                                              */
  const float* b1;      /* [E, H] or NULL                                                              */
  const void* W2;       /* [E, D, H] bf16                                                              */
  const float* b2;      /* [E, D] or NULL                                                              */
  int B, T, D, Demb, E, H, top_k;
  int gate_mode, act_type, dtype;
  int keep_expert_output; /* 1: do not multiply by the gate value (positionwise_feed_forward.py:257)  */
  float ff_scale;         /* 0.5 for the macaron Conformer wiring (fmoe_transformer.py:56-58,155)     */
  /* optional routing outputs for parity checks (NULL to skip): idx/score [B*T, top_k], counts [E],
   * mapping [B*T*top_k] */
  int* idx_out;
  float* score_out;
  int* counts_out;
  int* mapping_out;
  /* Arithmetic of the expert GEMMs.  B200MOE_COMPUTE_BF16 (0, default): W1 / W2 are the bf16-packed weights.
   * B200MOE_COMPUTE_TF32 (1): W1 / W2 are the reference's fp32 FMoELinear weights ([E, H, D] / [E, D, H]) after
   * b200moe_pack_tf32, activations must be fp32 (`dtype` 0), the tensor cores run tcgen05 kind::tf32 with fp32
   * accumulation and fp32 intermediates: outputs within rel-L2 1e-3 of the fp32 reference instead of 1e-2. */
  int compute;
} b200moe_layer_args;

int b200moe_forward(const b200moe_layer_args* args, void* ws, size_t ws_bytes, cudaStream_t stream);

/* ---- expert parallelism over peer-mapped memory (NVLink / NVSwitch) ---------------------------------------------
 * The reference's only multi-GPU mode of this path: experts are partitioned contiguously over the GPUs of one node
 * (expert e on rank e / E_local; `num_expert` is per worker: trainer_3m_fix/model/..._hier.py:259-273), tokens stay
 * data-parallel, and fmoe_cuda exchanges counts and rows with all-to-alls around the expert computation
 * (trainer_3m_fix/fmoe/functions.py:37-50 expert_exchange + .cpu(), :74-80 global_scatter, :185-191 global_gather).
 * Here the exchange is fused into the kernels on either side of it, over peer-mapped memory: one CTA of the dispatch
 * kernel stores this rank's per-expert counts into every rank; every CTA waits for all ranks' counts and stores its
 * token rows straight to their final place in the owner GPU's receive buffer (expert-major, so that each local expert's
 * rows are contiguous whatever rank they came from and the expert kernel runs full token tiles); the expert-FFN kernel's
 * second GEMM stores every result row straight to its source GPU; arrival is signalled with system-scope
 * release/acquire flags -- no collective call, no host synchronisation, CUDA-graph capturable.  One process per GPU;
 * the buffers are exchanged once at set-up:
 *
 *   every rank:  b200moe_ep_alloc(b200moe_ep_buffer_bytes(world, E_local, D, cap), &buf);   (cudaMalloc, zeroed)
 *                b200moe_ep_ipc_export(buf, handle);   all-gather the 64-byte handles by any means (host side)
 *                b200moe_ep_ipc_open(handle_of_rank_r, &bufs[r]) for r != rank;  bufs[rank] = buf
 *                ctx = b200moe_ep_create(rank, world, E_local, D, cap, bufs, timeout_ms);
 *   per layer:   b200moe_ep_forward(ctx, &args, ws, ws_bytes, stream);
 *
 * cap = the largest B*T*top_k any rank will ever pass.  args->E is the TOTAL number of experts (router width),
 * args->W1/b1/W2/b2 hold this rank's E_local experts.  Activations must be bf16.  Every rank must call
 * b200moe_ep_forward the same number of times with the same kind of arguments (ranks without tokens pass B*T = 0).  A
 * peer that does not show up within timeout_ms makes the waiting kernels give up, sets the status word
 * (b200moe_ep_status != 0) and poisons the layer's output with NaNs.
 *
 * Folded combine.  When the call is top-1, args->residual is NULL or args->x itself, there is no norm_final, and
 * args->out lies inside this rank's symmetric buffer (b200moe_ep_out_buffer: two slots, to be used alternately, the same
 * slot on every rank), the owner's epilogue writes the finished rows  residual + ff_scale * score * y  directly into the
 * source rank's `out` and no combine kernel runs at all (the reference's MOEGather + weighting + residual add,
 * functions.py:185-194, positionwise_feed_forward.py:257-258).  b200moe_ep_forward then ends with a one-CTA kernel that
 * waits for the owners' flags; b200moe_ep_forward_deferred leaves that wait to the next b200moe_ep_forward* call on the
 * context (which does it first thing) or to an explicit b200moe_ep_wait -- nothing else may touch `out` before. */
typedef struct b200moe_ep_ctx b200moe_ep_ctx;
size_t b200moe_ep_buffer_bytes(int world, int E_local, int D, int cap);
int b200moe_ep_alloc(size_t bytes, void** dev_ptr);
int b200moe_ep_free(void* dev_ptr);
int b200moe_ep_ipc_export(void* dev_ptr, void* handle64 /* host, 64 bytes */);
int b200moe_ep_ipc_open(const void* handle64 /* host */, void** peer_ptr);
int b200moe_ep_ipc_close(void* peer_ptr);
b200moe_ep_ctx* b200moe_ep_create(int rank, int world, int E_local, int D, int cap, void* const* bufs /* host [world] */,
                                  int timeout_ms);
void b200moe_ep_destroy(b200moe_ep_ctx* ctx);
size_t b200moe_ep_workspace_bytes(const b200moe_ep_ctx* ctx, int H);
int b200moe_ep_forward(b200moe_ep_ctx* ctx, const b200moe_layer_args* args, void* ws, size_t ws_bytes,
                       cudaStream_t stream);
int b200moe_ep_forward_deferred(b200moe_ep_ctx* ctx, const b200moe_layer_args* args, void* ws, size_t ws_bytes,
                                cudaStream_t stream);
int b200moe_ep_wait(b200moe_ep_ctx* ctx, cudaStream_t stream);
/* Output slot 0 / 1 inside this rank's symmetric buffer: [cap, D] bf16 each. */
int b200moe_ep_out_buffer(const b200moe_ep_ctx* ctx, int slot, void** dev_ptr, size_t* bytes);
/* The same layer, driven in stages (bit mask: 1 gate + this rank's counts to every rank, 8 wait for the counts + push
 * the rows, 2 wait for the rows + expert FFN + push back, 4 wait for the results (+ combine); 15 = b200moe_ep_forward).
 * Lets one process drive several ranks on ONE GPU stage by stage (tests): kernels that wait on one another must never
 * be queued on the same GPU in the wrong order. */
int b200moe_ep_forward_stages(b200moe_ep_ctx* ctx, const b200moe_layer_args* args, void* ws, size_t ws_bytes,
                              int stages, cudaStream_t stream);
/* Reads the context's device status word (synchronous copy): 0 ok, 1 a peer's rows did not arrive, 2 returned rows did
 * not arrive, 3 a peer's counts did not arrive, 4 the ranks disagree on folding the combine. */
int b200moe_ep_status(const b200moe_ep_ctx* ctx, int* host_status);

/* ---- plugin object: mirror of FMoEExpertPlugin / FMoEExpertPluginCreator --------------------------------------
 * Fields are the creator's (fmoe_expert_plugin.cpp:325-366): data_type (0 fp32, 1 fp16; anything else fails like
 * the reference creator, except 2 = bf16 which is new), num_expert, idim, hidden_units, act_type. */
typedef struct b200moe_plugin b200moe_plugin;

b200moe_plugin* b200moe_plugin_create(int data_type, int num_expert, int idim, int hidden_units, int act_type);
b200moe_plugin* b200moe_plugin_clone(const b200moe_plugin* p);
/* 32 bytes: data_type, num_expert, idim, hidden_units, act_type + 3 zero ints (fmoe_expert_plugin.cpp:288-304). */
size_t b200moe_plugin_serialization_size(const b200moe_plugin* p);
int b200moe_plugin_serialize(const b200moe_plugin* p, void* host_buffer);
b200moe_plugin* b200moe_plugin_deserialize(const void* host_data, size_t length);
void b200moe_plugin_destroy(b200moe_plugin* p);
/* getWorkspaceSize (fmoe_expert_plugin.cpp:224-239): scratch for S tokens.  Pure scratch: nothing is expected to survive
 * in it between two enqueues, so TensorRT may share it between layers and pass a different S every time. */
size_t b200moe_plugin_workspace_bytes(const b200moe_plugin* p, int S);
/* enqueue: the reference's six inputs in the reference's order -- input [S, idim], gate_idx [S] int32,
 * w1_weight [E, H, D], w1_bias [E, H], w2_weight [E, D, H], w2_bias [E, D] (all `data_type` except gate_idx) --
 * and its single un-weighted output [S, idim].  The bf16 / fp32 copies of the weights live in memory the plugin owns
 * (cudaMalloc by the first enqueue -- which therefore cannot be captured into a CUDA graph -- freed by destroy) and are
 * re-made only when one of the four weight pointers changes, or after b200moe_plugin_invalidate (weights updated in
 * place behind the same pointers, e.g. a TensorRT refit). */
int b200moe_plugin_enqueue(b200moe_plugin* p, const void* input, const int* gate_idx, const void* w1_weight,
                           const void* w1_bias, const void* w2_weight, const void* w2_bias, int S, void* output,
                           void* workspace, size_t workspace_bytes, cudaStream_t stream);
int b200moe_plugin_invalidate(b200moe_plugin* p);

/* ---- companion gate plugin: SoftmaxTopKPluginDynamic (softmax_topk_plugin.cpp:88-136) --------------------------
 * logits [B, T, E] in `data_type`, mask [B] int32 (valid lengths) -> value [B, T] (`data_type`), idx [B, T]
 * int32; top-1 only, value = 1 / sum(exp(l - max)).  Rows >= mask[b] get idx -1 / value 0. */
int b200moe_softmax_topk_enqueue(const void* logits, const int* mask, int B, int T, int E, int data_type,
                                 void* value, int* idx, cudaStream_t stream);

/* ---- the Conformer block's feed-forward part: LayerNorms either side of the layer --------
 * Replaces, in trainer_3m_fix/layer/fmoe_transformer.py:144-166 (FmoeConformerLayer.forward, normalize_before = True):
 *   :145-148  residual = x;  x = addLayerNorm(norm_ff, x)            -> norm_ff_gamma / norm_ff_beta
 *   :152      x = feed_forward(x, embed, x_len)                       -> the layer (b200moe_forward)
 *   :155-158  x = residual + ff_scale * x                             -> layer.residual / layer.ff_scale
 *   :164-166  x = addLayerNorm(norm_final, x)   (blocks with a conv module) -> norm_final_gamma / norm_final_beta
 * i.e.  out = norm_final( residual + ff_scale * MoE( norm_ff(x), embed ) ).
 * `layer.x` is the block input BEFORE norm_ff (normally also `layer.residual`); gamma / beta are fp32 [D]; a NULL gamma
 * skips that norm; eps is shared (the reference uses nn.LayerNorm(size, eps=1e-12) for both, :54-65).  LayerNorm
 * semantics: biased variance over the D features, y = (x - mean) * rsqrt(var + eps) * gamma + beta, statistics in fp32,
 * the normalised input rounded to `dtype` (what the reference's LayerNorm plugin hands to the next graph layer).
 * D: multiple of 8, at most 1024.  ws: b200moe_block_workspace_bytes (the layer's workspace + one [S, D] buffer).
 * Where the norms run.  norm_ff: as a row pass in front of the gate; or, when Wr_packed_ln is given (bf16 batches of up
 * to 2 * 32 * 148 tokens), folded into the fused gate + dispatch kernel algebraically: with mu, r the mean and
 * reciprocal standard deviation of a row,  LN(x) . Wr_x = r * (x . diag(gamma) Wr_x - mu * gamma^T Wr_x) + beta^T Wr_x,
 * so the router MMAs run on the raw rows against the pre-scaled router while idle warps compute mu and r, the arg-max
 * epilogue combines them, and the rows handed to the experts are normalised (and rounded to bf16) as they are scattered;
 * `x` itself stays as the residual.  The router then sees norm_ff(x) in fp32 rather than rounded to bf16 -- closer to
 * the fp32 reference; tokens whose two best logits differ by less than that rounding may route differently between the
 * two variants.  norm_final: inside the combine kernel when there is one (top_k > 1, expert parallelism), otherwise
 * as a row pass over `out` behind the expert kernel.  All of them share one register-level routine. */
typedef struct b200moe_block_args {
  b200moe_layer_args layer;
  const float* norm_ff_gamma;
  const float* norm_ff_beta;
  const float* norm_final_gamma;
  const float* norm_final_beta;
  float eps;
  /* optional: b200moe_pack_router_ln(Wr, gamma = norm_ff_gamma, beta = norm_ff_beta); lets norm_ff be folded into the
   * fused gate + dispatch kernel (see below); NULL = norm_ff as a row pass in front */
  const void* Wr_packed_ln;
} b200moe_block_args;

size_t b200moe_block_workspace_bytes(int S, int E, int D, int H, int top_k);
int b200moe_block_forward(const b200moe_block_args* args, void* ws, size_t ws_bytes, cudaStream_t stream);
/* The same block with the experts partitioned over the GPUs of `ctx` (see above). */
size_t b200moe_ep_block_workspace_bytes(const b200moe_ep_ctx* ctx, int H);
int b200moe_ep_block_forward(b200moe_ep_ctx* ctx, const b200moe_block_args* args, void* ws, size_t ws_bytes,
                             cudaStream_t stream);
/* Router packing for the folded norm_ff: like b200moe_pack_router with the last D rows of Wr [R, E] (the x part of the
 * concat) scaled by gamma, followed by c1 = gamma^T Wr_x and c0 = beta^T Wr_x (32 floats each). */
size_t b200moe_router_ln_pack_bytes(int R);
int b200moe_pack_router_ln(const float* Wr, int R, int E, int D, const float* gamma, const float* beta, void* packed,
                           cudaStream_t stream);
/* ---- the other encoder plugins: what sits between the fast_moe blocks in the reference's TensorRT graph ----
 * All in `dtype` (B200MOE_F32 / F16 / BF16) with fp32 arithmetic, caller's stream, in == out allowed.
 * b200moe_att_masked_softmax <- AttMaskedSoftmaxPluginDynamic (TRTAPI++/plugin/att_masked_softmax_plugin/
 *   att_masked_softmax_kernel.cu:198-277; called from layer/attention.py:199-239): in / out [B, N, S, ld] attention scores,
 *   mask [B] int32 valid keys (NULL = all); out = softmax(scale * in) over keys < mask[b], exactly 0 behind; ld <= 1024.
 *   A row with no valid key gives zeros (the reference divides by zero there).
 * b200moe_glu <- GluPluginDynamic (glu_plugin/glu_kernel.cu:24-37; layer/convolution.py:125): x [M, 2C, N] -> y [M, C, N],
 *   y = x[:, :C] * sigmoid(x[:, C:]).
 * b200moe_masked_fill <- MaskedFillPluginDynamic (masked_fill_plugin/masked_fill_kernel.cu:25-39; layer/convolution.py:
 *   89-113,153): in / out [B, dim, T], positions t >= mask[b] become `fill`.
 * b200moe_rel_pos_encoding <- RelPositionalEncodingPluginDynamic (rel_positional_encoding_plugin/
 *   rel_positional_encoding_kernel.cu:61-93; layer/positional_encoding.py:101-130): out [B, T, D] = in * scale,
 *   pos_emb [T, D] = pe[:T] (pe [max_len, D], max_len >= T). */
int b200moe_att_masked_softmax(const void* in, const int* mask, float scale, int B, int N, int S, int ld, int dtype,
                               void* out, cudaStream_t stream);
int b200moe_glu(const void* x, int M, int C, int N, int dtype, void* y, cudaStream_t stream);
int b200moe_masked_fill(const void* in, const int* mask, float fill, int B, int dim, int T, int dtype, void* out,
                        cudaStream_t stream);
int b200moe_rel_pos_encoding(const void* in, const void* pe, float scale, int B, int T, int D, int dtype, void* out,
                             void* pos_emb, cudaStream_t stream);

/* The LayerNorm alone (LayerNormPluginDynamic, TRTAPI++/plugin/layer_norm_plugin/layer_norm_kernel.cu, with the eps the
 * plugin drops): out [S, D] = LN(in [S, D]); in == out allowed. */
int b200moe_layernorm(const void* in, const float* gamma, const float* beta, float eps, int S, int D, int dtype,
                      void* out, cudaStream_t stream);

/* Run-time tunables (each also has an environment default, B200MOE_<KEY>): "route" 1/0 fused gate + dispatch kernel for
 * small batches; "pdl" / "pdl_trig" bit masks (1 gate, 2 dispatch, 4 expert FFN, 8 LayerNorm) for programmatic dependent launch;
 * "prefetch" 0/1/2 L2 prefetch of the expert weights from the gate kernel; "ln_fuse" 1/0 fold the block's norm_ff into
 * the fused gate + dispatch kernel when Wr_packed_ln is given, or always run it as a row pass in front.  Results do not depend on any of them. */
int b200moe_config(const char* key, int value);

/* Optional per-stage device timing with CUDA events around each stage of b200moe_forward (eager launches only, not
 * under graph capture). stage_ms / stage_calls are HOST arrays of 4: 0 gate, 1 dispatch, 2 expert_ffn (with the fused
 * combine epilogue when top_k == 1), 3 combine.  b200moe_profile_read waits for the recorded work and resets. */
int b200moe_profile_enable(int on);
int b200moe_profile_read(float* stage_ms, int* stage_calls);

/* Debug: while dev_buf is non-NULL every expert-FFN launch runs its tracing variant and writes a per-CTA timeline of
 * 16-byte records {tile, event, clock64 lo, hi}: 4 roles (TMA producer, MMA issuer, epilogue, publisher) x
 * records_per_cta/4 slots per CTA, CTA-major.  dev_buf must hold 148 * records_per_cta records.  See tools/ffn_trace.py. */
int b200moe_debug_ffn_trace(void* dev_buf, int records_per_cta);

/* Debug, across kernels: while dev_buf is non-NULL every launch of the fused gate + dispatch kernel and of the expert
 * kernel takes the next slot of 148 CTAs x 8 marks of 8 bytes (%globaltimer in ns; 0 = not reached): 0 CTA start,
 * 1 dependency wait passed, 2 first operand data on chip, 3 last weight request issued, 4 last MMA issued, 5 CTA end.
 * dev_buf holds max_launches slots, zeroed by the caller; slots are handed out in launch order (and frozen into a
 * captured CUDA graph).  b200moe_debug_timeline_kind(slot) -> 1 route, 2 expert kernel, 0 unused.  tools/timeline.py. */
int b200moe_debug_timeline(void* dev_buf, int max_launches);
int b200moe_debug_timeline_kind(int slot);

/* Debug: timeline of the fused gate + dispatch kernel, 16 records of 16 bytes per CTA (148 CTAs at most). */
int b200moe_debug_route_trace(void* dev_buf);

/* Number of kernels this library launched on behalf of the calling process (for bench accounting). */
unsigned long long b200moe_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200MOE_H_ */
