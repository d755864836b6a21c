// Engine behind the C ABI of include/mocr_b200.h: weights, device workspace, launch
// sequences for preprocess -> ViT encoder -> cross-K/V projection -> batched greedy decode.
//
// The path it replaces:
//   upstream MangaOcr.__call__  -> ViTImageProcessor  -> VisionEncoderDecoderModel.generate
//   (transformers/models/vit/modeling_vit.py:428-458, models/bert/modeling_bert.py:856-910,
//    generation/utils.py:2743-2805), called from reference/src/ui/main_window.py:9801.
//
// Everything on the device is a hand-written sm_100a kernel from this directory; there is
// no library GEMM and no CPU fallback: without a B200 every compute entry point fails.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <chrono>
#include <vector>

#include "../../include/mocr_b200.h"
#include "common.cuh"
#include "decode_stages.cuh"
#include "encoder_attn_tc.cuh"
#include "gemm_tcgen05.cuh"
#include "gemm_ksplit.cuh"
#include "preprocess.cuh"
#include "region_mask.cuh"
#include "beam_search.h"
#include "beam_kernels.cuh"
#include "beam_device.cuh"
#include "rowops.cuh"

using namespace mocr;

namespace {

constexpr int kClsId = 2, kSepId = 3;   // [CLS] starts a sequence, [SEP] ends it ([PAD] = 0 fills)
constexpr float kQScale = 0.125f;   // 1/sqrt(64), folded into every query projection (exact in bf16)

thread_local std::string g_create_error;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

// A Linear layer on the device: W [N, K] bf16 (K-major = the UMMA B operand as stored by
// nn.Linear), bias [N] fp32, and one TMA descriptor per N-tile width in use.
struct Linear {
  __nv_bfloat16* w = nullptr;
  float* bias = nullptr;
  int N = 0, K = 0;
  std::map<int, CUtensorMap> maps;
};

// A bf16 activation buffer [rows_cap, K] used as the UMMA A operand.
struct ActBuf {
  __nv_bfloat16* p = nullptr;
  int rows_cap = 0, K = 0;
  CUtensorMap map;
};

struct LnParams {
  float* g = nullptr;
  float* b = nullptr;
};

struct EncLayer {
  LnParams ln1, ln2;
  Linear qkv, out, fc1, fc2;
};
struct DecLayer {
  Linear self_qkv, self_out, cross_q, cross_out, fc1, fc2;
  LnParams ln_self, ln_cross, ln_ffn;
};

}  // namespace

struct mocr_handle {
  int device = 0;
  int max_batch = 0;
  int max_length = 0;
  int sms = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t stream_enc = nullptr;  // lower priority: encoder of the crops that are still waiting while the decoder already runs (slot refill)
  cudaStream_t stream_enc_hi = nullptr;   // the same at the decoder's priority (option pipeline = 2; admissions with sess_enc_hi = 1)
  cudaStream_t stream_fetch = nullptr;    // finished id rows of a session (mocr_session_fetch)
  int* d_sess_map = nullptr;              // [64] cache block of each crop of the admission being encoded
  const int* enc_crop_map = nullptr;      // set while an admission's encoder pass is launched
  cudaEvent_t ev_published = nullptr;     // the last admission's crops are in the device queue
  bool sess_wait_pub = false;             // the next chunk starts after that publication (the session was empty: no row to hold up)
  cudaEvent_t ev_staged = nullptr;        // the last admission's pixels and descriptors have been consumed (its preprocess has run)
  int sess_enc_hi = 1;                    // admissions of a session run at the decoder's stream priority
  cudaEvent_t ev_first = nullptr;     // first sub-chunk encoded and published
  std::mutex mu;
  std::string error;
  int64_t launches = 0;
  int taps = 0;
  bool finalized = false;
  std::vector<void*> allocs;
  std::map<std::string, HostTensor> staged;

  // tile widths (mocr_set_option)
  int enc_bn = 256;
  int dec_tc = 1;           // decoder GEMM stages on the encoder's tcgen05 kernel (bit 0: vocabulary, bit 1: FFN1, bit 2: QKV)
  int big_row_warps = 2;    // the same in the large-batch program (512 rows: 8 -> 333.0, 4 -> 330.4, 2 -> 329.5 us per step)
  int row_warps = 2;        // warps (= rows) per CTA of the decoder's row-wise stage kernels
  int carveout = -1;        // shared-memory carve-out (percent) forced on every decoder stage kernel; -1: driver default (set before the first decode)
  int kv_evict_first = -1;  // L2 evict-first hint on the decoder's K/V reads: bit 0 encoder K/V, bit 1 self-attention cache; -1 = by program:
                            // small 2 (the encoder K/V of <= 64 rows stay in L2 with the weights: 98.7 -> 94.6 us per step at 32 rows,
                            // 119.0 -> 118.6 at 64), large 3 (everything streams: 347.5 -> 344.1 at 512 rows)
  int resid_tma = 1;        // encoder residual adds through the TMA reduce-add epilogue (0: per-thread f32 loads/stores)
  int enc_bn768 = 256;      // tile width of the N = 768 encoder GEMMs (128 or 256; 192 would give 2.68 waves instead of 2.007, measured 3 % slower)
  int check_every = 26;
  int use_graph = 1;
  int use_pdl = 1;          // programmatic dependent launch between the decoder's stage kernels
  int pdl_mask = 0x1ff;     // ... per stage type (bit = PdStageType): kernels of a type whose bit is clear are launched in plain stream order
  int pdl_now = 1;          // (set per launch by decode_stage_step)
  int steps_per_graph = 13; // decode steps captured in one CUDA graph (299 = 23 x 13)
  int fuse_ln = 1;          // decoder projections that feed a LayerNorm as 16-CTA clusters that normalise the rows themselves (0: split-K partials + LayerNorm stage)
  int kv_prefetch = 0;      // 1: a layer's encoder K/V are requested into L2 by the layer's first stage (bulk prefetch before the dependency wait); measured 2.7 us per step SLOWER at 64 rows (the prefetch competes with the weights for L2)
#ifdef MOCR_GEMM_DBG
  int gemm_dbg = 0;
#endif
  int big_ksplit = 1;       // ... and split their K dimension over clusters of four CTAs (gemm_ksplit.cuh) while tiles x 4 <= SMs
  int big_accum = 1;        // large-batch program: the residual projections accumulate into x in place (EPI_F32_ACCUM) instead of y = x + ...
  // ---- session (admission into a running decode; mocr_session_*)
  bool sess_on = false;
  int sess_T = 0, sess_rows = 0, sess_order = 0;
  long long sess_published = 0;          // publications so far (the device queue counts them)
  std::vector<char> sess_used;           // slot occupied (published, result not yet released)
  PdParams sess_p{};
  PdParams sess_p_small{};                // the same session on its first kSessSmallRows rows only (mocr_session_rows)
  cudaGraphExec_t sess_exec_small = nullptr;
  int64_t sess_per_step_small = 0;
  int sess_rows_now = 0;                  // rows the following chunks step
  cudaGraphExec_t sess_exec = nullptr;
  int64_t sess_per_step = 0;
  int* d_ring = nullptr;                 // [max_batch] slots in publication order
  int* h_sess_lens[2] = {nullptr, nullptr};   // pinned [max_batch]: two length snapshots in flight
  cudaEvent_t sess_ev[2] = {nullptr, nullptr};
  long long snap_enq = 0, snap_read = 0;      // snapshots enqueued / handed to the caller
  std::vector<long long> sess_min_snap;       // per slot: the first snapshot that may speak for its current occupant
  int stage_chunk = 128;    // crops per staging chunk of a large batch (0 = one piece): see stage_encode
  int enc_tma_store = 1;    // encoder GEMMs with bf16 outputs (QKV, FFN1): rows leave as TMA stores from a staging tile (GemmArgs::out_tma)
  int big_attn_rows = 1;    // large-batch program: attention as one WARP per (row, head) unit (pd_attention_rows_kernel); 0 = the four-warp kernel
  int big_attn_grid = 384;  // CTAs of the attention stages in the large-batch program (0 = as many as there is work for); warp-per-unit kernel at 512 rows: 384 -> 348 us per step, 444 -> 355, 296 -> 375, 512 -> 405
  int big_bn768 = 32;       // tile width of the large-batch program's N = 768 GEMMs (32 or 64)
  int big_vocab_bn = 0;     // tile width of its vocabulary projection: 64, 128, 256, or 0 = by row count (about 96 tiles: 64 columns per 128-row
                            // m-tile; us per step 64 / 128 / 256: 128 rows 178.6 / 180.7 / 183.0, 256 rows 227.3 / 224.7 / 227.5, 512 rows 317.5 / 314.1 / 312.6)
  int big_rows = 112;       // more decoder rows than this: the large-batch program (every Linear on the tcgen05 GEMM, 128-row tiles).  Measured us per step,
                            // small / large program: 96 rows 157 / 178, 112 rows 184 / 182, 128 rows 199 / 188, 144 rows 244 / 216, 160 rows 248 / 220
  int pipeline = 0;         // with slot refill, 1 (2: equal stream priorities): encode the waiting crops in sub-chunks on a second stream while the decoder
                            // already runs.  Measured on the ragged 512-crop leg: 116 ms (80 ms at equal priorities) against 75 ms with the encoder
                            // serialised in front - the 200 KB GEMM CTAs and the decoder's stage kernels do not share SMs well - so it is off
  double prof_add[4] = {0, 0, 0, 0};   // MOCR_SESSION_PROF: seconds in mocr_session_add (wait for the previous pass, staging, launches) and passes
  int sub_i0 = 0, sub_n = 0;   // sub-range of the staged crops that preprocess / encode work on (sub_n = 0: all of them)
  int beam_device = 1;      // beam search with the selection on the device and the steps in a CUDA graph (0: host bookkeeping, one round trip per step)
  int beam_steps_per_graph = 8;
  int stage_threads = 6;    // host threads that copy a large batch of crops into the pinned arena (one per 8 MB, at most this many)
  int slots = 0;            // decoder rows of a greedy decode (0 = one per crop).  With fewer rows than crops a row that finishes takes
                            // the next waiting crop (in-flight slot refill): worth it when lengths are ragged (real text); with random-init
                            // weights, which never emit EOS, one row per crop on the large-batch program is faster
  int attn_grid = 384;      // CTAs of the decoder attention stage kernels (0 = one per (row, head) unit); 384 measured best at B = 64

  // ---- weights
  Linear patch;
  float* pos = nullptr;   // [197,768]
  float* cls = nullptr;   // [768]
  EncLayer enc[kEncLayers];
  LnParams enc_ln;
  Linear cross_kv;        // [3072,768]: K0 V0 K1 V1
  EmbedWeights emb{};
  DecLayer dec[kDecLayers];
  Linear head_t, head_dec;
  LnParams head_ln;
  float* lut = nullptr;   // [256] reference rescale+normalize values

  // ---- workspace
  int rows_cap = 0;       // encoder token rows (multiple of 128)
  int brow_cap = 0;       // decoder rows (multiple of 128)
  ActBuf patches, xn, ctx, mlp, enc_out;
  __nv_bfloat16* qkv = nullptr;       // [rows_cap, 2304]
  CUtensorMap map_qkv_q, map_qkv_kv;  // TMA views of qkv for the tcgen05 attention (boxes 64 x 128 and 64 x 208)
  float* hres = nullptr;              // [rows_cap, 768] fp32 residual stream
  float* enc_f32 = nullptr;           // tap
  __nv_bfloat16* crosskv = nullptr;   // [crop][layer][K|V][head][197][64]: cross-attention K/V, one contiguous block per (crop, layer, head)
  ActBuf d_xb, d_ctx, d_ffn, d_tb;
  float* d_x = nullptr;               // [brow_cap, 768]
  __nv_bfloat16* d_qkv = nullptr;     // [brow_cap, 2304]
  __nv_bfloat16* d_q = nullptr;       // [brow_cap, 768]
  __nv_bfloat16* self_k[kDecLayers] = {nullptr, nullptr};   // [max_batch, max_length, 768]
  __nv_bfloat16* self_v[kDecLayers] = {nullptr, nullptr};
  __nv_bfloat16* self_k2[kDecLayers] = {nullptr, nullptr};  // beam mode: second cache set (the gather that follows the beams ping-pongs)
  __nv_bfloat16* self_v2[kDecLayers] = {nullptr, nullptr};
  void* d_beam = nullptr;             // beam mode scratch: candidates, next tokens, parents, ban lists
  size_t d_beam_bytes = 0;
  void* d_beam_dev = nullptr;         // device-resident beam search: token rows, scores, control block (beam_device.cuh)
  size_t d_beam_dev_bytes = 0;
  int* h_beam_ctl = nullptr;          // pinned [8]
  float* part_max = nullptr;
  int* part_idx = nullptr;
  int* d_ids = nullptr;               // [max_batch, max_length]
  int* d_pos = nullptr;
  int* d_finished = nullptr;
  int* d_forced = nullptr;
  int* d_zero = nullptr;              // [max_batch] zeros
  int* d_lens = nullptr;              // [max_batch] valid ids per crop of the last decode
  int* d_slot_crop = nullptr;         // [max_batch] crop each decoder row is working on
  int* d_queue = nullptr;             // [4] next waiting crop | crops ready | crops finished
  int* h_queue = nullptr;             // pinned [4]
  int* h_flags = nullptr;             // pinned [max_batch]
  float* d_y = nullptr;               // [3, brow_cap, 768] split-K partials of the projections feeding a LayerNorm
  float* d_yq = nullptr;              // [3, brow_cap, 768] split-K partials of the cross-attention query
  long long* d_prof = nullptr;        // [4096] stage timeline of the decoder (option decode_prof)
  int decode_prof = 0;
  float* logits_tap = nullptr;        // [n, max_length-1, 6144]
  size_t logits_tap_bytes = 0;
  uint8_t* px_u8 = nullptr;           // taps [max_batch,224,224]
  float* px_f32 = nullptr;

  // ---- crop staging
  uint8_t* h_arena = nullptr;
  uint8_t* d_arena = nullptr;
  size_t arena_cap = 0;
  // region staging (polygon masks rasterised on the device)
  uint8_t* d_masks = nullptr;
  size_t masks_cap = 0;
  void* d_mask_meta = nullptr;        // jobs | edges | lines, one upload
  size_t mask_meta_cap = 0;
  std::vector<long long> region_mask_off;   // per staged region: mask offset or -1 (test hook)
  std::vector<int> region_mask_hw;          // per staged region: source crop h, w
  CropDesc* h_descs = nullptr;        // pinned [max_batch]
  CropDesc* d_descs = nullptr;
  std::vector<int> h_coefs;
  int* d_coefs = nullptr;
  size_t d_coefs_cap = 0, d_coefs_used = 0;
  struct TableRef { int offset, ksize, strip_rows; };
  std::unordered_map<int, TableRef> tables;
  int pre_pitch = 0, pre_tmp_rows = 0, pre_bgr = 0;

  // ---- batch state
  int n = 0;              // crops of the current batch
  int cur_len = 0;        // max_length of the last decode
  int last_steps = 0;
  int last_rows = 0;      // decoder rows of the last decode
  bool staged_ok = false, pre_ok = false, enc_ok = false, dec_ok = false;

  // ---- decode-step graphs keyed by (n, max_length, forced?, tap?)
  struct StepGraph { cudaGraphExec_t exec; int launches; };
  std::map<uint64_t, StepGraph> graphs;
  std::map<uint64_t, StepGraph> enc_graphs;   // encoder launches per (batch size, tap)
  std::map<uint64_t, StepGraph> beam_graphs;  // device-resident beam search: steps per (crops, beams, length, n-gram, early stopping, penalty)
};

namespace {

int fail(mocr_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h != nullptr) h->error = buf; else g_create_error = buf;
  return code;
}

#define CK(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      return fail(h, MOCR_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define TRY(call)             \
  do {                        \
    int r_ = (call);          \
    if (r_ != MOCR_OK) return r_; \
  } while (0)

template <typename T>
int dmalloc(mocr_handle* h, T** out, size_t count, bool zero = true) {
  void* p = nullptr;
  CK(cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 256)));
  h->allocs.push_back(p);
  if (zero) CK(cudaMemsetAsync(p, 0, std::max<size_t>(count * sizeof(T), 256), h->stream));
  *out = static_cast<T*>(p);
  return MOCR_OK;
}

int make_map(mocr_handle* h, CUtensorMap* m, const void* base, int rows, int cols, int box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return fail(h, MOCR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kGemmBK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, MOCR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d box=%d", (int)r, rows, cols, box_rows);
  return MOCR_OK;
}

// f32 [rows, cols] view for the TMA reduce-add epilogue: 32 x 32 boxes (128-byte rows, SWIZZLE_128B)
int make_map_f32_out(mocr_handle* h, CUtensorMap* m, const void* base, int rows, int cols) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return fail(h, MOCR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, MOCR_ERR_CUDA, "cuTensorMapEncodeTiled (f32 out) failed (%d) rows=%d cols=%d", (int)r, rows, cols);
  return MOCR_OK;
}

// bf16 [rows, cols] view of a GEMM output for the TMA-store epilogue: box 64 columns x 32 rows (one epilogue warp's staging tile)
int make_map_bf16_out(mocr_handle* h, CUtensorMap* m, const void* base, int rows, int cols) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return fail(h, MOCR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {64, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, MOCR_ERR_CUDA, "cuTensorMapEncodeTiled (bf16 out) failed (%d) rows=%d cols=%d", (int)r, rows, cols);
  return MOCR_OK;
}

int make_act(mocr_handle* h, ActBuf* a, int rows_cap, int K) {
  a->rows_cap = rows_cap;
  a->K = K;
  TRY(dmalloc(h, &a->p, static_cast<size_t>(rows_cap) * K));
  return make_map(h, &a->map, a->p, rows_cap, K, kGemmBM);
}

// ------------------------------------------------------------------ weights ---

const HostTensor* find(mocr_handle* h, const std::string& name) {
  auto it = h->staged.find(name);
  return it == h->staged.end() ? nullptr : &it->second;
}

int need(mocr_handle* h, const std::string& name, size_t numel, const HostTensor** out) {
  const HostTensor* t = find(h, name);
  if (t == nullptr) return fail(h, MOCR_ERR_WEIGHTS, "missing tensor %s", name.c_str());
  if (t->data.size() != numel) return fail(h, MOCR_ERR_WEIGHTS, "tensor %s has %zu elements, expected %zu", name.c_str(), t->data.size(), numel);
  *out = t;
  return MOCR_OK;
}

int upload_f32(mocr_handle* h, float** dst, const float* src, size_t n) {
  TRY(dmalloc(h, dst, n, false));
  CK(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
  return MOCR_OK;
}
int upload_vec(mocr_handle* h, float** dst, const std::string& name, size_t n) {
  const HostTensor* t;
  TRY(need(h, name, n, &t));
  return upload_f32(h, dst, t->data.data(), n);
}
int upload_ln(mocr_handle* h, LnParams* ln, const std::string& prefix) {
  TRY(upload_vec(h, &ln->g, prefix + ".weight", kD));
  return upload_vec(h, &ln->b, prefix + ".bias", kD);
}

// rows of `parts` are concatenated along N; row_scale[i] multiplies part i (weights and bias).
int upload_linear(mocr_handle* h, Linear* L, int K, const std::vector<std::string>& parts, const std::vector<int>& part_n,
                  const std::vector<float>& part_scale) {
  int N = 0;
  for (int v : part_n) N += v;
  std::vector<uint16_t> wb(static_cast<size_t>(N) * K);
  std::vector<float> bias(static_cast<size_t>(N));
  size_t row = 0;
  for (size_t i = 0; i < parts.size(); ++i) {
    const HostTensor *w, *b;
    TRY(need(h, parts[i] + ".weight", static_cast<size_t>(part_n[i]) * K, &w));
    TRY(need(h, parts[i] + ".bias", static_cast<size_t>(part_n[i]), &b));
    const float sc = part_scale[i];
    for (size_t j = 0; j < static_cast<size_t>(part_n[i]) * K; ++j) wb[row * K + j] = f32_to_bf16(w->data[j] * sc);
    for (int j = 0; j < part_n[i]; ++j) bias[row + j] = b->data[j] * sc;
    row += part_n[i];
  }
  L->N = N;
  L->K = K;
  TRY(dmalloc(h, &L->w, wb.size(), false));
  CK(cudaMemcpy(L->w, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
  return upload_f32(h, &L->bias, bias.data(), bias.size());
}
int upload_linear1(mocr_handle* h, Linear* L, int N, int K, const std::string& name, float scale = 1.0f) {
  return upload_linear(h, L, K, {name}, {N}, {scale});
}

int linear_map(mocr_handle* h, Linear* L, int bn, const CUtensorMap** out) {
  auto it = L->maps.find(bn);
  if (it == L->maps.end()) {
    CUtensorMap m;
    TRY(make_map(h, &m, L->w, L->N, L->K, bn));
    it = L->maps.emplace(bn, m).first;
  }
  *out = &it->second;
  return MOCR_OK;
}

int finalize_weights(mocr_handle* h) {
  const std::string e = "encoder.", el = "encoder.encoder.layer.";
  // ---- ViT patch embedding: fold the 3 equal input planes (K 768 -> 256) and the
  // rescale/normalize affine map x = (u8-128)*2/255 + 1/255 into weight and bias.
  {
    const HostTensor *w, *b;
    TRY(need(h, e + "embeddings.patch_embeddings.projection.weight", static_cast<size_t>(kD) * 3 * 256, &w));
    TRY(need(h, e + "embeddings.patch_embeddings.projection.bias", kD, &b));
    std::vector<uint16_t> wb(static_cast<size_t>(kD) * kPatchK);
    std::vector<float> bias(kD);
    for (int n = 0; n < kD; ++n) {
      double tot = 0.0;
      for (int k = 0; k < kPatchK; ++k) {
        const double s = static_cast<double>(w->data[(static_cast<size_t>(n) * 3 + 0) * 256 + k]) +
                         w->data[(static_cast<size_t>(n) * 3 + 1) * 256 + k] + w->data[(static_cast<size_t>(n) * 3 + 2) * 256 + k];
        tot += s;
        wb[static_cast<size_t>(n) * kPatchK + k] = f32_to_bf16(static_cast<float>(s * (2.0 / 255.0)));
      }
      bias[n] = static_cast<float>(b->data[n] + tot * (1.0 / 255.0));
    }
    h->patch.N = kD;
    h->patch.K = kPatchK;
    TRY(dmalloc(h, &h->patch.w, wb.size(), false));
    CK(cudaMemcpy(h->patch.w, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
    TRY(upload_f32(h, &h->patch.bias, bias.data(), bias.size()));
  }
  TRY(upload_vec(h, &h->pos, e + "embeddings.position_embeddings", static_cast<size_t>(kEncTokens) * kD));
  TRY(upload_vec(h, &h->cls, e + "embeddings.cls_token", kD));
  for (int i = 0; i < kEncLayers; ++i) {
    const std::string p = el + std::to_string(i) + ".";
    EncLayer& L = h->enc[i];
    TRY(upload_ln(h, &L.ln1, p + "layernorm_before"));
    TRY(upload_ln(h, &L.ln2, p + "layernorm_after"));
    TRY(upload_linear(h, &L.qkv, kD, {p + "attention.attention.query", p + "attention.attention.key", p + "attention.attention.value"},
                      {kD, kD, kD}, {kQScale, 1.f, 1.f}));
    TRY(upload_linear1(h, &L.out, kD, kD, p + "attention.output.dense"));
    TRY(upload_linear1(h, &L.fc1, kFFN, kD, p + "intermediate.dense"));
    TRY(upload_linear1(h, &L.fc2, kD, kFFN, p + "output.dense"));
  }
  TRY(upload_ln(h, &h->enc_ln, e + "layernorm"));

  const std::string d = "decoder.bert.", dl = "decoder.bert.encoder.layer.";
  {
    float* t;
    TRY(upload_vec(h, &t, d + "embeddings.word_embeddings.weight", static_cast<size_t>(kVocab) * kD));
    h->emb.word = t;
    TRY(upload_vec(h, &t, d + "embeddings.position_embeddings.weight", static_cast<size_t>(kMaxPos) * kD));
    h->emb.posemb = t;
    const HostTensor* tt;
    TRY(need(h, d + "embeddings.token_type_embeddings.weight", 2 * kD, &tt));
    TRY(upload_f32(h, &t, tt->data.data(), kD));   // row 0 only: token_type_ids are all zero
    h->emb.type0 = t;
    LnParams ln;
    TRY(upload_ln(h, &ln, d + "embeddings.LayerNorm"));
    h->emb.gamma = ln.g;
    h->emb.beta = ln.b;
  }
  std::vector<std::string> ckv;
  for (int i = 0; i < kDecLayers; ++i) {
    const std::string p = dl + std::to_string(i) + ".";
    DecLayer& L = h->dec[i];
    TRY(upload_linear(h, &L.self_qkv, kD, {p + "attention.self.query", p + "attention.self.key", p + "attention.self.value"},
                      {kD, kD, kD}, {kQScale, 1.f, 1.f}));
    TRY(upload_linear1(h, &L.self_out, kD, kD, p + "attention.output.dense"));
    TRY(upload_ln(h, &L.ln_self, p + "attention.output.LayerNorm"));
    TRY(upload_linear1(h, &L.cross_q, kD, kD, p + "crossattention.self.query", kQScale));
    TRY(upload_linear1(h, &L.cross_out, kD, kD, p + "crossattention.output.dense"));
    TRY(upload_ln(h, &L.ln_cross, p + "crossattention.output.LayerNorm"));
    TRY(upload_linear1(h, &L.fc1, kFFN, kD, p + "intermediate.dense"));
    TRY(upload_linear1(h, &L.fc2, kD, kFFN, p + "output.dense"));
    TRY(upload_ln(h, &L.ln_ffn, p + "output.LayerNorm"));
    ckv.push_back(p + "crossattention.self.key");
    ckv.push_back(p + "crossattention.self.value");
  }
  TRY(upload_linear(h, &h->cross_kv, kD, ckv, {kD, kD, kD, kD}, {1.f, 1.f, 1.f, 1.f}));
  const std::string c = "decoder.cls.predictions.";
  TRY(upload_linear1(h, &h->head_t, kD, kD, c + "transform.dense"));
  TRY(upload_ln(h, &h->head_ln, c + "transform.LayerNorm"));
  {
    // LM head: weight may be tied to the word embeddings; bias is predictions.bias.
    const HostTensor* w = find(h, c + "decoder.weight");
    if (w == nullptr) w = find(h, d + "embeddings.word_embeddings.weight");
    const HostTensor* b = find(h, c + "bias");
    if (b == nullptr) b = find(h, c + "decoder.bias");
    if (w == nullptr || b == nullptr || w->data.size() != static_cast<size_t>(kVocab) * kD || b->data.size() != static_cast<size_t>(kVocab))
      return fail(h, MOCR_ERR_WEIGHTS, "missing or mis-shaped LM head (decoder.cls.predictions.decoder.weight / .bias)");
    std::vector<uint16_t> wb(w->data.size());
    for (size_t j = 0; j < wb.size(); ++j) wb[j] = f32_to_bf16(w->data[j]);
    h->head_dec.N = kVocab;
    h->head_dec.K = kD;
    TRY(dmalloc(h, &h->head_dec.w, wb.size(), false));
    CK(cudaMemcpy(h->head_dec.w, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
    TRY(upload_f32(h, &h->head_dec.bias, b->data.data(), kVocab));
  }
  {
    // rescale + normalize as the reference computes them (transformers/image_transforms.py:
    // 89-124 then 384-442): float32(float64(v) * (1/255)), then (x - 0.5f) / 0.5f in float32.
    float lut[256];
    for (int v = 0; v < 256; ++v) {
      const float x = static_cast<float>(static_cast<double>(v) * (1.0 / 255.0));
      lut[v] = (x - 0.5f) / 0.5f;
    }
    TRY(upload_f32(h, &h->lut, lut, 256));
  }
  h->staged.clear();
  h->finalized = true;
  return MOCR_OK;
}

// ------------------------------------------------------------------ launches ---

template <int BN, int EPI>
int launch_gemm_t(mocr_handle* h, const CUtensorMap& ma, const CUtensorMap& mb, const GemmArgs& a) {
  using Cfg = GemmCfg<BN>;
  static bool attr_done[16] = {};
  if (!attr_done[h->device & 15]) {
    CK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes(Cfg::kSmemBytes, EPI)));
    attr_done[h->device & 15] = true;
  }
  const int tiles = ((a.M + kGemmBM - 1) / kGemmBM) * (a.N / BN);
  const int grid = std::min(tiles, h->sms);
  gemm_tcgen05_kernel<BN, EPI><<<grid, kGemmThreads, gemm_smem_bytes(Cfg::kSmemBytes, EPI), h->stream>>>(ma, mb, a);
  CK(cudaGetLastError());
  ++h->launches;
  return MOCR_OK;
}

template <int EPI>
int launch_gemm_bn(mocr_handle* h, int bn, const CUtensorMap& ma, const CUtensorMap& mb, const GemmArgs& a) {
  switch (bn) {
    case 32: return launch_gemm_t<32, EPI>(h, ma, mb, a);
    case 64: return launch_gemm_t<64, EPI>(h, ma, mb, a);
    case 128: return launch_gemm_t<128, EPI>(h, ma, mb, a);
    case 192: return launch_gemm_t<192, EPI>(h, ma, mb, a);
    case 256: return launch_gemm_t<256, EPI>(h, ma, mb, a);
    default: return fail(h, MOCR_ERR_INVALID, "unsupported GEMM tile width %d", bn);
  }
}

// out = epilogue(A[M,K] * W^T + bias)
int gemm(mocr_handle* h, int epi, int bn, const ActBuf& A, Linear& L, int M, GemmArgs a) {
  if (A.K != L.K || M > A.rows_cap || L.N % bn != 0 || L.K % kGemmBK != 0)
    return fail(h, MOCR_ERR_INVALID, "gemm shape mismatch: M=%d A.K=%d W=[%d,%d] bn=%d", M, A.K, L.N, L.K, bn);
  a.M = M;
  a.N = L.N;
  a.K = L.K;
  a.bias = L.bias;
#ifdef MOCR_GEMM_DBG
  a.dbg = h->gemm_dbg;
#endif
  const CUtensorMap* mb;
  TRY(linear_map(h, &L, bn, &mb));
  if (epi == EPI_F32_ACCUM) {
    if (a.ldo != L.N || bn > 256 || bn % 64 != 0) return fail(h, MOCR_ERR_INVALID, "accumulating epilogue needs a dense [M, N] output and a 64-column multiple tile");
    TRY(make_map_f32_out(h, &a.tmap_out, a.out, M, L.N));
  }
  if ((epi == EPI_BF16 || epi == EPI_BF16_GELU) && h->enc_tma_store && bn % 128 == 0 && a.ldo == L.N) {
    TRY(make_map_bf16_out(h, &a.tmap_out, a.out, M, L.N));
    a.out_tma = 1;
  }
  switch (epi) {
    case EPI_BF16: return launch_gemm_bn<EPI_BF16>(h, bn, A.map, *mb, a);
    case EPI_BF16_GELU: return launch_gemm_bn<EPI_BF16_GELU>(h, bn, A.map, *mb, a);
    case EPI_F32_RESID: return launch_gemm_bn<EPI_F32_RESID>(h, bn, A.map, *mb, a);
    case EPI_PATCH: return launch_gemm_bn<EPI_PATCH>(h, bn, A.map, *mb, a);
    case EPI_ARGMAX: return launch_gemm_bn<EPI_ARGMAX>(h, bn, A.map, *mb, a);
    case EPI_F32_GELU: return launch_gemm_bn<EPI_F32_GELU>(h, bn, A.map, *mb, a);
    case EPI_CROSSKV: return launch_gemm_bn<EPI_CROSSKV>(h, bn, A.map, *mb, a);
    case EPI_F32_ACCUM: return launch_gemm_bn<EPI_F32_ACCUM>(h, bn, A.map, *mb, a);
    default: return fail(h, MOCR_ERR_INVALID, "bad epilogue %d", epi);
  }
}

GemmArgs out_bf16(__nv_bfloat16* out, int ldo) {
  GemmArgs a{};
  a.out = out;
  a.ldo = ldo;
  return a;
}
GemmArgs out_f32(float* out, int ldo, const float* resid = nullptr, int ldr = 0) {
  GemmArgs a{};
  a.out = out;
  a.ldo = ldo;
  a.resid = resid;
  a.ldr = ldr;
  return a;
}

int layernorm(mocr_handle* h, const float* x, int rows, const LnParams& ln, __nv_bfloat16* ob, float* of) {
  layernorm_rows_kernel<<<(rows + 7) / 8, 256, 0, h->stream>>>(x, rows, ln.g, ln.b, ob, of);
  CK(cudaGetLastError());
  ++h->launches;
  return MOCR_OK;
}

// ------------------------------------------------------------------ preprocess ---

int table_for(mocr_handle* h, int in_size, mocr_handle::TableRef* out) {
  auto it = h->tables.find(in_size);
  if (it == h->tables.end()) {
    ResampleTable t = make_resample_table(in_size);
    mocr_handle::TableRef r{static_cast<int>(h->h_coefs.size()), t.ksize, t.max_strip_rows};
    h->h_coefs.insert(h->h_coefs.end(), t.data.begin(), t.data.end());
    it = h->tables.emplace(in_size, r).first;
  }
  *out = it->second;
  return MOCR_OK;
}

// chunk > 0 (large batches): the pixels travel in chunks of `chunk` crops on the copy stream, and after_chunk(i0, cnt) is called as
// soon as a chunk's copies are enqueued (the handle's stream already waits for them): the caller launches that chunk's preprocess
// and encoder there, so the host-side copy of chunk i+1 overlaps the GPU work on chunk i.
// desc_base (session mode): the crops take descriptor / crop indices [desc_base, desc_base + n) instead of [0, n).
int stage_crops(mocr_handle* h, const mocr_crop_t* crops, int n, int order, int chunk = 0, const std::function<int(int, int)>* after_chunk = nullptr,
                int desc_base = -1) {
  if (h->sess_on != (desc_base >= 0)) return fail(h, MOCR_ERR_INVALID, h->sess_on ? "a session is active on this handle (mocr_session_end first)" : "no session is active");
  if (desc_base < 0) desc_base = 0;
  if (!h->finalized) return fail(h, MOCR_ERR_INVALID, "weights are not finalized");
  if (n < 1 || desc_base + n > h->max_batch) return fail(h, MOCR_ERR_CAPACITY, "batch of %d crops, handle capacity is %d", n, h->max_batch);
  if (crops == nullptr) return fail(h, MOCR_ERR_INVALID, "crops is NULL");
  h->staged_ok = h->pre_ok = h->enc_ok = h->dec_ok = false;
  h->region_mask_off.clear();
  size_t total = 0;
  int max_w = 0, tmp_rows = kPreStripRows;
  for (int i = 0; i < n; ++i) {
    const mocr_crop_t& c = crops[i];
    if (c.data == nullptr || c.height < 1 || c.width < 1 || (c.channels != 1 && c.channels != 3 && c.channels != 4) ||
        c.stride < c.width * c.channels)
      return fail(h, MOCR_ERR_INVALID, "crop %d is malformed (h=%d w=%d stride=%d channels=%d)", i, c.height, c.width, c.stride, c.channels);
    if (c.height > 32768 || c.width > 32768) return fail(h, MOCR_ERR_CAPACITY, "crop %d is larger than 32768 px", i);
    total += (static_cast<size_t>(c.height) * c.width * c.channels + 15) & ~static_cast<size_t>(15);
    max_w = std::max(max_w, c.width);
  }
  // the previous batch may still be reading the arena / tables
  {
    const auto t0 = std::chrono::steady_clock::now();
    if (h->sess_on) CK(cudaEventSynchronize(h->ev_staged));    // (an admission: the previous one's encoder pass may still run - only its preprocess reads them)
    else CK(cudaStreamSynchronize(h->stream));
    h->prof_add[0] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  if (total > h->arena_cap) {
    if (h->h_arena) cudaFreeHost(h->h_arena);
    if (h->d_arena) cudaFree(h->d_arena);
    h->h_arena = nullptr;
    h->d_arena = nullptr;
    h->arena_cap = 0;
    const size_t cap = std::max<size_t>(total + total / 4, 8u << 20);
    CK(cudaMallocHost(reinterpret_cast<void**>(&h->h_arena), cap));
    CK(cudaMalloc(reinterpret_cast<void**>(&h->d_arena), cap));
    h->arena_cap = cap;
  }
  // descriptors + arena offsets
  std::vector<size_t> offs(static_cast<size_t>(n) + 1, 0);
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    const mocr_crop_t& c = crops[i];
    const size_t rowb = static_cast<size_t>(c.width) * c.channels;
    CropDesc& d = h->h_descs[desc_base + i];
    d.offset = static_cast<long long>(off);
    d.h = c.height;
    d.w = c.width;
    d.stride = static_cast<int>(rowb);
    d.channels = c.channels;
    d.hcoef = d.vcoef = -1;
    d.hks = d.vks = 0;
    d.region = 0;
    d.rot = kRotNone;
    d.ox = d.oy = 0;
    d.ph = c.height;
    d.pw = c.width;
    d.sw = c.width;
    d.mask = -1;
    mocr_handle::TableRef t;
    if (c.width != kImage) {
      TRY(table_for(h, c.width, &t));
      d.hcoef = t.offset;
      d.hks = t.ksize;
    }
    if (c.height != kImage) {
      TRY(table_for(h, c.height, &t));
      d.vcoef = t.offset;
      d.vks = t.ksize;
      tmp_rows = std::max(tmp_rows, t.strip_rows);
    }
    offs[i] = off;
    off += (rowb * c.height + 15) & ~static_cast<size_t>(15);
  }
  offs[n] = off;
  // resampling tables and descriptors first: a chunk's preprocess may start as soon as its pixels have arrived
  h->pre_pitch = round_up(max_w, 16);
  h->pre_tmp_rows = tmp_rows;
  h->pre_bgr = order == MOCR_BGR ? 1 : 0;
  const size_t smem = pre_smem_bytes(h->pre_pitch, tmp_rows);
  if (smem > 200 * 1024) return fail(h, MOCR_ERR_CAPACITY, "crop extents need %zu B of shared memory (limit 204800)", smem);
  if (h->h_coefs.size() > h->d_coefs_cap) {
    if (h->d_coefs) cudaFree(h->d_coefs);
    h->d_coefs = nullptr;
    h->d_coefs_cap = h->d_coefs_used = 0;
    const size_t cap = std::max<size_t>(h->h_coefs.size() * 2, 1u << 20);
    CK(cudaMalloc(reinterpret_cast<void**>(&h->d_coefs), cap * sizeof(int)));
    h->d_coefs_cap = cap;
  }
  if (h->h_coefs.size() > h->d_coefs_used) {
    CK(cudaMemcpyAsync(h->d_coefs + h->d_coefs_used, h->h_coefs.data() + h->d_coefs_used,
                       (h->h_coefs.size() - h->d_coefs_used) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    // h_coefs is pageable: the copy above is staged synchronously by the runtime, safe to reuse
    h->d_coefs_used = h->h_coefs.size();
  }
  CK(cudaMemcpyAsync(h->d_descs + desc_base, h->h_descs + desc_base, sizeof(CropDesc) * n, cudaMemcpyHostToDevice, h->stream));
  h->n = n;
  h->staged_ok = true;
  // pixels -> pinned arena -> device.  Workers take contiguous runs of crops of about equal bytes; each uploads its run
  // as soon as it is copied (slices of >= 1 MB), so host copy and H2D overlap and a page batch (100+ MB) is not bound by
  // one core's memcpy rate.
  const bool chunked = chunk > 0 && after_chunk != nullptr && n > chunk && h->stream_enc != nullptr;
  cudaStream_t copy_stream = chunked ? h->stream_enc : h->stream;
  auto copy_run = [&](int lo, int hi) -> cudaError_t {
    size_t sent = offs[lo];
    for (int i = lo; i < hi; ++i) {
      const mocr_crop_t& c = crops[i];
      const size_t rowb = static_cast<size_t>(c.width) * c.channels;
      if (static_cast<size_t>(c.stride) == rowb) {
        memcpy(h->h_arena + offs[i], c.data, rowb * c.height);
      } else {
        for (int y = 0; y < c.height; ++y) memcpy(h->h_arena + offs[i] + y * rowb, c.data + static_cast<size_t>(y) * c.stride, rowb);
      }
      if (offs[i + 1] - sent >= (1u << 20) || i + 1 == hi) {
        const cudaError_t e = cudaMemcpyAsync(h->d_arena + sent, h->h_arena + sent, offs[i + 1] - sent, cudaMemcpyHostToDevice, copy_stream);
        if (e != cudaSuccess) return e;
        sent = offs[i + 1];
      }
    }
    return cudaSuccess;
  };
  if (chunked) CK(cudaStreamSynchronize(h->stream_enc));
  for (int c0 = 0; c0 < n; c0 += chunked ? chunk : n) {
  const int c1 = chunked ? std::min(n, c0 + chunk) : n;
  const size_t cbytes = offs[c1] - offs[c0];
  const int workers = static_cast<int>(std::min<size_t>(h->stage_threads, std::max<size_t>(1, cbytes >> 23)));     // one per 8 MB, at most stage_threads
  if (workers <= 1) {
    CK(copy_run(c0, c1));
  } else {
    std::vector<std::thread> pool;
    std::vector<cudaError_t> errs(workers, cudaSuccess);
    int lo = c0;
    for (int w = 0; w < workers; ++w) {
      const size_t target = offs[c0] + cbytes * (w + 1) / workers;
      int hi = lo;
      while (hi < c1 && (offs[hi + 1] <= target || w == workers - 1)) ++hi;
      if (w == workers - 1) hi = c1;
      const int device = h->device;
      try {
        pool.emplace_back([&, w, lo, hi, device]() {
          cudaSetDevice(device);
          errs[w] = copy_run(lo, hi);
        });
      } catch (...) {          // no thread to be had: this run is copied here (nothing may cross the C ABI as an exception)
        errs[w] = copy_run(lo, hi);
      }
      lo = hi;
    }
    for (auto& th : pool) th.join();
    for (cudaError_t e : errs)
      if (e != cudaSuccess) return fail(h, MOCR_ERR_CUDA, "crop upload failed: %s", cudaGetErrorString(e));
  }
  if (chunked) {
    CK(cudaEventRecord(h->ev_first, copy_stream));
    CK(cudaStreamWaitEvent(h->stream, h->ev_first, 0));
    TRY((*after_chunk)(c0, c1 - c0));
  }
  }
  return MOCR_OK;
}

// Region staging: ONE upload of the page, then every selection is a view of it -
// crop box, optional polygon (rasterised on the device into a mask), optional 90-degree rotation -
// resolved inside the preprocess kernel's reads.  Replaces, per selection, the reference's
// PIL crop + RGB->BGR + fillPoly/bitwise_and/add composite + cv2.rotate + BGR->RGB host copies
// (reference/src/ui/main_window.py:6497-6506, 9789-9800).
int stage_regions(mocr_handle* h, const mocr_crop_t* page, const mocr_region_t* regions, int n, int order) {
  if (!h->finalized) return fail(h, MOCR_ERR_INVALID, "weights are not finalized");
  if (n < 1 || n > h->max_batch) return fail(h, MOCR_ERR_CAPACITY, "batch of %d regions, handle capacity is %d", n, h->max_batch);
  if (page == nullptr || regions == nullptr) return fail(h, MOCR_ERR_INVALID, "page or regions is NULL");
  if (h->sess_on) return fail(h, MOCR_ERR_INVALID, "a session is active on this handle (mocr_session_end first)");
  const mocr_crop_t& pg = *page;
  if (pg.data == nullptr || pg.height < 1 || pg.width < 1 || (pg.channels != 1 && pg.channels != 3 && pg.channels != 4) ||
      pg.stride < pg.width * pg.channels)
    return fail(h, MOCR_ERR_INVALID, "page is malformed (h=%d w=%d stride=%d channels=%d)", pg.height, pg.width, pg.stride, pg.channels);
  h->staged_ok = h->pre_ok = h->enc_ok = h->dec_ok = false;
  CK(cudaStreamSynchronize(h->stream));     // the previous batch may still be reading the pinned descriptors / arena / tables / masks
  std::vector<MaskJob> jobs;
  std::vector<MaskEdge> edges;
  std::vector<MaskLine> lines;
  std::vector<int32_t> rel;
  size_t mask_bytes = 0;
  int max_w = 0, tmp_rows = kPreStripRows, max_mask_h = 0;
  h->region_mask_off.assign(n, -1);
  h->region_mask_hw.assign(2 * n, 0);
  for (int i = 0; i < n; ++i) {
    const mocr_region_t& r = regions[i];
    const long long sw = static_cast<long long>(r.right) - r.left, sh = static_cast<long long>(r.bottom) - r.top;
    if (sw < 1 || sh < 1 || sw > 32768 || sh > 32768) return fail(h, MOCR_ERR_INVALID, "region %d has an empty or oversized box (%lld x %lld)", i, sw, sh);
    if (r.rotate < 0 || r.rotate > 2) return fail(h, MOCR_ERR_INVALID, "region %d: rotate must be 0, 1 (clockwise) or 2 (counter-clockwise)", i);
    if (r.n_points < 0 || r.n_points > kMaskMaxEdges || (r.n_points > 0 && r.polygon == nullptr))
      return fail(h, r.n_points > kMaskMaxEdges ? MOCR_ERR_CAPACITY : MOCR_ERR_INVALID, "region %d: bad polygon (%d points, limit %d)", i, r.n_points, kMaskMaxEdges);
    CropDesc& d = h->h_descs[i];
    d.offset = 0;
    d.h = static_cast<int>(r.rotate == kRotNone ? sh : sw);
    d.w = static_cast<int>(r.rotate == kRotNone ? sw : sh);
    d.stride = pg.width * pg.channels;      // the page is packed into the arena
    d.channels = pg.channels;
    d.hcoef = d.vcoef = -1;
    d.hks = d.vks = 0;
    d.region = 1;
    d.rot = r.rotate;
    d.ox = r.left;
    d.oy = r.top;
    d.ph = pg.height;
    d.pw = pg.width;
    d.sw = static_cast<int>(sw);
    d.mask = -1;
    h->region_mask_hw[2 * i] = static_cast<int>(sh);
    h->region_mask_hw[2 * i + 1] = static_cast<int>(sw);
    if (r.n_points > 0) {
      rel.resize(2 * static_cast<size_t>(r.n_points));
      for (int k = 0; k < r.n_points; ++k) {
        rel[2 * k] = r.polygon[2 * k] - r.left;
        rel[2 * k + 1] = r.polygon[2 * k + 1] - r.top;
      }
      MaskJob j;
      j.mask_off = static_cast<long long>(mask_bytes);
      j.h = static_cast<int>(sh);
      j.w = static_cast<int>(sw);
      j.edge0 = static_cast<int>(edges.size());
      j.line0 = static_cast<int>(lines.size());
      collect_poly_edges(j.h, j.w, rel.data(), r.n_points, edges, lines);
      j.n_edges = static_cast<int>(edges.size()) - j.edge0;
      j.n_lines = static_cast<int>(lines.size()) - j.line0;
      jobs.push_back(j);
      d.mask = j.mask_off;
      h->region_mask_off[i] = j.mask_off;
      mask_bytes += (static_cast<size_t>(sh) * sw + 15) & ~static_cast<size_t>(15);
      max_mask_h = std::max(max_mask_h, j.h);
    }
    mocr_handle::TableRef t;
    if (d.w != kImage) {
      TRY(table_for(h, d.w, &t));
      d.hcoef = t.offset;
      d.hks = t.ksize;
    }
    if (d.h != kImage) {
      TRY(table_for(h, d.h, &t));
      d.vcoef = t.offset;
      d.vks = t.ksize;
      tmp_rows = std::max(tmp_rows, t.strip_rows);
    }
    max_w = std::max(max_w, d.w);
  }
  h->pre_pitch = round_up(max_w, 16);
  h->pre_tmp_rows = tmp_rows;
  h->pre_bgr = order == MOCR_BGR ? 1 : 0;
  const size_t smem = pre_smem_bytes(h->pre_pitch, tmp_rows);
  if (smem > 200 * 1024) return fail(h, MOCR_ERR_CAPACITY, "region extents need %zu B of shared memory (limit 204800)", smem);

  const size_t rowb = static_cast<size_t>(pg.width) * pg.channels, total = rowb * pg.height;
  if (total > h->arena_cap) {
    if (h->h_arena) cudaFreeHost(h->h_arena);
    if (h->d_arena) cudaFree(h->d_arena);
    h->h_arena = nullptr;
    h->d_arena = nullptr;
    h->arena_cap = 0;
    const size_t cap = std::max<size_t>(total + total / 4, 8u << 20);
    CK(cudaMallocHost(reinterpret_cast<void**>(&h->h_arena), cap));
    CK(cudaMalloc(reinterpret_cast<void**>(&h->d_arena), cap));
    h->arena_cap = cap;
  }
  if (static_cast<size_t>(pg.stride) == rowb) memcpy(h->h_arena, pg.data, total);
  else for (int y = 0; y < pg.height; ++y) memcpy(h->h_arena + y * rowb, pg.data + static_cast<size_t>(y) * pg.stride, rowb);
  CK(cudaMemcpyAsync(h->d_arena, h->h_arena, total, cudaMemcpyHostToDevice, h->stream));
  if (h->h_coefs.size() > h->d_coefs_cap) {
    if (h->d_coefs) cudaFree(h->d_coefs);
    h->d_coefs = nullptr;
    h->d_coefs_cap = h->d_coefs_used = 0;
    const size_t cap = std::max<size_t>(h->h_coefs.size() * 2, 1u << 20);
    CK(cudaMalloc(reinterpret_cast<void**>(&h->d_coefs), cap * sizeof(int)));
    h->d_coefs_cap = cap;
  }
  if (h->h_coefs.size() > h->d_coefs_used) {
    CK(cudaMemcpyAsync(h->d_coefs + h->d_coefs_used, h->h_coefs.data() + h->d_coefs_used,
                       (h->h_coefs.size() - h->d_coefs_used) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    h->d_coefs_used = h->h_coefs.size();
  }
  CK(cudaMemcpyAsync(h->d_descs, h->h_descs, sizeof(CropDesc) * n, cudaMemcpyHostToDevice, h->stream));
  if (!jobs.empty()) {
    if (mask_bytes > h->masks_cap) {
      if (h->d_masks) cudaFree(h->d_masks);
      h->d_masks = nullptr;
      h->masks_cap = 0;
      const size_t cap = std::max<size_t>(mask_bytes + mask_bytes / 4, 4u << 20);
      CK(cudaMalloc(reinterpret_cast<void**>(&h->d_masks), cap));
      h->masks_cap = cap;
    }
    // jobs | edges | lines in one pageable upload (staged synchronously by the runtime)
    const size_t jb = round_up(static_cast<int>(jobs.size() * sizeof(MaskJob)), 16), eb = round_up(static_cast<int>(edges.size() * sizeof(MaskEdge)), 16),
                 lb = round_up(static_cast<int>(lines.size() * sizeof(MaskLine)), 16), meta = jb + eb + lb + 16;
    if (meta > h->mask_meta_cap) {
      if (h->d_mask_meta) cudaFree(h->d_mask_meta);
      h->d_mask_meta = nullptr;
      h->mask_meta_cap = 0;
      CK(cudaMalloc(&h->d_mask_meta, meta * 2));
      h->mask_meta_cap = meta * 2;
    }
    std::vector<uint8_t> blob(meta, 0);
    memcpy(blob.data(), jobs.data(), jobs.size() * sizeof(MaskJob));
    if (!edges.empty()) memcpy(blob.data() + jb, edges.data(), edges.size() * sizeof(MaskEdge));
    if (!lines.empty()) memcpy(blob.data() + jb + eb, lines.data(), lines.size() * sizeof(MaskLine));
    CK(cudaMemcpyAsync(h->d_mask_meta, blob.data(), meta, cudaMemcpyHostToDevice, h->stream));
    const MaskJob* dj = static_cast<const MaskJob*>(h->d_mask_meta);
    const MaskEdge* de = reinterpret_cast<const MaskEdge*>(static_cast<const uint8_t*>(h->d_mask_meta) + jb);
    const MaskLine* dl = reinterpret_cast<const MaskLine*>(static_cast<const uint8_t*>(h->d_mask_meta) + jb + eb);
    dim3 grid((max_mask_h + kMaskRowsPerCta - 1) / kMaskRowsPerCta, static_cast<unsigned>(jobs.size()));
    region_mask_fill_kernel<<<grid, 32 * kMaskRowsPerCta, 0, h->stream>>>(dj, de, h->d_masks);
    CK(cudaGetLastError());
    ++h->launches;
    if (!lines.empty()) {
      region_mask_lines_kernel<<<(static_cast<int>(lines.size()) + 127) / 128, 128, 0, h->stream>>>(dj, static_cast<int>(jobs.size()), dl,
                                                                                                  static_cast<int>(lines.size()), h->d_masks);
      CK(cudaGetLastError());
      ++h->launches;
    }
  }
  h->n = n;
  h->staged_ok = true;
  return MOCR_OK;
}

int preprocess(mocr_handle* h) {
  if (!h->staged_ok) return fail(h, MOCR_ERR_INVALID, "no crops staged");
  const size_t smem = pre_smem_bytes(h->pre_pitch, h->pre_tmp_rows);
  static size_t smem_set[16] = {};
  if (smem > 48 * 1024 && smem > smem_set[h->device & 15]) {
    CK(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    smem_set[h->device & 15] = 200 * 1024;
  }
  const bool tap = (h->taps & MOCR_TAP_PIXELS) != 0;
  const int i0 = h->sub_n > 0 ? h->sub_i0 : 0, cnt = h->sub_n > 0 ? h->sub_n : h->n;
  dim3 grid(kImage / kPreStripRows, cnt);
  preprocess_kernel<<<grid, kPreThreads, smem, h->stream>>>(h->d_arena, h->d_descs + i0, h->d_coefs, h->pre_bgr, h->pre_pitch, h->d_masks, h->patches.p,
                                                           tap ? h->px_u8 : nullptr, tap ? h->px_f32 : nullptr, h->lut);
  CK(cudaGetLastError());
  ++h->launches;
  h->pre_ok = true;
  h->enc_ok = h->dec_ok = false;
  return MOCR_OK;
}

// ------------------------------------------------------------------ encoder ---

int attention197(mocr_handle* h, int n) {
  static bool done[16] = {};
  if (!done[h->device & 15]) {
    CK(cudaFuncSetAttribute(encoder_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtcSmemBytes));
    done[h->device & 15] = true;
  }
  encoder_attention_tc_kernel<<<dim3(2, kHeads, n), kAtcThreads, kAtcSmemBytes, h->stream>>>(h->map_qkv_q, h->map_qkv_kv, h->ctx.p);
  CK(cudaGetLastError());
  ++h->launches;
  return MOCR_OK;
}

int encode_launches(mocr_handle* h) {
  const int i0 = h->sub_n > 0 ? h->sub_i0 : 0;
  const int n = h->sub_n > 0 ? h->sub_n : h->n, M = n * kEncTokens, bn = h->enc_bn, bn7 = h->enc_bn768;
  // embeddings: patch rows -> h[b*197+1+p] = conv + pos ; h[b*197] = cls + pos[0]   (modeling_vit.py:100-128)
  {
    GemmArgs a = out_f32(h->hres, kD);
    a.pos = h->pos;
    TRY(gemm(h, EPI_PATCH, bn7, h->patches, h->patch, n * kPatches, a));
    cls_rows_kernel<<<n, 192, 0, h->stream>>>(h->hres, h->cls, h->pos);
    CK(cudaGetLastError());
    ++h->launches;
  }
  for (int l = 0; l < kEncLayers; ++l) {
    EncLayer& L = h->enc[l];
    TRY(layernorm(h, h->hres, M, L.ln1, h->xn.p, nullptr));                                   // modeling_vit.py:333
    TRY(gemm(h, EPI_BF16, bn, h->xn, L.qkv, M, out_bf16(h->qkv, 3 * kD)));                   // :228-230
    TRY(attention197(h, n));                                                                  // :236-246
    TRY(gemm(h, h->resid_tma ? EPI_F32_ACCUM : EPI_F32_RESID, bn7, h->ctx, L.out, M, out_f32(h->hres, kD, h->hres, kD)));    // :266, :337
    TRY(layernorm(h, h->hres, M, L.ln2, h->xn.p, nullptr));                                   // :340
    TRY(gemm(h, EPI_BF16_GELU, bn, h->xn, L.fc1, M, out_bf16(h->mlp.p, kFFN)));              // :297-298
    TRY(gemm(h, h->resid_tma ? EPI_F32_ACCUM : EPI_F32_RESID, bn7, h->mlp, L.fc2, M, out_f32(h->hres, kD, h->hres, kD)));    // :309-311
  }
  TRY(layernorm(h, h->hres, M, h->enc_ln, h->enc_out.p, (h->taps & MOCR_TAP_ENCODER) ? h->enc_f32 : nullptr));   // :455
  // cross-attention K/V of both decoder layers, once per crop (modeling_bert.py:252-267)
  {
    GemmArgs a = out_bf16(h->crosskv + static_cast<size_t>(i0) * kEncTokens * 4 * kD, 4 * kD);
    a.crop_map = h->enc_crop_map;        // an admission's crops go to the cache blocks of their slots
    TRY(gemm(h, EPI_CROSSKV, bn, h->enc_out, h->cross_kv, M, a));
  }
  return MOCR_OK;
}

// The 88 encoder launches are host-bound when issued one by one (~6 us of CPU per launch with two
// 128-byte tensor maps as parameters): they are captured once per batch size and replayed as a graph.
int encode(mocr_handle* h) {
  if (!h->pre_ok) return fail(h, MOCR_ERR_INVALID, "preprocess has not run on the staged crops");
  if (!h->use_graph) {
    TRY(encode_launches(h));
  } else {
    const uint64_t key = (static_cast<uint64_t>(h->sub_n > 0 ? h->sub_n : h->n) << 32) | (static_cast<uint64_t>(h->sub_n > 0 ? h->sub_i0 : 0) << 8) |
                         ((h->taps & MOCR_TAP_ENCODER) ? 1u : 0u) | (h->enc_crop_map != nullptr ? 2u : 0u);
    auto it = h->enc_graphs.find(key);
    if (it == h->enc_graphs.end()) {
      const int64_t l0 = h->launches;
      TRY(encode_launches(h));            // warm run: function attributes, tensor maps (result is valid, kept)
      const int per = static_cast<int>(h->launches - l0);
      cudaGraph_t graph;
      cudaGraphExec_t exec = nullptr;
      CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
      const int r = encode_launches(h);
      h->launches = l0 + per;
      const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
      if (r != MOCR_OK) {
        if (ce == cudaSuccess) cudaGraphDestroy(graph);
        return r;
      }
      if (ce != cudaSuccess) return fail(h, MOCR_ERR_CUDA, "encoder stream capture failed: %s", cudaGetErrorString(ce));
      CK(cudaGraphInstantiate(&exec, graph, 0));
      cudaGraphDestroy(graph);
      if (h->enc_graphs.size() >= 160) {
        for (auto& g : h->enc_graphs) cudaGraphExecDestroy(g.second.exec);
        h->enc_graphs.clear();
      }
      h->enc_graphs[key] = mocr_handle::StepGraph{exec, per};
    } else {
      CK(cudaGraphLaunch(it->second.exec, h->stream));
      h->launches += it->second.launches;
    }
  }
  h->enc_ok = true;
  h->dec_ok = false;
  return MOCR_OK;
}

// ------------------------------------------------------------------ decoder ---

PdLinear pd_lin(const Linear& L) { return PdLinear{L.w, L.bias}; }
PdLn pd_ln(const LnParams& l) { return PdLn{l.g, l.b}; }

// n = decoder rows, n_crops = crops they work through (0: one crop per row)
// tile width of the large-batch program's vocabulary projection for n rows
int big_vocab_tile(const mocr_handle* h, int n) {
  if (h->big_vocab_bn != 0) return h->big_vocab_bn;
  const int m_tiles = (n + kGemmBM - 1) / kGemmBM;
  return m_tiles <= 1 ? 64 : (m_tiles <= 3 ? 128 : 256);
}

PdParams make_pd_params(mocr_handle* h, int n, int max_length, bool forced, bool tap, int n_crops = 0) {
  PdParams p{};
  p.B = n;
  p.n_crops = n_crops > n ? n_crops : n;
  p.lens = h->d_lens;
  p.slot_crop = h->d_slot_crop;
  p.queue = h->d_queue;
  p.max_len = max_length;
  p.cache_len = h->max_length;
  p.kv_div = 1;
  p.big = n > h->big_rows ? 1 : 0;
  p.n_partials = (p.big || (h->dec_tc & 1)) ? 2 * (kVocab / (p.big ? big_vocab_tile(h, n) : 64)) : kPdVocabTiles;
  p.logits_cur = 0;
  p.kv_evict_first = h->kv_evict_first >= 0 ? h->kv_evict_first : (p.big ? 3 : 2);
  p.fuse_ln = h->fuse_ln;
  p.big_accum = h->big_accum;
  p.kv_prefetch = h->kv_prefetch;
  p.eos_id = kSepId;
  for (int l = 0; l < kDecLayers; ++l) {
    DecLayer& L = h->dec[l];
    PdLayer& P = p.layer[l];
    P.qkv = pd_lin(L.self_qkv);
    P.self_out = pd_lin(L.self_out);
    P.cross_q = pd_lin(L.cross_q);
    P.cross_out = pd_lin(L.cross_out);
    P.fc1 = pd_lin(L.fc1);
    P.fc2 = pd_lin(L.fc2);
    P.ln_self = pd_ln(L.ln_self);
    P.ln_cross = pd_ln(L.ln_cross);
    P.ln_ffn = pd_ln(L.ln_ffn);
    P.self_k = h->self_k[l];
    P.self_v = h->self_v[l];
  }
  p.head_t = pd_lin(h->head_t);
  p.head_dec = pd_lin(h->head_dec);
  p.head_ln = pd_ln(h->head_ln);
  p.emb = h->emb;
  p.crosskv = h->crosskv;
  p.ids = h->d_ids;
  p.pos = h->d_pos;
  p.finished = h->d_finished;
  p.forced = forced ? h->d_forced : nullptr;
  p.x = h->d_x;
  p.xb = h->d_xb.p;
  p.tb = h->d_tb.p;
  p.q = h->d_q;
  p.y = h->d_y;
  p.yq = h->d_yq;
  p.qkv = h->d_qkv;
  p.ctx = h->d_ctx.p;
  p.ffn = h->d_ffn.p;
  p.part_max = h->part_max;
  p.part_idx = h->part_idx;
  p.logits = tap ? h->logits_tap : nullptr;
  p.prof = h->decode_prof ? h->d_prof : nullptr;
  return p;
}

// Launch with programmatic stream serialization: the kernel may start while its predecessor is
// still running; it synchronises with griddepcontrol.wait before touching dependent data.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(mocr_handle* h, void (*kernel)(KArgs...), int grid, int block, size_t smem, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = (h->use_pdl && h->pdl_now) ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl_grid(mocr_handle* h, void (*kernel)(KArgs...), dim3 grid, int block, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(block);
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = (h->use_pdl && h->pdl_now) ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// The encoder's tcgen05 GEMM as a decoder stage: launched with programmatic stream serialization (the kernel requests its
// weight tiles before griddepcontrol.wait), A map clipped to the live rows (TMA zero-fills rows >= n without reading them).
template <int BN, int EPI>
int launch_gemm_stage_tc(mocr_handle* h, const __nv_bfloat16* a_ptr, int K, Linear& L, int rows, GemmArgs a) {
  using Cfg = GemmCfg<BN>;
  static bool attr_done[16] = {};
  if (!attr_done[h->device & 15]) {
    CK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes(Cfg::kSmemBytes, EPI)));
    attr_done[h->device & 15] = true;
  }
  CUtensorMap ma;
  TRY(make_map(h, &ma, a_ptr, rows, K, kGemmBM));
  const CUtensorMap* mb;
  TRY(linear_map(h, &L, BN, &mb));
  a.M = rows;
  a.N = L.N;
  a.K = L.K;
  a.bias = L.bias;
  a.pdl = 1;
  if (EPI == EPI_F32_ACCUM) TRY(make_map_f32_out(h, &a.tmap_out, a.out, rows, L.N));
  if ((EPI == EPI_BF16 || EPI == EPI_BF16_GELU) && BN % 128 == 0 && h->enc_tma_store && a.ldo == L.N) {
    TRY(make_map_bf16_out(h, &a.tmap_out, a.out, rows, L.N));
    a.out_tma = 1;
  }
#ifdef MOCR_GEMM_DBG
  a.dbg = h->gemm_dbg;
#endif
  const int tiles = ((rows + kGemmBM - 1) / kGemmBM) * (L.N / BN);
  CK(launch_pdl(h, gemm_tcgen05_kernel<BN, EPI>, std::min(tiles, h->sms), kGemmThreads, gemm_smem_bytes(Cfg::kSmemBytes, EPI), ma, *mb, a));
  return MOCR_OK;
}

// The residual projections of the large-batch program with K split over clusters of four CTAs (gemm_ksplit.cuh).
int launch_gemm_ksplit(mocr_handle* h, const __nv_bfloat16* a_ptr, int K, Linear& L, int rows, float* x) {
  static bool attr_done[16] = {};
  if (!attr_done[h->device & 15]) {
    CK(cudaFuncSetAttribute(gemm_ksplit_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKsSmemBytes));
    attr_done[h->device & 15] = true;
  }
  CUtensorMap ma;
  TRY(make_map(h, &ma, a_ptr, rows, K, kGemmBM));
  const CUtensorMap* mb;
  TRY(linear_map(h, &L, kKsBN, &mb));
  GemmArgs a{};
  a.M = rows;
  a.N = L.N;
  a.K = L.K;
  a.bias = L.bias;
  a.out = x;
  a.ldo = L.N;
  a.pdl = 1;
  TRY(make_map_f32_out(h, &a.tmap_out, x, rows, L.N));
  const int tiles = ((rows + kGemmBM - 1) / kGemmBM) * (L.N / kKsBN);
  CK(launch_pdl(h, gemm_ksplit_accum_kernel, tiles * kKsSplit, kKsThreads, kKsSmemBytes, ma, *mb, a));
  return MOCR_OK;
}

Linear* dec_linear(mocr_handle* h, int lin) {
  if (lin == PD_LIN_HEAD_T) return &h->head_t;
  if (lin == PD_LIN_HEAD_DEC) return &h->head_dec;
  DecLayer& L = h->dec[lin / PD_LIN_PER_LAYER];
  switch (lin % PD_LIN_PER_LAYER) {
    case PD_LIN_QKV: return &L.self_qkv;
    case PD_LIN_SELF_OUT: return &L.self_out;
    case PD_LIN_CROSS_Q: return &L.cross_q;
    case PD_LIN_CROSS_OUT: return &L.cross_out;
    case PD_LIN_FC1: return &L.fc1;
    default: return &L.fc2;
  }
}

// A Linear of the large-batch program on the tcgen05 GEMM.  Tile widths are chosen for the tile count at a few hundred rows
// (128-row tiles): 32 columns for the N = 768 layers (96 tiles at 512 rows; 64-column tiles left 100 SMs idle: 16 us per
// launch), 64 for QKV (144 tiles), 128 for FFN1 (96 tiles).
int launch_tc_stage(mocr_handle* h, const PdParams& p, const PdStage& st) {
  Linear& L = *dec_linear(h, st.lin);
  const bool wide = h->big_bn768 == 64;
  switch (st.epi) {
    case EPI_BF16:
      if (st.N == kD && !wide) return launch_gemm_stage_tc<32, EPI_BF16>(h, st.A, st.K, L, p.B, out_bf16(st.ob, st.ldo));
      return launch_gemm_stage_tc<64, EPI_BF16>(h, st.A, st.K, L, p.B, out_bf16(st.ob, st.ldo));
    case EPI_BF16_GELU:
      return launch_gemm_stage_tc<128, EPI_BF16_GELU>(h, st.A, st.K, L, p.B, out_bf16(st.ob, st.ldo));
    case EPI_F32_RESID:
      if (!wide) return launch_gemm_stage_tc<32, EPI_F32_RESID>(h, st.A, st.K, L, p.B, out_f32(st.of, st.ldo, st.resid, kD));
      return launch_gemm_stage_tc<64, EPI_F32_RESID>(h, st.A, st.K, L, p.B, out_f32(st.of, st.ldo, st.resid, kD));
    case EPI_F32_ACCUM:      // x += A W^T + bias in place (TMA reduce-add: the SM never reads the residual)
      // (measured at 512 rows: FFN2, K = 3072, 17.4 -> 11.2 us; the K = 768 projections 6.6 -> 8.4 us: the cluster's fixed costs
      //  - two cluster barriers, the exchange - outweigh nine saved k-blocks, so only long K goes this way)
      if (h->big_ksplit && st.K >= 2048 && st.N % kKsBN == 0 && st.K % (kGemmBK * kKsSplit) == 0 &&
          ((p.B + kGemmBM - 1) / kGemmBM) * (st.N / kKsBN) * kKsSplit <= h->sms)
        return launch_gemm_ksplit(h, st.A, st.K, L, p.B, st.of);
      if (!wide) return launch_gemm_stage_tc<32, EPI_F32_ACCUM>(h, st.A, st.K, L, p.B, out_f32(st.of, st.ldo));
      return launch_gemm_stage_tc<64, EPI_F32_ACCUM>(h, st.A, st.K, L, p.B, out_f32(st.of, st.ldo));
    case EPI_F32_GELU:
      if (!wide) return launch_gemm_stage_tc<32, EPI_F32_GELU>(h, st.A, st.K, L, p.B, out_f32(st.of, st.ldo));
      return launch_gemm_stage_tc<64, EPI_F32_GELU>(h, st.A, st.K, L, p.B, out_f32(st.of, st.ldo));
    default:
      return fail(h, MOCR_ERR_INVALID, "decoder stage with an unsupported tcgen05 epilogue %d", st.epi);
  }
}

// The attention stage of the per-token program: the four-warp kernel (small program), or one warp per (row, head) unit
// (large-batch program), on the grid the options give.
cudaError_t launch_attention_stage(mocr_handle* h, const PdParams& p, const PdStage& st) {
  const int units = p.B * kHeads;
  const bool self = st.type == PD_ATTN_SELF;
  if (p.big && h->big_attn_rows) {
    const int grid = std::min(h->big_attn_grid > 0 ? h->big_attn_grid : units, (units + 3) / 4);
    return self ? launch_pdl(h, pd_attention_rows_kernel<true>, grid, 128, kPdAttnRowsSmemBytes, p, st)
                : launch_pdl(h, pd_attention_rows_kernel<false>, grid, 128, kPdAttnRowsSmemBytes, p, st);
  }
  const int want = p.big ? h->big_attn_grid : h->attn_grid;
  const int grid = want > 0 ? std::min(want, units) : units;
  return self ? launch_pdl(h, pd_attention_kernel<true>, grid, 128, kPdAttnSmemBytes, p, st)
              : launch_pdl(h, pd_attention_kernel<false>, grid, 128, kPdAttnSmemBytes, p, st);
}

// One greedy step as a sequence of stage kernels (decode_stages.cuh), one launch per stage.
int decode_stage_step(mocr_handle* h, const PdParams& p, bool skip_next = false) {
  static bool done[16] = {};
  if (!done[h->device & 15]) {
    CK(cudaFuncSetAttribute(pd_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPdAttnSmemBytes));
    CK(cudaFuncSetAttribute(pd_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPdAttnSmemBytes));
    CK(cudaFuncSetAttribute(pd_attention_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPdAttnRowsSmemBytes));
    CK(cudaFuncSetAttribute(pd_attention_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPdAttnRowsSmemBytes));
    CK(cudaFuncSetAttribute(pd_gemm_kernel<48, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, pd_gemm_smem_bytes(48)));
    CK(cudaFuncSetAttribute(pd_proj_ln_kernel<8, 3, 1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    if (h->carveout >= 0) {
      // one shared-memory carve-out for every stage kernel: kernels with different L1/smem splits cannot overlap on an SM
      CK(cudaFuncSetAttribute(pd_attention_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, h->carveout));
      CK(cudaFuncSetAttribute(pd_attention_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, h->carveout));
      CK(cudaFuncSetAttribute(pd_gemm_kernel<16, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, h->carveout));
      CK(cudaFuncSetAttribute(pd_gemm_kernel<32, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, h->carveout));
      CK(cudaFuncSetAttribute(pd_gemm_kernel<48, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, h->carveout));
      CK(cudaFuncSetAttribute(pd_proj_ln_kernel<8, 3, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, h->carveout));
      CK(cudaFuncSetAttribute(pd_ln_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, h->carveout));
      CK(cudaFuncSetAttribute(pd_next_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, h->carveout));
      CK(cudaFuncSetAttribute(pd_begin_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, h->carveout));
    }
    done[h->device & 15] = true;
  }
  PdStage prog[kPdMaxStages];
  const int n_stages = pd_build_program(p, prog);
  // row-wise stages (LayerNorm, next token): one warp per row; few warps per CTA so that the rows spread over
  // many SMs (8 CTAs of 8 warps made 8 SMs pull 98 KB each at ~75 GB/s per SM: 1.3 us of the stage's 2.6)
  const int row_warps = std::max(1, std::min(p.big ? h->big_row_warps : h->row_warps, kPdWarps));
  const int row_ctas = (p.B + row_warps - 1) / row_warps;
  struct PdlReset { mocr_handle* h; ~PdlReset() { h->pdl_now = 1; } } pdl_reset{h};
  for (int i = 0; i < n_stages; ++i) {
    const PdStage& st = prog[i];
    h->pdl_now = (h->pdl_mask >> st.type) & 1;
    if (st.type == PD_TC) {
      TRY(launch_tc_stage(h, p, st));
    } else if (st.type == PD_GEMM16 || st.type == PD_GEMM32 || st.type == PD_GEMM48) {
      if ((st.epi == PD_ARGMAX) && (p.big || (h->dec_tc & 1))) {
        // vocabulary projection + per-tile arg-max on the tcgen05 kernel (96 CTAs, weights by TMA before the dependency wait)
        GemmArgs a{};
        a.part_max = p.part_max;
        a.part_idx = p.part_idx;
        a.logits = p.logits;
        a.step = p.logits_cur ? h->d_zero : p.pos;          // beam mode taps the current step only: [B, 6144]
        a.tap_steps = p.logits_cur ? 1 : p.max_len - 1;
        const int vbn = p.big ? big_vocab_tile(h, p.B) : 64;
        if (vbn == 256) TRY((launch_gemm_stage_tc<256, EPI_ARGMAX>(h, st.A, kD, h->head_dec, p.B, a)));
        else if (vbn == 128) TRY((launch_gemm_stage_tc<128, EPI_ARGMAX>(h, st.A, kD, h->head_dec, p.B, a)));
        else TRY((launch_gemm_stage_tc<64, EPI_ARGMAX>(h, st.A, kD, h->head_dec, p.B, a)));
      } else {
        const int nt = st.type == PD_GEMM16 ? 16 : (st.type == PD_GEMM32 ? 32 : 48);
        const int grid = (st.N / nt) * st.ksplit;
        if (st.type == PD_GEMM16) CK(launch_pdl(h, pd_gemm_kernel<16, 2>, grid, 128 * kPdStageKS, pd_gemm_smem_bytes(16), p, st));
        else if (st.type == PD_GEMM32) CK(launch_pdl(h, pd_gemm_kernel<32, 2>, grid, 128 * kPdStageKS, pd_gemm_smem_bytes(32), p, st));
        else CK(launch_pdl(h, pd_gemm_kernel<48, 1>, grid, 128 * kPdStageKS, pd_gemm_smem_bytes(48), p, st));
      }
    } else if (st.type == PD_PROJ_LN) {
      const int groups = (p.B + kPlRows - 1) / kPlRows;
      CK(launch_pdl(h, pd_proj_ln_kernel<8, 3, 1>, groups * kPlCluster, 256, pd_proj_ln_smem_bytes(8), p, st));
    } else if (st.type == PD_ATTN_SELF || st.type == PD_ATTN_CROSS) {
      CK(launch_attention_stage(h, p, st));
    } else if (st.type == PD_LN) {
      CK(launch_pdl(h, pd_ln_kernel, row_ctas, 32 * row_warps, 0, p, st));
    } else {
      if (skip_next) continue;
      CK(launch_pdl(h, pd_next_kernel, row_ctas, 32 * row_warps, 0, p));
    }
    ++h->launches;
  }
  return MOCR_OK;
}

// slot_crop[r] = r for r < rows: rows that address their own crop (unit hooks, kernel timing, beam mode)
int identity_slots(mocr_handle* h, int rows) {
  std::vector<int> iota(rows);
  for (int i = 0; i < rows; ++i) iota[i] = i;
  CK(cudaMemcpyAsync(h->d_slot_crop, iota.data(), sizeof(int) * rows, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));      // (pageable source)
  return MOCR_OK;
}

int decode_begin(mocr_handle* h, const PdParams& pdp) {
  CK(launch_pdl(h, pd_begin_kernel, std::min((pdp.n_crops + kPdWarps - 1) / kPdWarps, 4 * h->sms), kPdThreads, 0, pdp));
  ++h->launches;
  return MOCR_OK;
}

// The CUDA graph of steps_per_graph token steps for this (crops, rows, length, mode), created on first use.  Creation
// runs one step outside capture (function attributes, tensor maps) on throw-away state and re-runs decode_begin.
int decode_step_graph(mocr_handle* h, const PdParams& pdp, bool forced, bool tap, cudaGraphExec_t* exec, int64_t* per_step) {
  *exec = nullptr;
  *per_step = 0;
  if (!h->use_graph) return MOCR_OK;
  const int n = pdp.n_crops, rows = pdp.B, max_length = pdp.max_len;
  const uint64_t key = (static_cast<uint64_t>(n) << 40) | (static_cast<uint64_t>(rows) << 20) | (static_cast<uint64_t>(max_length) << 8) |
                       (pdp.queue_slots != nullptr ? 4u : 0u) | (forced ? 2u : 0u) | (tap ? 1u : 0u);
  const int spg = std::max(1, std::min(h->steps_per_graph, max_length - 1));
  auto it = h->graphs.find(key);
  if (it != h->graphs.end()) {
    *exec = it->second.exec;
    *per_step = it->second.launches;
    return MOCR_OK;
  }
  PdParams warm = pdp;
  warm.ext_queue = 0;
  const int64_t l0 = h->launches;
  TRY(decode_stage_step(h, pdp));
  *per_step = h->launches - l0;
  TRY(decode_begin(h, warm));
  if (tap) CK(cudaMemsetAsync(h->logits_tap, 0, static_cast<size_t>(n) * (max_length - 1) * kVocab * sizeof(float), h->stream));
  cudaGraph_t graph;
  CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  const int64_t l1 = h->launches;
  int r = MOCR_OK;
  for (int s = 0; s < spg && r == MOCR_OK; ++s) r = decode_stage_step(h, pdp);
  h->launches = l1;
  *per_step *= spg;
  cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
  if (r != MOCR_OK) {
    if (ce == cudaSuccess) cudaGraphDestroy(graph);
    return r;
  }
  if (ce != cudaSuccess) return fail(h, MOCR_ERR_CUDA, "stream capture failed: %s", cudaGetErrorString(ce));
  CK(cudaGraphInstantiate(exec, graph, 0));
  cudaGraphDestroy(graph);
  if (h->graphs.size() >= 64) {
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.second.exec);
    h->graphs.clear();
  }
  h->graphs[key] = mocr_handle::StepGraph{*exec, static_cast<int>(*per_step)};
  return MOCR_OK;
}

// Decoder rows for n crops: one per crop, or fewer (option "slots"): a row that finishes then takes the next waiting crop.
// Teacher forcing and the logits tap address rows by crop, so they keep one row per crop.
int decode_rows(const mocr_handle* h, int n, bool forced, bool tap) { return (h->slots > 0 && !forced && !tap) ? std::min(n, h->slots) : n; }

// Token steps until every crop of the decode has finished (polled every check_every steps).
int decode_loop(mocr_handle* h, const PdParams& pdp, bool forced, cudaGraphExec_t exec, int64_t per_step, cudaStream_t feeder = nullptr) {
  const int n = pdp.n_crops, rows = pdp.B, max_length = pdp.max_len;
  // every crop takes at most max_length - 1 steps of one row: an upper bound of the step count for any length mix
  const int steps = (max_length - 1) * ((n + rows - 1) / rows) + (pdp.ext_queue ? 4 * h->check_every : 0);
  // A graph replays steps_per_graph steps; steps past the end are no-ops for the result
  // (every row is finished by then: no id, position or cache row is written).
  const int spg = exec != nullptr ? std::max(1, std::min(h->steps_per_graph, max_length - 1)) : 1;
  int done_steps = 0, budget = steps;
  bool finished = false, extended = false;
  while (!finished) {
    if (done_steps >= budget) {
      // fed by the encoder stream: the bound only holds once every crop has been published (idle steps before that do not count)
      if (!pdp.ext_queue || extended) break;
      if (feeder != nullptr) CK(cudaStreamSynchronize(feeder));
      budget = done_steps + steps;
      extended = true;
    }
    const int chunk = forced ? budget - done_steps : std::min(std::max(h->check_every, spg), budget - done_steps);
    int ran = 0;
    while (ran < chunk) {
      if (exec != nullptr) {
        CK(cudaGraphLaunch(exec, h->stream));
        h->launches += per_step;
      } else {
        TRY(decode_stage_step(h, pdp));
      }
      ran += spg;
    }
    done_steps += ran;
    if (!forced && (done_steps < budget || pdp.ext_queue)) {
      // every crop finished? (generation/utils.py:2805 does this check, with a host sync, every step)
      CK(cudaMemcpyAsync(h->h_queue, h->d_queue, sizeof(int) * 4, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      finished = h->h_queue[2] >= n;
    }
  }
  h->last_steps = std::min(done_steps, budget);
  h->last_rows = rows;
  h->cur_len = max_length;
  h->dec_ok = true;
  return MOCR_OK;
}

int decode(mocr_handle* h, int max_length, const int32_t* forced_ids) {
  if (h->sess_on) return fail(h, MOCR_ERR_INVALID, "a session is active on this handle (mocr_session_end first)");
  if (!h->enc_ok) return fail(h, MOCR_ERR_INVALID, "encode has not run on the staged crops");
  if (max_length < 2 || max_length > h->max_length)
    return fail(h, MOCR_ERR_CAPACITY, "max_length %d outside [2, %d]", max_length, h->max_length);
  const int n = h->n;
  const bool forced = forced_ids != nullptr;
  const bool tap = (h->taps & MOCR_TAP_LOGITS) != 0;
  if (forced) CK(cudaMemcpyAsync(h->d_forced, forced_ids, sizeof(int) * n * max_length, cudaMemcpyHostToDevice, h->stream));
  if (tap) {
    const size_t need_b = static_cast<size_t>(n) * (max_length - 1) * kVocab * sizeof(float);
    if (need_b > h->logits_tap_bytes) {
      CK(cudaStreamSynchronize(h->stream));
      if (h->logits_tap) cudaFree(h->logits_tap);
      h->logits_tap = nullptr;
      h->logits_tap_bytes = 0;
      CK(cudaMalloc(reinterpret_cast<void**>(&h->logits_tap), need_b));
      h->logits_tap_bytes = need_b;
      for (auto& g : h->graphs) cudaGraphExecDestroy(g.second.exec);   // they captured the old tap pointer
      h->graphs.clear();
    }
    CK(cudaMemsetAsync(h->logits_tap, 0, need_b, h->stream));
  }
  if (h->decode_prof) CK(cudaMemsetAsync(h->d_prof, 0, 8, h->stream));
  const PdParams pdp = make_pd_params(h, decode_rows(h, n, forced, tap), max_length, forced, tap, n);
  TRY(decode_begin(h, pdp));
  cudaGraphExec_t exec = nullptr;
  int64_t per_step = 0;
  TRY(decode_step_graph(h, pdp, forced, tap, &exec, &per_step));
  return decode_loop(h, pdp, forced, exec, per_step);
}

// Slot refill with the encoder overlapped: all m crops are staged at once; the first sub-chunk is preprocessed and encoded,
// then the decoder starts on its own (higher-priority) stream while the encoder stream works through the remaining
// sub-chunks and publishes them to the device-side queue; rows that finish (or idle rows) pick them up.
int recognize_pipelined(mocr_handle* h, const mocr_crop_t* crops, int m, int order, int max_length) {
  if (max_length < 2 || max_length > h->max_length)
    return fail(h, MOCR_ERR_CAPACITY, "max_length %d outside [2, %d]", max_length, h->max_length);
  const int rows = std::min(m, h->slots);
  const int sub = std::max(rows, 64);                   // crops per encoder sub-chunk
  // the decode graph first: creating it runs a warm-up step that may not touch the live queue
  PdParams pdp = make_pd_params(h, rows, max_length, false, false, m);
  cudaGraphExec_t exec = nullptr;
  int64_t per_step = 0;
  h->n = m;
  TRY(decode_begin(h, pdp));
  TRY(decode_step_graph(h, pdp, false, false, &exec, &per_step));
  CK(cudaStreamSynchronize(h->stream));
  cudaStream_t dec_stream = h->stream;
  cudaStream_t enc_stream = h->pipeline == 2 ? h->stream_enc_hi : h->stream_enc;
  struct Restore {      // also on an error path: the handle's stream is the decoder stream again and the encoder stream has drained
    mocr_handle* h; cudaStream_t s; cudaStream_t e;
    ~Restore() { h->stream = s; h->sub_i0 = 0; h->sub_n = 0; cudaStreamSynchronize(e); }
  } restore{h, dec_stream, enc_stream};
  h->stream = enc_stream;                               // everything encoder-side goes to the encoder stream
  TRY(stage_crops(h, crops, m, order));
  for (int i0 = 0; i0 < m; i0 += sub) {
    h->sub_i0 = i0;
    h->sub_n = std::min(sub, m - i0);
    TRY(preprocess(h));
    TRY(encode(h));
    pd_publish_kernel<<<1, 1, 0, h->stream>>>(h->d_queue, i0 == 0 ? rows : -1, i0 + h->sub_n);
    CK(cudaGetLastError());
    ++h->launches;
    if (i0 == 0) {
      CK(cudaEventRecord(h->ev_first, h->stream));
      // the decoder starts as soon as the first sub-chunk is ready; its launches are issued from here on while the
      // host keeps feeding the encoder stream below
      CK(cudaStreamWaitEvent(dec_stream, h->ev_first, 0));
    }
  }
  h->stream = dec_stream;
  h->sub_i0 = h->sub_n = 0;
  pdp.ext_queue = 1;
  TRY(decode_begin(h, pdp));
  TRY(decode_loop(h, pdp, false, exec, per_step, enc_stream));
  CK(cudaStreamSynchronize(enc_stream));
  return MOCR_OK;
}

int fetch_ids(mocr_handle* h, int32_t* out_ids, int32_t* out_lens) {
  if (!h->dec_ok) return fail(h, MOCR_ERR_INVALID, "no decode result to fetch");
  const int n = h->n, T = h->cur_len;
  CK(cudaMemcpyAsync(out_ids, h->d_ids, sizeof(int) * n * T, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->h_flags, h->d_lens, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->h_flags + n, h->d_pos, sizeof(int) * std::min(n, std::max(h->last_rows, 1)), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (out_lens != nullptr)
    for (int i = 0; i < n; ++i) {
      // a crop that finished has its length recorded; a teacher-forced row never "finishes" before max_length: position + 1
      const int by_pos = i < h->last_rows ? std::min(h->h_flags[n + i] + 1, T) : T;
      out_lens[i] = h->h_flags[i] > 0 ? h->h_flags[i] : by_pos;
    }
  return MOCR_OK;
}

int check_handle(mocr_handle* h) {
  if (h == nullptr) return fail(nullptr, MOCR_ERR_INVALID, "handle is NULL");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return fail(h, MOCR_ERR_CUDA, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e));
  return MOCR_OK;
}

int create_impl(mocr_handle* h) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(h, MOCR_ERR_NO_DEVICE, "no CUDA device (%s); this engine has no CPU path", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
  if (h->device < 0 || h->device >= count) return fail(h, MOCR_ERR_NO_DEVICE, "device %d out of range (%d devices)", h->device, count);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->device));
  if (prop.major != 10)
    return fail(h, MOCR_ERR_NO_DEVICE, "device %d is sm_%d%d; the kernels are built for sm_100a (B200) only", h->device, prop.major, prop.minor);
  CK(cudaSetDevice(h->device));
  h->sms = prop.multiProcessorCount;
  if (const char* e = getenv("MOCR_DEC_TC")) h->dec_tc = atoi(e) & 7;     // (A/B switch for the test-suite)
  {
    int lo = 0, hi = 0;     // (numerically lower = higher priority)
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, hi));
    CK(cudaStreamCreateWithPriority(&h->stream_enc, cudaStreamNonBlocking, lo));
    CK(cudaStreamCreateWithPriority(&h->stream_enc_hi, cudaStreamNonBlocking, hi));
    CK(cudaStreamCreateWithPriority(&h->stream_fetch, cudaStreamNonBlocking, hi));
    CK(cudaEventCreateWithFlags(&h->ev_first, cudaEventDisableTiming));
  }
  const int B = h->max_batch, T = h->max_length;
  h->rows_cap = round_up(B * kEncTokens, kGemmBM);
  h->brow_cap = round_up(B, kGemmBM);
  TRY(make_act(h, &h->patches, round_up(B * kPatches, kGemmBM), kPatchK));
  TRY(make_act(h, &h->xn, h->rows_cap, kD));
  TRY(make_act(h, &h->ctx, h->rows_cap, kD));
  TRY(make_act(h, &h->mlp, h->rows_cap, kFFN));
  TRY(make_act(h, &h->enc_out, h->rows_cap, kD));
  TRY(dmalloc(h, &h->qkv, static_cast<size_t>(h->rows_cap) * 3 * kD));
  TRY(make_map(h, &h->map_qkv_q, h->qkv, h->rows_cap, 3 * kD, 128));
  TRY(make_map(h, &h->map_qkv_kv, h->qkv, h->rows_cap, 3 * kD, kAtcKeys));
  TRY(dmalloc(h, &h->hres, static_cast<size_t>(h->rows_cap) * kD));
  TRY(dmalloc(h, &h->crosskv, static_cast<size_t>(h->rows_cap) * 4 * kD));
  TRY(make_act(h, &h->d_xb, h->brow_cap, kD));
  TRY(make_act(h, &h->d_ctx, h->brow_cap, kD));
  TRY(make_act(h, &h->d_ffn, h->brow_cap, kFFN));
  TRY(make_act(h, &h->d_tb, h->brow_cap, kD));
  TRY(dmalloc(h, &h->d_x, static_cast<size_t>(h->brow_cap) * kD));
  TRY(dmalloc(h, &h->d_y, static_cast<size_t>(h->brow_cap) * kD * kPdSplit));
  TRY(dmalloc(h, &h->d_yq, static_cast<size_t>(h->brow_cap) * kD * kPdSplit));
  TRY(dmalloc(h, &h->d_qkv, static_cast<size_t>(h->brow_cap) * 3 * kD));
  TRY(dmalloc(h, &h->d_q, static_cast<size_t>(h->brow_cap) * kD));
  for (int l = 0; l < kDecLayers; ++l) {
    TRY(dmalloc(h, &h->self_k[l], static_cast<size_t>(B) * T * kD));
    TRY(dmalloc(h, &h->self_v[l], static_cast<size_t>(B) * T * kD));
  }
  TRY(dmalloc(h, &h->part_max, static_cast<size_t>(h->brow_cap) * 2 * (kVocab / 32)));
  TRY(dmalloc(h, &h->part_idx, static_cast<size_t>(h->brow_cap) * 2 * (kVocab / 32)));
  TRY(dmalloc(h, &h->d_ids, static_cast<size_t>(B) * T));
  TRY(dmalloc(h, &h->d_forced, static_cast<size_t>(B) * T));
  TRY(dmalloc(h, &h->d_pos, static_cast<size_t>(B)));
  TRY(dmalloc(h, &h->d_finished, static_cast<size_t>(B)));
  TRY(dmalloc(h, &h->d_zero, static_cast<size_t>(B)));
  TRY(dmalloc(h, &h->d_lens, static_cast<size_t>(B)));
  TRY(dmalloc(h, &h->d_slot_crop, static_cast<size_t>(B)));
  TRY(dmalloc(h, &h->d_queue, 4));
  CK(cudaMallocHost(reinterpret_cast<void**>(&h->h_queue), sizeof(int) * 4));
  TRY(dmalloc(h, &h->d_prof, 4096));
  TRY(dmalloc(h, &h->d_descs, static_cast<size_t>(B)));
  CK(cudaMallocHost(reinterpret_cast<void**>(&h->h_flags), sizeof(int) * 2 * B));
  CK(cudaMallocHost(reinterpret_cast<void**>(&h->h_descs), sizeof(CropDesc) * B));
  CK(cudaStreamSynchronize(h->stream));
  return MOCR_OK;
}

int ensure_taps(mocr_handle* h) {
  const size_t B = h->max_batch;
  if ((h->taps & MOCR_TAP_PIXELS) && h->px_u8 == nullptr) {
    TRY(dmalloc(h, &h->px_u8, B * kImage * kImage));
    TRY(dmalloc(h, &h->px_f32, B * kImage * kImage));
  }
  if ((h->taps & MOCR_TAP_ENCODER) && h->enc_f32 == nullptr) TRY(dmalloc(h, &h->enc_f32, static_cast<size_t>(h->rows_cap) * kD));
  return MOCR_OK;
}

int d2h(mocr_handle* h, void* dst, const void* src, size_t bytes) {
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return MOCR_OK;
}

// No C++ exception may cross the C ABI (std::bad_alloc from a staging vector, std::system_error from a thread).
template <class F>
int guarded(mocr_handle* h, F&& f) {
  try {
    return f();
  } catch (const std::bad_alloc&) {
    return fail(h, MOCR_ERR_CAPACITY, "out of host memory");
  } catch (const std::exception& e) {
    return fail(h, MOCR_ERR_INVALID, "unexpected exception: %s", e.what());
  } catch (...) {
    return fail(h, MOCR_ERR_INVALID, "unexpected exception");
  }
}

}  // namespace

// ===================================================================== C ABI ===

extern "C" {

int mocr_abi_version(void) { return MOCR_ABI_VERSION; }

int mocr_create(int device, int max_batch, int max_length, mocr_handle_t** out) {
  if (out == nullptr) return fail(nullptr, MOCR_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (max_batch < 1 || max_batch > 4096) return fail(nullptr, MOCR_ERR_INVALID, "max_batch %d outside [1, 4096]", max_batch);
  if (max_length < 2 || max_length > kMaxPos) return fail(nullptr, MOCR_ERR_INVALID, "max_length %d outside [2, %d]", max_length, kMaxPos);
  mocr_handle* h = new (std::nothrow) mocr_handle();
  if (h == nullptr) return fail(nullptr, MOCR_ERR_INVALID, "out of host memory");
  h->device = device;
  h->max_batch = max_batch;
  h->max_length = max_length;
  int r = create_impl(h);
  if (r != MOCR_OK) {
    g_create_error = h->error;
    mocr_destroy(h);
    return r;
  }
  *out = h;
  return MOCR_OK;
}

int mocr_destroy(mocr_handle_t* h) {
  if (h == nullptr) return MOCR_OK;
  { std::lock_guard<std::mutex> lock(h->mu); }      // a call still running on another thread finishes first (callers must not start new ones)
  if (cudaSetDevice(h->device) == cudaSuccess) {
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->stream_enc) cudaStreamSynchronize(h->stream_enc);
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.second.exec);
    for (auto& g : h->enc_graphs) cudaGraphExecDestroy(g.second.exec);
    for (auto& g : h->beam_graphs) cudaGraphExecDestroy(g.second.exec);
    for (void* p : h->allocs) cudaFree(p);
    if (h->d_beam_dev) cudaFree(h->d_beam_dev);
    if (h->h_beam_ctl) cudaFreeHost(h->h_beam_ctl);
    if (h->logits_tap) cudaFree(h->logits_tap);
    for (int i = 0; i < 2; ++i) {
      if (h->h_sess_lens[i]) cudaFreeHost(h->h_sess_lens[i]);
      if (h->sess_ev[i]) cudaEventDestroy(h->sess_ev[i]);
    }
    if (h->d_arena) cudaFree(h->d_arena);
    if (h->h_arena) cudaFreeHost(h->h_arena);
    if (h->d_coefs) cudaFree(h->d_coefs);
    if (h->d_masks) cudaFree(h->d_masks);
    if (h->d_mask_meta) cudaFree(h->d_mask_meta);
    if (h->d_beam) cudaFree(h->d_beam);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    if (h->h_descs) cudaFreeHost(h->h_descs);
    if (h->h_queue) cudaFreeHost(h->h_queue);
    if (h->stream_enc) cudaStreamDestroy(h->stream_enc);
    if (h->stream_enc_hi) cudaStreamDestroy(h->stream_enc_hi);
    if (h->stream_fetch) cudaStreamDestroy(h->stream_fetch);
    if (h->ev_first) cudaEventDestroy(h->ev_first);
    if (h->ev_staged) cudaEventDestroy(h->ev_staged);
    if (h->ev_published) cudaEventDestroy(h->ev_published);
    if (h->stream) cudaStreamDestroy(h->stream);
  }
  delete h;
  return MOCR_OK;
}

int mocr_set_weight(mocr_handle_t* h, const char* name, const float* data, const int64_t* shape, int ndim) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (name == nullptr || data == nullptr || shape == nullptr || ndim < 1 || ndim > 8) return fail(h, MOCR_ERR_INVALID, "bad tensor argument");
  if (h->finalized) return fail(h, MOCR_ERR_INVALID, "weights are already finalized");
  HostTensor t;
  size_t numel = 1;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] < 1) return fail(h, MOCR_ERR_INVALID, "tensor %s has a non-positive extent", name);
    t.shape.push_back(shape[i]);
    numel *= static_cast<size_t>(shape[i]);
  }
  if (numel > (1u << 28)) return fail(h, MOCR_ERR_INVALID, "tensor %s is implausibly large", name);
  t.data.assign(data, data + numel);
  h->staged[name] = std::move(t);
  return MOCR_OK;
}

int mocr_finalize_weights(mocr_handle_t* h) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (h->finalized) return fail(h, MOCR_ERR_INVALID, "weights are already finalized");
  return finalize_weights(h);
}

int mocr_stage_crops(mocr_handle_t* h, const mocr_crop_t* crops, int n, int channel_order) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  return guarded(h, [&]() -> int { return stage_crops(h, crops, n, channel_order); });
}

int mocr_stage_regions(mocr_handle_t* h, const mocr_crop_t* page, const mocr_region_t* regions, int n, int channel_order) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  return guarded(h, [&]() -> int { return stage_regions(h, page, regions, n, channel_order); });
}

int mocr_recognize_regions(mocr_handle_t* h, const mocr_crop_t* page, const mocr_region_t* regions, int n, int channel_order, int max_length,
                           int32_t* out_ids, int32_t* out_lens) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (n < 0 || (n > 0 && (page == nullptr || regions == nullptr || out_ids == nullptr))) return fail(h, MOCR_ERR_INVALID, "bad argument");
  return guarded(h, [&]() -> int {
    for (int i0 = 0; i0 < n; i0 += h->max_batch) {
      const int m = std::min(h->max_batch, n - i0);
      TRY(stage_regions(h, page, regions + i0, m, channel_order));
      TRY(preprocess(h));
      TRY(encode(h));
      TRY(decode(h, max_length, nullptr));
      TRY(fetch_ids(h, out_ids + static_cast<size_t>(i0) * max_length, out_lens ? out_lens + i0 : nullptr));
    }
    return MOCR_OK;
  });
}

int mocr_get_region_mask(mocr_handle_t* h, int index, uint8_t* out) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->staged_ok || index < 0 || index >= static_cast<int>(h->region_mask_off.size()) || index >= h->n || out == nullptr)
    return fail(h, MOCR_ERR_INVALID, "no staged region %d", index);
  if (h->region_mask_off[index] < 0) return fail(h, MOCR_ERR_INVALID, "region %d has no polygon", index);
  return d2h(h, out, h->d_masks + h->region_mask_off[index], static_cast<size_t>(h->region_mask_hw[2 * index]) * h->region_mask_hw[2 * index + 1]);
}

int mocr_preprocess(mocr_handle_t* h) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  return guarded(h, [&]() -> int { return preprocess(h); });
}

int mocr_encode(mocr_handle_t* h) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  return guarded(h, [&]() -> int { return encode(h); });
}

int mocr_decode_greedy(mocr_handle_t* h, int max_length, const int32_t* forced_ids) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  return guarded(h, [&]() -> int { return decode(h, max_length, forced_ids); });
}

int mocr_fetch_ids(mocr_handle_t* h, int32_t* out_ids, int32_t* out_lens) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (out_ids == nullptr) return fail(h, MOCR_ERR_INVALID, "out_ids is NULL");
  return fetch_ids(h, out_ids, out_lens);
}

int mocr_run_resident(mocr_handle_t* h, int max_length) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  TRY(preprocess(h));
  TRY(encode(h));
  return decode(h, max_length, nullptr);
}

// Stage, preprocess and encode m crops.  Batches of at least two chunks (option "stage_chunk", 128 crops: 197 m-tiles fill the
// encoder GEMMs' waves exactly) are copied chunk by chunk on the copy stream while the encoder already works on the chunks that
// have arrived; smaller ones go in one piece.
int stage_encode(mocr_handle* h, const mocr_crop_t* crops, int m, int order) {
  const int chunk = h->stage_chunk;
  if (chunk <= 0 || m < 2 * chunk || h->taps != 0) {
    TRY(stage_crops(h, crops, m, order));
    TRY(preprocess(h));
    return encode(h);
  }
  struct Restore {
    mocr_handle* h;
    ~Restore() { h->sub_i0 = h->sub_n = 0; }
  } restore{h};
  const std::function<int(int, int)> after = [h](int i0, int cnt) -> int {
    h->sub_i0 = i0;
    h->sub_n = cnt;
    TRY(preprocess(h));
    return encode(h);
  };
  TRY(stage_crops(h, crops, m, order, chunk, &after));
  h->sub_i0 = h->sub_n = 0;
  return MOCR_OK;          // pre_ok / enc_ok were set by the last chunk; every chunk's work is ordered on the handle's stream
}

int mocr_recognize(mocr_handle_t* h, const mocr_crop_t* crops, int n, int channel_order, int max_length, int32_t* out_ids,
                   int32_t* out_lens) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (n < 0 || (n > 0 && (crops == nullptr || out_ids == nullptr))) return fail(h, MOCR_ERR_INVALID, "bad argument");
  return guarded(h, [&]() -> int {
    for (int i0 = 0; i0 < n; i0 += h->max_batch) {
      const int m = std::min(h->max_batch, n - i0);
      if (h->slots > 0 && h->pipeline && m > h->slots && h->taps == 0) {
        TRY(recognize_pipelined(h, crops + i0, m, channel_order, max_length));
      } else {
        TRY(stage_encode(h, crops + i0, m, channel_order));
        TRY(decode(h, max_length, nullptr));
      }
      TRY(fetch_ids(h, out_ids + static_cast<size_t>(i0) * max_length, out_lens ? out_lens + i0 : nullptr));
    }
    return MOCR_OK;
  });
}

// ---- admission into a running decode (include/mocr_b200.h: mocr_session_*) ------------------------------------------
// A session keeps `rows` decoder rows stepping; crops are added while it runs: each takes a free SLOT (crop index: encoder K/V,
// id row, length), is preprocessed and encoded between two step chunks and published to the device queue, where the first idle
// row picks it up (pd_next_kernel).  The host reads the slots' lengths after every chunk, fetches the rows that finished and
// releases their slots for reuse: no batch boundary, a caller waits for its own crop only.  ids per crop are those of
// mocr_recognize (rows are independent).

constexpr int kSessSmallRows = 16;

int session_end(mocr_handle* h) {
  h->sess_on = false;
  if (getenv("MOCR_SESSION_PROF") != nullptr && h->prof_add[3] > 0)
    fprintf(stderr, "[mocr] session: %.0f encoder passes; in mocr_session_add: %.1f ms waiting for the previous pass, %.1f ms staging, %.1f ms launching\n",
            h->prof_add[3], h->prof_add[0] * 1e3, h->prof_add[1] * 1e3, h->prof_add[2] * 1e3);
  h->prof_add[0] = h->prof_add[1] = h->prof_add[2] = h->prof_add[3] = 0;
  h->sub_i0 = h->sub_n = 0;
  h->staged_ok = h->pre_ok = h->enc_ok = h->dec_ok = false;
  cudaStreamSynchronize(h->stream_enc);
  cudaStreamSynchronize(h->stream_enc_hi);
  cudaStreamSynchronize(h->stream_fetch);
  cudaStreamSynchronize(h->stream);
  return MOCR_OK;
}

int mocr_session_begin(mocr_handle_t* h, int channel_order, int max_length, int rows) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  return guarded(h, [&]() -> int {
    if (!h->finalized) return fail(h, MOCR_ERR_INVALID, "weights are not finalized");
    if (h->sess_on) return fail(h, MOCR_ERR_INVALID, "a session is already active");
    if (h->taps != 0) return fail(h, MOCR_ERR_INVALID, "sessions do not run with parity taps");
    if (max_length < 2 || max_length > h->max_length) return fail(h, MOCR_ERR_CAPACITY, "max_length %d outside [2, %d]", max_length, h->max_length);
    if (rows < 1 || rows > h->max_batch) return fail(h, MOCR_ERR_CAPACITY, "%d decoder rows, handle capacity is %d", rows, h->max_batch);
    if (h->d_ring == nullptr) {
      TRY(dmalloc(h, &h->d_ring, static_cast<size_t>(h->max_batch)));
      CK(cudaMemsetAsync(h->d_ring, 0, sizeof(int) * h->max_batch, h->stream));
    }
    if (h->d_sess_map == nullptr) TRY(dmalloc(h, &h->d_sess_map, 64));
    if (h->ev_staged == nullptr) CK(cudaEventCreateWithFlags(&h->ev_staged, cudaEventDisableTiming));
    if (h->ev_published == nullptr) CK(cudaEventCreateWithFlags(&h->ev_published, cudaEventDisableTiming));
    h->sess_wait_pub = false;
    CK(cudaStreamSynchronize(h->stream));
    h->staged_ok = h->pre_ok = h->enc_ok = h->dec_ok = false;
    pd_publish_kernel<<<1, 1, 0, h->stream>>>(h->d_queue, 0, 0);                     // (the graph's warm-up step must find an empty queue)
    CK(cudaGetLastError());
    if (h->graphs.size() + 2 >= 64) {       // (the two programs of this session must not evict each other from the cache)
      for (auto& g : h->graphs) cudaGraphExecDestroy(g.second.exec);
      h->graphs.clear();
    }
    PdParams p = make_pd_params(h, rows, max_length, false, false, h->max_batch);
    p.queue_slots = h->d_ring;
    p.ext_queue = 1;
    p.idle_start = 1;
    TRY(decode_step_graph(h, p, false, false, &h->sess_exec, &h->sess_per_step));     // (its warm-up step runs on throw-away state)
    h->sess_exec_small = nullptr;
    if (rows > kSessSmallRows) {       // the program of a lightly loaded session (mocr_session_rows): same state, fewer rows per step
      PdParams ps = make_pd_params(h, kSessSmallRows, max_length, false, false, h->max_batch);
      ps.queue_slots = h->d_ring;
      ps.ext_queue = 1;
      ps.idle_start = 1;
      TRY(decode_step_graph(h, ps, false, false, &h->sess_exec_small, &h->sess_per_step_small));
      h->sess_p_small = ps;
    }
    h->sess_rows_now = rows;
    TRY(decode_begin(h, p));                                                         // every row idle, every id row [CLS] PAD ...
    pd_publish_kernel<<<1, 1, 0, h->stream>>>(h->d_queue, 0, 0);                     // nothing published, nothing finished
    CK(cudaGetLastError());
    ++h->launches;
    h->sess_p = p;
    h->sess_T = max_length;
    h->sess_rows = rows;
    h->sess_order = channel_order;
    h->sess_published = 0;
    h->sess_used.assign(static_cast<size_t>(h->max_batch), 0);
    h->sess_min_snap.assign(static_cast<size_t>(h->max_batch), 0);
    h->snap_enq = h->snap_read = 0;
    for (int i = 0; i < 2; ++i) {
      if (h->h_sess_lens[i] == nullptr) CK(cudaMallocHost(reinterpret_cast<void**>(&h->h_sess_lens[i]), sizeof(int) * h->max_batch));
      if (h->sess_ev[i] == nullptr) CK(cudaEventCreateWithFlags(&h->sess_ev[i], cudaEventDisableTiming));
    }
    h->sess_on = true;
    return MOCR_OK;
  });
}

int mocr_session_add(mocr_handle_t* h, const mocr_crop_t* crops, int n, int32_t* out_slots) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  return guarded(h, [&]() -> int {
    if (!h->sess_on) return fail(h, MOCR_ERR_INVALID, "no session is active");
    if (n < 1 || crops == nullptr || out_slots == nullptr) return fail(h, MOCR_ERR_INVALID, "bad argument");
    int n_free = 0;
    for (char u : h->sess_used) n_free += u ? 0 : 1;
    if (n > n_free) return fail(h, MOCR_ERR_CAPACITY, "%d crops added, %d slots are free", n, n_free);
    const bool was_empty = n_free == static_cast<int>(h->sess_used.size());
    {
      // all or nothing: every crop is checked before the first run of slots is published (a failure must not leave crops of an
      // earlier run decoding into slots the caller never learns about)
      int max_w = 0, tmp_rows = kPreStripRows;
      for (int i = 0; i < n; ++i) {
        const mocr_crop_t& c = crops[i];
        if (c.data == nullptr || c.height < 1 || c.width < 1 || (c.channels != 1 && c.channels != 3 && c.channels != 4) ||
            c.stride < c.width * c.channels)
          return fail(h, MOCR_ERR_INVALID, "crop %d is malformed (h=%d w=%d stride=%d channels=%d)", i, c.height, c.width, c.stride, c.channels);
        if (c.height > 32768 || c.width > 32768) return fail(h, MOCR_ERR_CAPACITY, "crop %d is larger than 32768 px", i);
        max_w = std::max(max_w, c.width);
        if (c.height != kImage) {
          mocr_handle::TableRef t;
          TRY(table_for(h, c.height, &t));
          tmp_rows = std::max(tmp_rows, t.strip_rows);
        }
      }
      const size_t smem = pre_smem_bytes(round_up(max_w, 16), tmp_rows);
      if (smem > 200 * 1024) return fail(h, MOCR_ERR_CAPACITY, "crop extents need %zu B of shared memory (limit 204800)", smem);
    }
    // Everything an admission launches - pixel upload, preprocess, the encoder, the publication - goes to the (lower-priority)
    // encoder stream: the decode steps already queued on the handle's stream keep running meanwhile, and the rows pick the new
    // crops up at their next next-token stage.  The encoder touches nothing the decoder reads except the K/V slots being added.
    struct Restore {
      mocr_handle* h;
      cudaStream_t s;
      ~Restore() { h->stream = s; h->sub_i0 = h->sub_n = 0; }
    } restore{h, h->stream};
    h->stream = h->sess_enc_hi ? h->stream_enc_hi : h->stream_enc;
    int done = 0;
    while (done < n) {
      // ONE staging + encoder pass for the crops (64 at a time): they take whichever slots are free - the pass works on its own
      // contiguous activations, only the cross-K/V epilogue and the publication go by slot (one pass per run of adjacent free slots
      // cost 1.3-1.6 passes of ~1 ms per admission, each waiting for the one before)
      const int len = std::min(n - done, 64);
      PdSlotList l{};
      l.n = len;
      for (int i = 0, s0 = 0; i < len; ++i, ++s0) {
        while (h->sess_used[s0]) ++s0;
        l.slot[i] = s0;
      }
      const auto t0 = std::chrono::steady_clock::now();
      const double w0 = h->prof_add[0];
      TRY(stage_crops(h, crops + done, len, h->sess_order, 0, nullptr, 0));
      const auto t1 = std::chrono::steady_clock::now();
      h->sub_i0 = 0;
      h->sub_n = len;
      TRY(preprocess(h));
      CK(cudaEventRecord(h->ev_staged, h->stream));
      pd_slot_map_kernel<<<1, 64, 0, h->stream>>>(h->d_sess_map, l);
      CK(cudaGetLastError());
      ++h->launches;
      h->enc_crop_map = h->d_sess_map;
      // (a pass always works on activations [0, len) and the same map: up to 4 crops - a lone caller's latency is mostly these 88
      //  launches - it is replayed from a graph per crop count; larger counts are launched directly, a capture costs ~7 ms once each)
      const int re = len <= 4 ? encode(h) : encode_launches(h);
      h->enc_crop_map = nullptr;
      TRY(re);
      h->prof_add[1] += std::chrono::duration<double>(t1 - t0).count() - (h->prof_add[0] - w0);
      h->prof_add[2] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
      h->prof_add[3] += 1;
      pd_publish_slots_kernel<<<1, 256, 0, h->stream>>>(h->d_queue, h->d_ring, h->max_batch, static_cast<int>(h->sess_published), h->d_ids,
                                                        h->d_lens, h->sess_T, l);
      CK(cudaGetLastError());
      ++h->launches;
      h->sess_published += len;
      CK(cudaEventRecord(h->ev_published, h->stream));
      for (int i = 0; i < len; ++i) {
        h->sess_used[l.slot[i]] = 1;               // (which snapshots may speak for the slot was settled when it was released)
        out_slots[done + i] = l.slot[i];
      }
      done += len;
    }
    // An empty session has no row to hold up: its next chunk starts when these crops are in the queue instead of stepping idle rows
    // while the encoder pass runs (a lone caller's text then has the whole chunk to end in)
    h->sess_wait_pub = was_empty;
    return MOCR_OK;
  });
}

int mocr_session_run(mocr_handle_t* h, int steps, int32_t* out_lens) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  return guarded(h, [&]() -> int {
    if (!h->sess_on) return fail(h, MOCR_ERR_INVALID, "no session is active");
    if (steps < 0 || (steps == 0 && out_lens == nullptr)) return fail(h, MOCR_ERR_INVALID, "bad argument");
    if (steps > 0 && h->sess_wait_pub) {
      CK(cudaStreamWaitEvent(h->stream, h->ev_published, 0));
      h->sess_wait_pub = false;
    }
    const bool small = h->sess_rows_now < h->sess_rows;
    const cudaGraphExec_t exec = small ? h->sess_exec_small : h->sess_exec;
    const int spg = exec != nullptr ? std::max(1, std::min(h->steps_per_graph, h->sess_T - 1)) : 1;
    for (int ran = 0; ran < steps; ran += spg) {
      if (exec != nullptr) {
        CK(cudaGraphLaunch(exec, h->stream));
        h->launches += small ? h->sess_per_step_small : h->sess_per_step;
      } else {
        TRY(decode_stage_step(h, small ? h->sess_p_small : h->sess_p));
      }
    }
    // A snapshot of the slots' lengths follows every chunk of steps; up to two are in flight, so the host can read the snapshot of
    // chunk k (an event wait, not a stream sync) while chunk k + 1 is already queued: the GPU never waits for the host.
    auto enqueue_snapshot = [&]() -> int {
      if (h->snap_enq - h->snap_read >= 2) return fail(h, MOCR_ERR_INVALID, "two length snapshots are pending: read one first");
      const int k = static_cast<int>(h->snap_enq & 1);
      CK(cudaMemcpyAsync(h->h_sess_lens[k], h->d_lens, sizeof(int) * h->max_batch, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaEventRecord(h->sess_ev[k], h->stream));
      ++h->snap_enq;
      return MOCR_OK;
    };
    if (steps > 0) TRY(enqueue_snapshot());
    if (out_lens == nullptr) return MOCR_OK;       // launch only: the caller admits crops meanwhile and reads later (steps = 0)
    if (h->snap_enq == h->snap_read) TRY(enqueue_snapshot());
    const long long j = h->snap_read;              // the oldest unread snapshot
    CK(cudaEventSynchronize(h->sess_ev[j & 1]));
    const int* lens = h->h_sess_lens[j & 1];
    for (int i = 0; i < h->max_batch; ++i) out_lens[i] = (h->sess_used[i] && j >= h->sess_min_snap[i]) ? lens[i] : 0;
    ++h->snap_read;
    return MOCR_OK;
  });
}

int mocr_session_rows(mocr_handle_t* h, int rows) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->sess_on) return fail(h, MOCR_ERR_INVALID, "no session is active");
  if (rows < 1 || rows > h->sess_rows) return fail(h, MOCR_ERR_CAPACITY, "%d rows, the session has %d", rows, h->sess_rows);
  const int want = (rows <= kSessSmallRows && h->sess_rows > kSessSmallRows) ? kSessSmallRows : h->sess_rows;
  if (want < h->sess_rows_now)       // rows above the new count must be idle: certain only when no crop is in the session
    for (char u : h->sess_used)
      if (u) return fail(h, MOCR_ERR_INVALID, "the row count of a session shrinks only while no slot is in use");
  h->sess_rows_now = want;
  return want;
}

int mocr_session_fetch(mocr_handle_t* h, const int32_t* slots, int n, int32_t* out_ids, int release) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  return guarded(h, [&]() -> int {
    if (!h->sess_on) return fail(h, MOCR_ERR_INVALID, "no session is active");
    if (n < 0 || (n > 0 && (slots == nullptr || out_ids == nullptr))) return fail(h, MOCR_ERR_INVALID, "bad argument");
    // The rows of finished slots are final: they are copied on a side stream, without waiting for the decode steps that are queued
    // on the handle's stream.  A released slot's length is zeroed here, synchronously, before the slot can be handed out again
    // (its next occupant is published from the encoder stream; snapshots enqueued before that are ignored for it).
    cudaStream_t side = h->stream_fetch;
    for (int i = 0; i < n; ++i) {
      if (slots[i] < 0 || slots[i] >= h->max_batch || !h->sess_used[slots[i]]) return fail(h, MOCR_ERR_INVALID, "slot %d is not in use", slots[i]);
      CK(cudaMemcpyAsync(out_ids + static_cast<size_t>(i) * h->sess_T, h->d_ids + static_cast<size_t>(slots[i]) * h->sess_T, sizeof(int) * h->sess_T,
                         cudaMemcpyDeviceToHost, side));
      if (release) CK(cudaMemsetAsync(h->d_lens + slots[i], 0, sizeof(int), side));
    }
    CK(cudaStreamSynchronize(side));
    if (release)
      for (int i = 0; i < n; ++i) {
        h->sess_used[slots[i]] = 0;
        // The length was zeroed synchronously just now: a snapshot enqueued from here on is taken after that and shows 0 or the next
        // occupant's length; one enqueued earlier may already have been taken with this occupant's length in it and must not speak
        // for the next one.  (Counting from the next ADMISSION instead cost short texts a whole chunk: the snapshot that follows
        // the chunk they are decoded in is usually enqueued before they are admitted.)
        h->sess_min_snap[slots[i]] = h->snap_enq;
      }
    return MOCR_OK;
  });
}

int mocr_session_end(mocr_handle_t* h) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->sess_on) return fail(h, MOCR_ERR_INVALID, "no session is active");
  return session_end(h);
}

int decode_beam_device(mocr_handle* h, int beams, int max_length, int ngram, float length_penalty, int early, int32_t* out_ids, int32_t* out_lens,
                       float* out_scores);

// Beam search over the encoded crops (include/mocr_b200.h: mocr_decode_beam).
int decode_beam(mocr_handle* h, int beams, int max_length, int ngram, float length_penalty, int early, int32_t* out_ids, int32_t* out_lens,
                float* out_scores) {
  if (!h->enc_ok) return fail(h, MOCR_ERR_INVALID, "encode has not run on the staged crops");
  if (beams < 1 || 2 * beams > kBeamMaxK || ngram < 0 || early < 0 || early > 2 || out_ids == nullptr)
    return fail(h, MOCR_ERR_INVALID, "bad beam-search argument");
  if (max_length < 2 || max_length > h->max_length) return fail(h, MOCR_ERR_CAPACITY, "max_length %d outside [2, %d]", max_length, h->max_length);
  const int n = h->n, R = n * beams, K = 2 * beams;
  if (R > h->max_batch) return fail(h, MOCR_ERR_CAPACITY, "%d crops x %d beams = %d rows, handle capacity is %d", n, beams, R, h->max_batch);
  if (h->beam_device) {
    const size_t need_tap_dev = static_cast<size_t>(h->max_batch) * kVocab * sizeof(float);
    if (need_tap_dev > h->logits_tap_bytes) {
      CK(cudaStreamSynchronize(h->stream));
      if (h->logits_tap) cudaFree(h->logits_tap);
      h->logits_tap = nullptr;
      h->logits_tap_bytes = 0;
      CK(cudaMalloc(reinterpret_cast<void**>(&h->logits_tap), need_tap_dev));
      h->logits_tap_bytes = need_tap_dev;
      for (auto& g : h->graphs) cudaGraphExecDestroy(g.second.exec);   // they captured the old tap pointer
      h->graphs.clear();
      for (auto& g : h->beam_graphs) cudaGraphExecDestroy(g.second.exec);
      h->beam_graphs.clear();
    }
    return decode_beam_device(h, beams, max_length, ngram, length_penalty, early, out_ids, out_lens, out_scores);
  }
  const size_t cache_elems = static_cast<size_t>(h->max_batch) * h->max_length * kD;
  for (int l = 0; l < kDecLayers; ++l) {
    if (h->self_k2[l] == nullptr) TRY(dmalloc(h, &h->self_k2[l], cache_elems));
    if (h->self_v2[l] == nullptr) TRY(dmalloc(h, &h->self_v2[l], cache_elems));
  }
  const size_t need_tap = static_cast<size_t>(h->max_batch) * kVocab * sizeof(float);
  if (need_tap > h->logits_tap_bytes) {
    CK(cudaStreamSynchronize(h->stream));
    if (h->logits_tap) cudaFree(h->logits_tap);
    h->logits_tap = nullptr;
    h->logits_tap_bytes = 0;
    CK(cudaMalloc(reinterpret_cast<void**>(&h->logits_tap), need_tap));
    h->logits_tap_bytes = need_tap;
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.second.exec);   // they captured the old tap pointer
    h->graphs.clear();
  }
  // scratch: cand_lp [R][K] f32 | cand_tok [R][K] | next [R] | parent [R] | ban_cnt [R] | ban_tok [R][max_length]
  const size_t Rm = h->max_batch;
  const size_t off_tok = Rm * kBeamMaxK * 4, off_next = off_tok + Rm * kBeamMaxK * 4, off_parent = off_next + Rm * 4, off_bcnt = off_parent + Rm * 4,
               off_btok = off_bcnt + Rm * 4, total = off_btok + Rm * h->max_length * 4;
  if (total > h->d_beam_bytes) {
    if (h->d_beam) cudaFree(h->d_beam);
    h->d_beam = nullptr;
    h->d_beam_bytes = 0;
    CK(cudaMalloc(&h->d_beam, total));
    h->d_beam_bytes = total;
  }
  uint8_t* sb = static_cast<uint8_t*>(h->d_beam);
  float* d_lp = reinterpret_cast<float*>(sb);
  int* d_tok = reinterpret_cast<int*>(sb + off_tok);
  int* d_next = reinterpret_cast<int*>(sb + off_next);
  int* d_parent = reinterpret_cast<int*>(sb + off_parent);
  int* d_bcnt = reinterpret_cast<int*>(sb + off_bcnt);
  int* d_btok = reinterpret_cast<int*>(sb + off_btok);

  BeamSearch bs;
  bs.init(n, beams, max_length, ngram, length_penalty, early, kClsId);
  std::vector<float> lp(static_cast<size_t>(R) * K);
  std::vector<int32_t> tok(static_cast<size_t>(R) * K), next(R), parent(R), bcnt(R, 0), btok;
  __nv_bfloat16 *ck[kDecLayers], *cv[kDecLayers], *ak[kDecLayers], *av[kDecLayers];    // current / alternate caches
  for (int l = 0; l < kDecLayers; ++l) { ck[l] = h->self_k[l]; cv[l] = h->self_v[l]; ak[l] = h->self_k2[l]; av[l] = h->self_v2[l]; }
  PdParams p = make_pd_params(h, R, max_length, false, true);
  p.kv_div = beams;
  p.logits_cur = 1;
  CK(launch_pdl(h, pd_begin_kernel, (R + kPdWarps - 1) / kPdWarps, kPdThreads, 0, p));
  ++h->launches;
  CK(cudaMemsetAsync(d_bcnt, 0, sizeof(int) * R, h->stream));
  int ban_cap = 1, steps = 0;
  for (int t = 0; t < max_length - 1 && bs.unfinished; ++t) {
    for (int l = 0; l < kDecLayers; ++l) { p.layer[l].self_k = ck[l]; p.layer[l].self_v = cv[l]; }
    TRY(decode_stage_step(h, p, true));
    beam_topk_kernel<<<R, 256, 0, h->stream>>>(h->logits_tap, d_btok, d_bcnt, ban_cap, K, d_lp, d_tok);
    CK(cudaGetLastError());
    ++h->launches;
    CK(cudaMemcpyAsync(lp.data(), d_lp, sizeof(float) * R * K, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(tok.data(), d_tok, sizeof(int) * R * K, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    ++steps;
    if (!bs.step(lp.data(), tok.data(), next.data(), parent.data())) break;
    // ---- next step's inputs: tokens, cache rows, n-gram bans
    bool moved = false;
    for (int r = 0; r < R; ++r) moved = moved || parent[r] != r;
    CK(cudaMemcpyAsync(d_next, next.data(), sizeof(int) * R, cudaMemcpyHostToDevice, h->stream));
    if (moved) {
      CK(cudaMemcpyAsync(d_parent, parent.data(), sizeof(int) * R, cudaMemcpyHostToDevice, h->stream));
      BeamCaches bc;
      for (int l = 0; l < kDecLayers; ++l) {
        bc.src[2 * l] = ck[l]; bc.dst[2 * l] = ak[l];
        bc.src[2 * l + 1] = cv[l]; bc.dst[2 * l + 1] = av[l];
      }
      beam_kv_gather_kernel<<<dim3(R, 2 * kDecLayers), 256, 0, h->stream>>>(bc, d_parent, t + 1, h->max_length);
      CK(cudaGetLastError());
      ++h->launches;
      for (int l = 0; l < kDecLayers; ++l) { std::swap(ck[l], ak[l]); std::swap(cv[l], av[l]); }
    }
    int mx = 0;
    if (ngram > 0) {
      btok.assign(static_cast<size_t>(R) * max_length, 0);
      for (int r = 0; r < R; ++r) {
        bcnt[r] = bs.banned(r, &btok[static_cast<size_t>(r) * max_length], max_length);
        mx = std::max(mx, bcnt[r]);
      }
    }
    ban_cap = std::max(mx, 1);
    if (mx > 0) {     // pack [R][ban_cap]
      std::vector<int32_t> packed(static_cast<size_t>(R) * ban_cap, 0);
      for (int r = 0; r < R; ++r) std::copy_n(&btok[static_cast<size_t>(r) * max_length], std::min(bcnt[r], ban_cap), &packed[static_cast<size_t>(r) * ban_cap]);
      CK(cudaMemcpyAsync(d_btok, packed.data(), sizeof(int) * packed.size(), cudaMemcpyHostToDevice, h->stream));
      CK(cudaMemcpyAsync(d_bcnt, bcnt.data(), sizeof(int) * R, cudaMemcpyHostToDevice, h->stream));
      CK(cudaStreamSynchronize(h->stream));      // (pageable sources)
    } else {
      CK(cudaMemsetAsync(d_bcnt, 0, sizeof(int) * R, h->stream));
    }
    beam_advance_kernel<<<(R + 7) / 8, 256, 0, h->stream>>>(p, d_next, t + 1);
    CK(cudaGetLastError());
    ++h->launches;
  }
  CK(cudaStreamSynchronize(h->stream));
  bs.result(out_ids, out_lens, out_scores);
  h->last_steps = steps;
  h->cur_len = max_length;
  h->dec_ok = false;        // the greedy-path getters (ids, step logits) do not describe this run
  return MOCR_OK;
}

// Beam search with the selection on the device (beam_device.cuh): a step = the greedy path's stage kernels (all but the
// next-token stage) + top-k with the n-gram ban + selection + cache-row fix-up + embed, captured beam_steps_per_graph steps
// at a time; the host only reads the 32-byte control block after every graph.
int decode_beam_device(mocr_handle* h, int beams, int max_length, int ngram, float length_penalty, int early, int32_t* out_ids, int32_t* out_lens,
                       float* out_scores) {
  const int n = h->n, R = n * beams, K = 2 * beams, T = max_length;
  const size_t Rm = h->max_batch, Tm = h->max_length;
  // scratch layout (ints / floats of 4 bytes)
  size_t off = 0;
  auto take = [&](size_t count) { const size_t o = off; off += (count * 4 + 255) & ~static_cast<size_t>(255); return o; };
  const size_t o_run0 = take(Rm * Tm), o_run1 = take(Rm * Tm), o_fin0 = take(Rm * Tm), o_fin1 = take(Rm * Tm), o_rs = take(Rm), o_fs = take(Rm),
               o_fl = take(Rm), o_if = take(Rm), o_un = take(Rm), o_ctl = take(8), o_clp = take(Rm * kBeamMaxK), o_ctk = take(Rm * kBeamMaxK),
               o_nx = take(Rm), o_pa = take(Rm), o_ph = take(Rm), o_cs = take(Rm), o_cd = take(Rm);
  if (off > h->d_beam_dev_bytes) {
    CK(cudaStreamSynchronize(h->stream));
    if (h->d_beam_dev) cudaFree(h->d_beam_dev);
    h->d_beam_dev = nullptr;
    h->d_beam_dev_bytes = 0;
    CK(cudaMalloc(&h->d_beam_dev, off));
    h->d_beam_dev_bytes = off;
    for (auto& g : h->beam_graphs) cudaGraphExecDestroy(g.second.exec);
    h->beam_graphs.clear();
  }
  if (h->h_beam_ctl == nullptr) CK(cudaMallocHost(reinterpret_cast<void**>(&h->h_beam_ctl), 8 * sizeof(int)));
  uint8_t* base = static_cast<uint8_t*>(h->d_beam_dev);
  BeamDev d{};
  d.n = n; d.beams = beams; d.K = K; d.T = T; d.ngram = ngram; d.early = early; d.length_penalty = length_penalty;
  d.eos = kSepId; d.fill = kSepId;      // pad id 0 is falsy in the reference: finished rows are filled with EOS (beam_search.h)
  d.prompt_len = 1; d.start_token = kClsId;
  d.run[0] = reinterpret_cast<int*>(base + o_run0); d.run[1] = reinterpret_cast<int*>(base + o_run1);
  d.fin[0] = reinterpret_cast<int*>(base + o_fin0); d.fin[1] = reinterpret_cast<int*>(base + o_fin1);
  d.run_score = reinterpret_cast<float*>(base + o_rs); d.fin_score = reinterpret_cast<float*>(base + o_fs);
  d.fin_len = reinterpret_cast<int*>(base + o_fl); d.is_fin = reinterpret_cast<int*>(base + o_if); d.unsat = reinterpret_cast<int*>(base + o_un);
  d.ctl = reinterpret_cast<int*>(base + o_ctl);
  d.cand_lp = reinterpret_cast<float*>(base + o_clp); d.cand_tok = reinterpret_cast<int*>(base + o_ctk);
  d.next = reinterpret_cast<int*>(base + o_nx); d.parent = reinterpret_cast<int*>(base + o_pa); d.phys = reinterpret_cast<int*>(base + o_ph);
  d.copy_src = reinterpret_cast<int*>(base + o_cs); d.copy_dst = reinterpret_cast<int*>(base + o_cd);

  PdParams p = make_pd_params(h, R, max_length, false, true);
  p.kv_div = beams;
  p.logits_cur = 1;
  p.kv_row = d.phys;
  BeamCaches bc;
  for (int l = 0; l < kDecLayers; ++l) {
    bc.src[2 * l] = bc.dst[2 * l] = h->self_k[l];
    bc.src[2 * l + 1] = bc.dst[2 * l + 1] = h->self_v[l];
  }
  auto one_step = [&]() -> int {
    TRY(decode_stage_step(h, p, true));
    // (programmatic dependent launch like the stage kernels: each waits for its predecessor with griddepcontrol.wait)
    CK(launch_pdl(h, beam_topk_dev_kernel, R, 256, 0, d, static_cast<const float*>(h->logits_tap)));
    CK(launch_pdl(h, beam_select_kernel, n, 128, 0, p, d));
    CK(launch_pdl_grid(h, beam_kv_copy_kernel, dim3(R, 2 * kDecLayers, kBeamCopySplit), 256, d, bc, h->max_length));
    h->launches += 3;
    return MOCR_OK;
  };
  auto begin = [&]() -> int {
    CK(launch_pdl(h, pd_begin_kernel, (R + kPdWarps - 1) / kPdWarps, kPdThreads, 0, p));
    beam_init_kernel<<<std::min(2 * h->sms, (R * T + 255) / 256), 256, 0, h->stream>>>(d);
    CK(cudaGetLastError());
    h->launches += 2;
    return MOCR_OK;
  };
  TRY(begin());
  const int spg = std::max(1, std::min(h->beam_steps_per_graph, max_length - 1));
  cudaGraphExec_t exec = nullptr;
  int64_t per_graph = 0;
  if (h->use_graph) {
    uint32_t lp_bits;
    memcpy(&lp_bits, &length_penalty, 4);
    const uint64_t key = (static_cast<uint64_t>(n) << 52) ^ (static_cast<uint64_t>(beams) << 44) ^ (static_cast<uint64_t>(max_length) << 34) ^
                         (static_cast<uint64_t>(ngram) << 30) ^ (static_cast<uint64_t>(early) << 28) ^ (static_cast<uint64_t>(lp_bits) >> 4);
    auto it = h->beam_graphs.find(key);
    if (it == h->beam_graphs.end()) {
      const int64_t l0 = h->launches;
      TRY(one_step());                        // warm: function attributes, tensor maps
      per_graph = (h->launches - l0) * spg;
      TRY(begin());
      cudaGraph_t graph;
      CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
      const int64_t l1 = h->launches;
      int r = MOCR_OK;
      for (int s = 0; s < spg && r == MOCR_OK; ++s) r = one_step();
      h->launches = l1;
      const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
      if (r != MOCR_OK) {
        if (ce == cudaSuccess) cudaGraphDestroy(graph);
        return r;
      }
      if (ce != cudaSuccess) return fail(h, MOCR_ERR_CUDA, "beam-search stream capture failed: %s", cudaGetErrorString(ce));
      CK(cudaGraphInstantiate(&exec, graph, 0));
      cudaGraphDestroy(graph);
      if (h->beam_graphs.size() >= 16) {
        for (auto& g : h->beam_graphs) cudaGraphExecDestroy(g.second.exec);
        h->beam_graphs.clear();
      }
      h->beam_graphs[key] = mocr_handle::StepGraph{exec, static_cast<int>(per_graph)};
    } else {
      exec = it->second.exec;
      per_graph = it->second.launches;
    }
  }
  int launched = 0;
  while (launched < max_length - 1) {
    if (exec != nullptr) {
      CK(cudaGraphLaunch(exec, h->stream));
      h->launches += per_graph;
      launched += spg;
    } else {
      TRY(one_step());
      launched += 1;
    }
    CK(cudaMemcpyAsync(h->h_beam_ctl, d.ctl, 8 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->h_beam_ctl[BC_UNFINISHED] == 0) break;
  }
  if (h->h_beam_ctl[BC_UNFINISHED] != 0) return fail(h, MOCR_ERR_CUDA, "beam search did not terminate within max_length steps");
  const int par = h->h_beam_ctl[BC_PARITY];
  std::vector<int> fl(R);
  std::vector<float> fs(R);
  CK(cudaMemcpyAsync(fl.data(), d.fin_len, sizeof(int) * R, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(fs.data(), d.fin_score, sizeof(float) * R, cudaMemcpyDeviceToHost, h->stream));
  // best hypothesis of every crop = row 0 of its finished set
  CK(cudaMemcpy2DAsync(out_ids, sizeof(int) * T, d.fin[par], sizeof(int) * T * beams, sizeof(int) * T, n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int b = 0; b < n; ++b) {
    if (out_lens) out_lens[b] = 1 + fl[static_cast<size_t>(b) * beams];
    if (out_scores) out_scores[b] = fs[static_cast<size_t>(b) * beams];
  }
  h->last_steps = h->h_beam_ctl[BC_STEPS];
  h->last_rows = R;
  h->cur_len = max_length;
  h->dec_ok = false;
  return MOCR_OK;
}

int mocr_decode_beam(mocr_handle_t* h, int num_beams, int max_length, int no_repeat_ngram_size, float length_penalty, int early_stopping,
                     int32_t* out_ids, int32_t* out_lens, float* out_scores) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  return guarded(h, [&]() -> int { return decode_beam(h, num_beams, max_length, no_repeat_ngram_size, length_penalty, early_stopping, out_ids, out_lens, out_scores); });
}

// Crops / selections -> beam-search hypotheses in ONE call: the handle's mutex is held from staging to the result, so
// concurrent callers of the same handle cannot interleave between stage, encode and decode (chunks of
// max_batch / num_beams crops).  page == NULL: `crops` are host crops; page != NULL: `regions` of that page.
int recognize_beam_impl(mocr_handle* h, const mocr_crop_t* crops, const mocr_crop_t* page, const mocr_region_t* regions, int n, int order,
                        int max_length, int beams, int ngram, float length_penalty, int early, int32_t* out_ids, int32_t* out_lens,
                        float* out_scores) {
  if (n < 0 || (n > 0 && out_ids == nullptr) || (n > 0 && page == nullptr && crops == nullptr) || (n > 0 && page != nullptr && regions == nullptr))
    return fail(h, MOCR_ERR_INVALID, "bad argument");
  if (beams < 1 || beams > h->max_batch) return fail(h, MOCR_ERR_CAPACITY, "num_beams %d outside [1, max_batch = %d]", beams, h->max_batch);
  const int per = std::max(1, h->max_batch / beams);
  for (int i0 = 0; i0 < n; i0 += per) {
    const int m = std::min(per, n - i0);
    if (page != nullptr) TRY(stage_regions(h, page, regions + i0, m, order));
    else TRY(stage_crops(h, crops + i0, m, order));
    TRY(preprocess(h));
    TRY(encode(h));
    TRY(decode_beam(h, beams, max_length, ngram, length_penalty, early, out_ids + static_cast<size_t>(i0) * max_length,
                    out_lens ? out_lens + i0 : nullptr, out_scores ? out_scores + i0 : nullptr));
  }
  return MOCR_OK;
}

int mocr_recognize_beam(mocr_handle_t* h, const mocr_crop_t* crops, int n, int channel_order, int max_length, int num_beams,
                        int no_repeat_ngram_size, float length_penalty, int early_stopping, int32_t* out_ids, int32_t* out_lens,
                        float* out_scores) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  return guarded(h, [&]() -> int { return recognize_beam_impl(h, crops, nullptr, nullptr, n, channel_order, max_length, num_beams, no_repeat_ngram_size, length_penalty,
                             early_stopping, out_ids, out_lens, out_scores); });
}

int mocr_recognize_regions_beam(mocr_handle_t* h, const mocr_crop_t* page, const mocr_region_t* regions, int n, int channel_order,
                                int max_length, int num_beams, int no_repeat_ngram_size, float length_penalty, int early_stopping,
                                int32_t* out_ids, int32_t* out_lens, float* out_scores) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (page == nullptr) return fail(h, MOCR_ERR_INVALID, "page is NULL");
  return recognize_beam_impl(h, nullptr, page, regions, n, channel_order, max_length, num_beams, no_repeat_ngram_size, length_penalty,
                             early_stopping, out_ids, out_lens, out_scores);
}

// ---- beam-search bookkeeping (host only; usable without a device: the CPU tests drive it with the oracle's logits)
struct mocr_beam {
  BeamSearch s;
};

int mocr_beam_create(int n, int num_beams, int max_length, int no_repeat_ngram, float length_penalty, int early_stopping, mocr_beam_t** out) {
  if (out == nullptr || n < 1 || num_beams < 1 || num_beams > 16 || max_length < 2 || max_length > kMaxPos || no_repeat_ngram < 0 ||
      early_stopping < 0 || early_stopping > 2)
    return MOCR_ERR_INVALID;
  mocr_beam* b = new (std::nothrow) mocr_beam();
  if (b == nullptr) return MOCR_ERR_CAPACITY;
  b->s.init(n, num_beams, max_length, no_repeat_ngram, length_penalty, early_stopping, 2 /*[CLS]*/);
  *out = b;
  return MOCR_OK;
}
int mocr_beam_destroy(mocr_beam_t* b) {
  delete b;
  return MOCR_OK;
}
int mocr_beam_banned(const mocr_beam_t* b, int row, int32_t* out, int cap) {
  if (b == nullptr || row < 0 || row >= b->s.n * b->s.beams || (cap > 0 && out == nullptr) || cap < 0) return MOCR_ERR_INVALID;
  return b->s.banned(row, out, cap);
}
int mocr_beam_step(mocr_beam_t* b, const float* cand_logprob, const int32_t* cand_token, int32_t* next_tokens, int32_t* parents) {
  if (b == nullptr || cand_logprob == nullptr || cand_token == nullptr || next_tokens == nullptr || parents == nullptr) return MOCR_ERR_INVALID;
  if (!b->s.unfinished) return 0;
  return b->s.step(cand_logprob, cand_token, next_tokens, parents) ? 1 : 0;
}
int mocr_beam_result(const mocr_beam_t* b, int32_t* ids, int32_t* lens, float* scores) {
  if (b == nullptr || ids == nullptr) return MOCR_ERR_INVALID;
  b->s.result(ids, lens, scores);
  return MOCR_OK;
}

int mocr_set_taps(mocr_handle_t* h, int taps) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  h->taps = taps;
  return ensure_taps(h);
}

int mocr_get_pixels_u8(mocr_handle_t* h, uint8_t* out) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->pre_ok || h->px_u8 == nullptr || out == nullptr) return fail(h, MOCR_ERR_INVALID, "pixel tap not available");
  return d2h(h, out, h->px_u8, static_cast<size_t>(h->n) * kImage * kImage);
}

int mocr_get_pixel_values(mocr_handle_t* h, float* out) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->pre_ok || h->px_f32 == nullptr || out == nullptr) return fail(h, MOCR_ERR_INVALID, "pixel tap not available");
  return d2h(h, out, h->px_f32, static_cast<size_t>(h->n) * kImage * kImage * sizeof(float));
}

int mocr_get_encoder_hidden(mocr_handle_t* h, float* out) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->enc_ok || h->enc_f32 == nullptr || out == nullptr) return fail(h, MOCR_ERR_INVALID, "encoder tap not available");
  return d2h(h, out, h->enc_f32, static_cast<size_t>(h->n) * kEncTokens * kD * sizeof(float));
}

int mocr_get_step_logits(mocr_handle_t* h, float* out) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->dec_ok || h->logits_tap == nullptr || out == nullptr) return fail(h, MOCR_ERR_INVALID, "logits tap not available");
  return d2h(h, out, h->logits_tap, static_cast<size_t>(h->n) * (h->cur_len - 1) * kVocab * sizeof(float));
}

void* mocr_stream(mocr_handle_t* h) { return h ? static_cast<void*>(h->stream) : nullptr; }

int mocr_sync(mocr_handle_t* h) {
  TRY(check_handle(h));
  CK(cudaStreamSynchronize(h->stream));
  return MOCR_OK;
}

int64_t mocr_launch_count(mocr_handle_t* h) { return h ? h->launches : 0; }
int mocr_last_steps(mocr_handle_t* h) {
  if (h == nullptr) return 0;
  return h->last_steps;
}

int mocr_set_option(mocr_handle_t* h, const char* key, int value) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (key == nullptr) return fail(h, MOCR_ERR_INVALID, "key is NULL");
  if (h->sess_on) return fail(h, MOCR_ERR_INVALID, "options cannot change while a session is active (its CUDA graph is in use)");
  const std::string k = key;
  auto bn_ok = [](int v) { return v == 32 || v == 64 || v == 128 || v == 192 || v == 256; };
  if (k == "enc_bn" && bn_ok(value) && kD % value == 0) h->enc_bn = value;
  else if (k == "enc_bn768" && (value == 128 || value == 192 || value == 256)) h->enc_bn768 = value;
  else if (k == "check_every" && value >= 1) h->check_every = value;
  else if (k == "use_graph") h->use_graph = value != 0;
  else if (k == "use_pdl") h->use_pdl = value != 0;
  else if (k == "pdl_mask" && value >= 0) h->pdl_mask = value;
  else if (k == "resid_tma") h->resid_tma = value != 0;
  else if (k == "row_warps" && value >= 1 && value <= 8) h->row_warps = value;
  else if (k == "big_row_warps" && value >= 1 && value <= 8) h->big_row_warps = value;
  else if (k == "dec_tc" && value >= 0 && value <= 7) h->dec_tc = value;
  else if (k == "carveout" && value >= -1 && value <= 100) h->carveout = value;
  else if (k == "kv_evict_first" && value >= -1 && value <= 3) h->kv_evict_first = value;
  else if (k == "attn_grid" && value >= 0) h->attn_grid = value;
  else if (k == "slots" && value >= 0) h->slots = value;
  else if (k == "stage_threads" && value >= 1 && value <= 64) h->stage_threads = value;
  else if (k == "beam_device") h->beam_device = value != 0;
  else if (k == "beam_steps_per_graph" && value >= 1 && value <= 64) h->beam_steps_per_graph = value;
  else if (k == "pipeline" && value >= 0 && value <= 2) h->pipeline = value;
  else if (k == "fuse_ln") h->fuse_ln = value != 0;
  else if (k == "kv_prefetch") h->kv_prefetch = value != 0;
  else if (k == "big_rows" && value >= 1) h->big_rows = value;
  else if (k == "big_attn_grid" && value >= 0) h->big_attn_grid = value;
  else if (k == "big_attn_rows") h->big_attn_rows = value != 0;
  else if (k == "enc_tma_store") h->enc_tma_store = value != 0;
  else if (k == "stage_chunk" && value >= 0) h->stage_chunk = value;
  else if (k == "big_accum") h->big_accum = value != 0;
  else if (k == "big_ksplit") h->big_ksplit = value != 0;
#ifdef MOCR_GEMM_DBG
  else if (k == "gemm_dbg") h->gemm_dbg = value;
#endif
  else if (k == "big_bn768" && (value == 32 || value == 64)) h->big_bn768 = value;
  else if (k == "big_vocab_bn" && (value == 0 || value == 64 || value == 128 || value == 256)) h->big_vocab_bn = value;
  else if (k == "sess_enc_hi") h->sess_enc_hi = value != 0;
  else if (k == "steps_per_graph" && value >= 1 && value <= 64) h->steps_per_graph = value;
  else if (k == "decode_prof") h->decode_prof = value != 0;
  else return fail(h, MOCR_ERR_INVALID, "unknown option or bad value: %s=%d", key, value);
  for (auto& g : h->graphs) cudaGraphExecDestroy(g.second.exec);
  h->graphs.clear();
  for (auto& g : h->enc_graphs) cudaGraphExecDestroy(g.second.exec);
  h->enc_graphs.clear();
  for (auto& g : h->beam_graphs) cudaGraphExecDestroy(g.second.exec);
  h->beam_graphs.clear();
  return MOCR_OK;
}

int mocr_time_kernel(mocr_handle_t* h, const char* kernel, int iters, float* ms_per_launch, double* algo_bytes, double* algo_flops) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (kernel == nullptr || iters < 1 || ms_per_launch == nullptr) return fail(h, MOCR_ERR_INVALID, "bad argument");
  if (!h->enc_ok) return fail(h, MOCR_ERR_INVALID, "run a batch through mocr_encode first");
  const std::string k = kernel;
  const int n = h->n, M = n * kEncTokens;
  double bytes = 0, flops = 0;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int r = MOCR_OK;
  for (int i = -2; i < iters && r == MOCR_OK; ++i) {
    if (i == 0) CK(cudaEventRecord(e0, h->stream));
    if (k == "enc_fc1") {
      r = gemm(h, EPI_BF16_GELU, h->enc_bn, h->xn, h->enc[0].fc1, M, out_bf16(h->mlp.p, kFFN));
      flops = 2.0 * M * kFFN * kD;
      bytes = 2.0 * (static_cast<double>(M) * kD + static_cast<double>(kFFN) * kD + static_cast<double>(M) * kFFN);
    } else if (k == "enc_fc2") {
      // (as in the encoder: accumulates into the f32 residual stream, which nothing reads after the encode)
      r = gemm(h, h->resid_tma ? EPI_F32_ACCUM : EPI_F32_RESID, h->enc_bn768, h->mlp, h->enc[0].fc2, M, out_f32(h->hres, kD, h->hres, kD));
      flops = 2.0 * M * kFFN * kD;
      bytes = 2.0 * (static_cast<double>(M) * kFFN + static_cast<double>(kFFN) * kD) + 8.0 * M * kD;
    } else if (k == "enc_out") {
      r = gemm(h, h->resid_tma ? EPI_F32_ACCUM : EPI_F32_RESID, h->enc_bn768, h->ctx, h->enc[0].out, M, out_f32(h->hres, kD, h->hres, kD));
      flops = 2.0 * M * kD * kD;
      bytes = 2.0 * (static_cast<double>(M) * kD + static_cast<double>(kD) * kD) + 8.0 * M * kD;
    } else if (k == "enc_ln") {
      r = layernorm(h, h->hres, M, h->enc[0].ln1, h->xn.p, nullptr);
      bytes = 6.0 * M * kD;
    } else if (k == "enc_qkv") {
      r = gemm(h, EPI_BF16, h->enc_bn, h->xn, h->enc[0].qkv, M, out_bf16(h->qkv, 3 * kD));
      flops = 2.0 * M * 3 * kD * kD;
      bytes = 2.0 * (static_cast<double>(M) * kD + 3.0 * kD * kD + static_cast<double>(M) * 3 * kD);
    } else if (k == "enc_attn") {
      r = attention197(h, n);
      flops = 4.0 * n * kHeads * kEncTokens * kEncTokens * kHeadDim;
      bytes = 2.0 * (static_cast<double>(M) * 3 * kD + static_cast<double>(M) * kD);
    } else if (k.rfind("dec_", 0) == 0) {
      // one stage kernel of the decoder's per-token program, on the state the last decode left
      // (finished flags cleared through a forced-decoding view so that every row is processed)
      if (!h->dec_ok) { r = fail(h, MOCR_ERR_INVALID, "run a decode first"); break; }
      const int rows = std::max(1, h->last_rows);
      PdParams p = make_pd_params(h, rows, h->cur_len, true, false);
      if (i == -2 && (r = identity_slots(h, rows)) != MOCR_OK) break;
      PdStage prog[kPdMaxStages];
      pd_build_program(p, prog);
      // the first stage of the named kind in the per-token program
      auto find_stage = [&](int type, int N, int K) {
        for (int j = 0; j < kPdMaxStages && prog[j].type != PD_NEXT; ++j)
          if (prog[j].type == type && (N == 0 || prog[j].N == N) && (K == 0 || prog[j].K == K)) return j;
        return -1;
      };
      int idx = -1;
      const double w768 = 2.0 * kD * kD, act = 2.0 * n * kD;
      if (k == "dec_qkv") { idx = find_stage(p.big ? PD_TC : PD_GEMM16, 3 * kD, kD); bytes = 3 * w768 + act + 3 * act; }
      else if (k == "dec_self_attn") {     // on the state the last decode left: every row holds last_steps cached tokens (K and V rows of one layer)
        idx = find_stage(PD_ATTN_SELF, 0, 0);
        bytes = static_cast<double>(n) * std::max(h->last_steps, 1) * 2.0 * kD * 2 + 2.0 * act;
      }
      else if (k == "dec_self_out") { idx = find_stage(p.big ? PD_TC : (p.fuse_ln ? PD_PROJ_LN : PD_GEMM16), kD, kD); bytes = w768 + act + (p.fuse_ln ? 3 * 2 * act : kPdSplit * 2 * act); }
      else if (k == "dec_ln") { idx = find_stage(PD_LN, 0, 0); bytes = (kPdSplit + 1) * 2 * act + 2 * act + act; }
      else if (k == "dec_cross_attn") { idx = find_stage(PD_ATTN_CROSS, 0, 0); bytes = static_cast<double>(n) * (2.0 * kEncTokens * kD * 2 + 2.0 * kD * 2); }
      else if (k == "dec_fc1") { idx = find_stage(p.big ? PD_TC : PD_GEMM32, kFFN, kD); bytes = 4 * w768 + act + 4 * act; }
      else if (k == "dec_fc2") { idx = find_stage(p.big ? PD_TC : PD_GEMM16, kD, kFFN); bytes = 4 * w768 + 4 * act + kPdSplit * 2 * act; }
      else if (k == "dec_vocab") { idx = find_stage(PD_GEMM48, kVocab, kD); bytes = 2.0 * kVocab * kD + act; }
      if (idx < 0) { r = fail(h, MOCR_ERR_INVALID, "no stage %s in the current decoder program", kernel); break; }
      const PdStage& st = prog[idx];
      cudaError_t le = cudaSuccess;
      // launched exactly as in the decode loop (programmatic dependent launch: the next launch's
      // constant / K-V prefetch overlaps the tail of the previous one)
      switch (st.type) {
        case PD_TC: r = launch_tc_stage(h, p, st); break;
        case PD_GEMM16: le = launch_pdl(h, pd_gemm_kernel<16, 2>, (st.N / 16) * st.ksplit, 128 * kPdStageKS, pd_gemm_smem_bytes(16), p, st); break;
        case PD_GEMM32: le = launch_pdl(h, pd_gemm_kernel<32, 2>, (st.N / 32) * st.ksplit, 128 * kPdStageKS, pd_gemm_smem_bytes(32), p, st); break;
        case PD_GEMM48:
          if (p.big || (h->dec_tc & 1)) {
            GemmArgs ga{};
            ga.part_max = p.part_max;
            ga.part_idx = p.part_idx;
            ga.step = p.pos;
            ga.tap_steps = p.max_len - 1;
            const int vbn = p.big ? big_vocab_tile(h, p.B) : 64;
            if (vbn == 256) r = launch_gemm_stage_tc<256, EPI_ARGMAX>(h, st.A, kD, h->head_dec, p.B, ga);
            else if (vbn == 128) r = launch_gemm_stage_tc<128, EPI_ARGMAX>(h, st.A, kD, h->head_dec, p.B, ga);
            else r = launch_gemm_stage_tc<64, EPI_ARGMAX>(h, st.A, kD, h->head_dec, p.B, ga);
          } else {
            le = launch_pdl(h, pd_gemm_kernel<48, 1>, (st.N / 48) * st.ksplit, 128 * kPdStageKS, pd_gemm_smem_bytes(48), p, st);
          }
          break;
        case PD_PROJ_LN: le = launch_pdl(h, pd_proj_ln_kernel<8, 3, 1>, ((p.B + kPlRows - 1) / kPlRows) * kPlCluster, 256, pd_proj_ln_smem_bytes(8), p, st); break;
        case PD_ATTN_SELF:
        case PD_ATTN_CROSS: le = launch_attention_stage(h, p, st); break;
        default: le = launch_pdl(h, pd_ln_kernel, (p.B + kPdWarps - 1) / kPdWarps, kPdThreads, 0, p, st); break;
      }
      if (le != cudaSuccess) r = fail(h, MOCR_ERR_CUDA, "stage kernel launch failed: %s", cudaGetErrorString(le));
      ++h->launches;
      flops = 0;
    } else {
      r = fail(h, MOCR_ERR_INVALID, "unknown kernel name %s", kernel);
    }
  }
  if (r == MOCR_OK) {
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_launch = ms / iters;
    if (algo_bytes) *algo_bytes = bytes;
    if (algo_flops) *algo_flops = flops;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return r;
}

// ---- kernel-level unit hooks (tests only): run ONE product kernel on caller-supplied data ----

int mocr_test_gemm(mocr_handle_t* h, int epi, int bn, int M, int N, int K, const float* A, const float* Wt, const float* bias,
                   const float* resid, float* out, int32_t* out_argmax) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (M < 1 || N < 1 || K < 1 || A == nullptr || Wt == nullptr || bias == nullptr || out == nullptr || (bn != 0 && N % bn != 0) || K % kGemmBK != 0 ||
      (bn == 0 && epi != EPI_F32_ACCUM))
    return fail(h, MOCR_ERR_INVALID, "bad test_gemm argument");
  if (epi == EPI_PATCH) return fail(h, MOCR_ERR_INVALID, "EPI_PATCH is covered by the encoder test");
  const int Mp = round_up(M, kGemmBM);
  std::vector<uint16_t> ab(static_cast<size_t>(Mp) * K, 0), wb(static_cast<size_t>(N) * K);
  for (size_t i = 0; i < static_cast<size_t>(M) * K; ++i) ab[i] = f32_to_bf16(A[i]);
  for (size_t i = 0; i < wb.size(); ++i) wb[i] = f32_to_bf16(Wt[i]);
  ActBuf a;
  Linear L;
  void *d_out = nullptr, *d_res = nullptr, *d_pm = nullptr, *d_pi = nullptr, *d_step = nullptr;
  const size_t mn = static_cast<size_t>(M) * N;
  int r = MOCR_OK;
  auto body = [&]() -> int {
    CK(cudaMalloc(reinterpret_cast<void**>(&a.p), ab.size() * 2));
    CK(cudaMalloc(reinterpret_cast<void**>(&L.w), wb.size() * 2));
    CK(cudaMalloc(reinterpret_cast<void**>(&L.bias), N * sizeof(float)));
    CK(cudaMalloc(&d_out, mn * sizeof(float)));
    CK(cudaMemcpy(a.p, ab.data(), ab.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(L.w, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(L.bias, bias, N * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0, mn * sizeof(float)));
    a.rows_cap = Mp;
    a.K = K;
    L.N = N;
    L.K = K;
    TRY(make_map(h, &a.map, a.p, Mp, K, kGemmBM));
    GemmArgs g{};
    g.out = d_out;
    g.ldo = N;
    if (epi == EPI_F32_RESID) {
      if (resid == nullptr) return fail(h, MOCR_ERR_INVALID, "resid is NULL");
      CK(cudaMalloc(&d_res, mn * sizeof(float)));
      CK(cudaMemcpy(d_res, resid, mn * sizeof(float), cudaMemcpyHostToDevice));
      g.resid = static_cast<const float*>(d_res);
      g.ldr = N;
    }
    if (epi == EPI_F32_ACCUM) {     // out starts as the residual and is accumulated in place
      if (resid == nullptr) return fail(h, MOCR_ERR_INVALID, "resid is NULL");
      CK(cudaMemcpy(d_out, resid, mn * sizeof(float), cudaMemcpyHostToDevice));
    }
    const int parts = bn > 0 ? 2 * (N / bn) : 0;
    if (epi == EPI_ARGMAX) {
      CK(cudaMalloc(&d_pm, static_cast<size_t>(M) * parts * sizeof(float)));
      CK(cudaMalloc(&d_pi, static_cast<size_t>(M) * parts * sizeof(int)));
      CK(cudaMalloc(&d_step, static_cast<size_t>(M) * sizeof(int)));
      CK(cudaMemset(d_step, 0, static_cast<size_t>(M) * sizeof(int)));
      g.part_max = static_cast<float*>(d_pm);
      g.part_idx = static_cast<int*>(d_pi);
      g.logits = static_cast<float*>(d_out);
      g.step = static_cast<const int*>(d_step);
      g.tap_steps = 1;
    }
    // the set-up above went through the legacy default stream (cudaMemset / cudaMemcpy), which does not order against the handle's
    // non-blocking stream: without this, a 150 MB memset of `out` can still be running when the GEMM writes it
    CK(cudaDeviceSynchronize());
    if (epi == EPI_F32_ACCUM && bn == 0) {     // the cluster K-split kernel of the large-batch decoder program
      if (N % kKsBN != 0 || K % (kGemmBK * kKsSplit) != 0) return fail(h, MOCR_ERR_INVALID, "K-split GEMM needs N %% 128 == 0 and K %% 256 == 0");
      h->pdl_now = 0;
      const int kr = launch_gemm_ksplit(h, a.p, K, L, M, static_cast<float*>(d_out));
      h->pdl_now = 1;
      TRY(kr);
    } else {
      TRY(gemm(h, epi, bn, a, L, M, g));
    }
    CK(cudaStreamSynchronize(h->stream));
    if (epi == EPI_BF16 || epi == EPI_BF16_GELU) {
      std::vector<uint16_t> ob(mn);
      CK(cudaMemcpy(ob.data(), d_out, mn * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < mn; ++i) {
        const uint32_t u = static_cast<uint32_t>(ob[i]) << 16;
        memcpy(&out[i], &u, 4);
      }
    } else {
      CK(cudaMemcpy(out, d_out, mn * sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (epi == EPI_ARGMAX && out_argmax != nullptr) {
      std::vector<float> pm(static_cast<size_t>(M) * parts);
      std::vector<int> pi(static_cast<size_t>(M) * parts);
      CK(cudaMemcpy(pm.data(), d_pm, pm.size() * sizeof(float), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(pi.data(), d_pi, pi.size() * sizeof(int), cudaMemcpyDeviceToHost));
      for (int m = 0; m < M; ++m) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int p = 0; p < parts; ++p) {
          const float v = pm[static_cast<size_t>(m) * parts + p];
          const int ix = pi[static_cast<size_t>(m) * parts + p];
          if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
        }
        out_argmax[m] = bi;
      }
    }
    return MOCR_OK;
  };
  r = body();
  cudaStreamSynchronize(h->stream);
  cudaFree(a.p); cudaFree(L.w); cudaFree(L.bias); cudaFree(d_out); cudaFree(d_res); cudaFree(d_pm); cudaFree(d_pi); cudaFree(d_step);
  return r;
}

int mocr_test_encoder_attention(mocr_handle_t* h, int n, const float* qkv, float* out) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (n < 1 || n > h->max_batch || qkv == nullptr || out == nullptr) return fail(h, MOCR_ERR_INVALID, "bad test_encoder_attention argument");
  const size_t rows = static_cast<size_t>(n) * kEncTokens;
  std::vector<uint16_t> qb(rows * 3 * kD);
  for (size_t i = 0; i < qb.size(); ++i) qb[i] = f32_to_bf16(qkv[i]);
  CK(cudaMemcpyAsync(h->qkv, qb.data(), qb.size() * 2, cudaMemcpyHostToDevice, h->stream));
  TRY(attention197(h, n));
  std::vector<uint16_t> ob(rows * kD);
  CK(cudaMemcpyAsync(ob.data(), h->ctx.p, ob.size() * 2, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i < ob.size(); ++i) {
    const uint32_t u = static_cast<uint32_t>(ob[i]) << 16;
    memcpy(&out[i], &u, 4);
  }
  h->pre_ok = h->enc_ok = h->dec_ok = false;   // scratch buffers were overwritten
  return MOCR_OK;
}

// ---- decoder stage kernels on caller-supplied data (tests only) ----

namespace {
std::vector<uint16_t> to_bf16(const float* src, size_t n) {
  std::vector<uint16_t> out(n);
  for (size_t i = 0; i < n; ++i) out[i] = f32_to_bf16(src[i]);
  return out;
}
void from_bf16(const uint16_t* src, float* dst, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    const uint32_t u = static_cast<uint32_t>(src[i]) << 16;
    memcpy(&dst[i], &u, 4);
  }
}
}  // namespace

int mocr_test_decode_attention(mocr_handle_t* h, int mode, int n_rows, int n_ctx, const int32_t* pos, const float* q, const float* k, const float* v,
                               const float* new_k, const float* new_v, float* out_ctx, float* out_k_row, float* out_v_row) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  const bool rows_kernel = mode >= 3;         // 3: self, 4: cross (bf16 query rows) through the warp-per-unit kernel of the large-batch program
  if (mode == 3) mode = 1;
  if (mode == 4) mode = 2;
  const bool self = mode == 1;
  if (mode < 0 || mode > 2 || n_rows < 1 || n_rows > h->max_batch || q == nullptr || k == nullptr || v == nullptr || out_ctx == nullptr)
    return fail(h, MOCR_ERR_INVALID, "bad test_decode_attention argument");
  if (self && (pos == nullptr || new_k == nullptr || new_v == nullptr || n_ctx < 1 || n_ctx > h->max_length))
    return fail(h, MOCR_ERR_INVALID, "self-attention test needs pos, new_k, new_v and 1 <= n_ctx <= max_length");
  if (!self && n_ctx != kEncTokens) return fail(h, MOCR_ERR_INVALID, "cross-attention runs over %d keys", kEncTokens);
  h->pre_ok = h->enc_ok = h->dec_ok = false;      // engine buffers are overwritten
  PdParams p = make_pd_params(h, n_rows, h->max_length, false, false);
  PdStage st{};
  st.layer = 0;
  std::vector<int> zeros(n_rows, 0);
  CK(cudaMemcpyAsync(h->d_finished, zeros.data(), sizeof(int) * n_rows, cudaMemcpyHostToDevice, h->stream));
  TRY(identity_slots(h, n_rows));
  float* d_zero_bias = nullptr;
  static bool done[16] = {};
  if (!done[h->device & 15]) {
    CK(cudaFuncSetAttribute(pd_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPdAttnSmemBytes));
    CK(cudaFuncSetAttribute(pd_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPdAttnSmemBytes));
    CK(cudaFuncSetAttribute(pd_attention_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPdAttnRowsSmemBytes));
    CK(cudaFuncSetAttribute(pd_attention_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPdAttnRowsSmemBytes));
    done[h->device & 15] = true;
  }
  const int grid = rows_kernel ? std::min(h->big_attn_grid > 0 ? h->big_attn_grid : n_rows * kHeads, (n_rows * kHeads + 3) / 4)
                               : (h->attn_grid > 0 ? std::min(h->attn_grid, n_rows * kHeads) : n_rows * kHeads);
  int r = MOCR_OK;
  auto body = [&]() -> int {
    if (self) {
      for (int b = 0; b < n_rows; ++b)
        if (pos[b] < 0 || pos[b] >= n_ctx) return fail(h, MOCR_ERR_INVALID, "pos[%d] = %d outside [0, n_ctx)", b, pos[b]);
      st.type = PD_ATTN_SELF;
      // cache rows [b][j][768], j < n_ctx (the row at pos[b] is overwritten by the kernel with this step's K/V)
      const std::vector<uint16_t> kb = to_bf16(k, static_cast<size_t>(n_rows) * n_ctx * kD), vb = to_bf16(v, static_cast<size_t>(n_rows) * n_ctx * kD);
      for (int b = 0; b < n_rows; ++b) {
        CK(cudaMemcpyAsync(h->self_k[0] + static_cast<size_t>(b) * h->max_length * kD, kb.data() + static_cast<size_t>(b) * n_ctx * kD,
                           static_cast<size_t>(n_ctx) * kD * 2, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->self_v[0] + static_cast<size_t>(b) * h->max_length * kD, vb.data() + static_cast<size_t>(b) * n_ctx * kD,
                           static_cast<size_t>(n_ctx) * kD * 2, cudaMemcpyHostToDevice, h->stream));
      }
      std::vector<uint16_t> qkv(static_cast<size_t>(n_rows) * 3 * kD);
      for (int b = 0; b < n_rows; ++b)
        for (int c = 0; c < kD; ++c) {
          qkv[(static_cast<size_t>(b) * 3 + 0) * kD + c] = f32_to_bf16(q[static_cast<size_t>(b) * kD + c]);
          qkv[(static_cast<size_t>(b) * 3 + 1) * kD + c] = f32_to_bf16(new_k[static_cast<size_t>(b) * kD + c]);
          qkv[(static_cast<size_t>(b) * 3 + 2) * kD + c] = f32_to_bf16(new_v[static_cast<size_t>(b) * kD + c]);
        }
      CK(cudaMemcpyAsync(h->d_qkv, qkv.data(), qkv.size() * 2, cudaMemcpyHostToDevice, h->stream));
      CK(cudaMemcpyAsync(h->d_pos, pos, sizeof(int) * n_rows, cudaMemcpyHostToDevice, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      if (rows_kernel) CK(launch_pdl(h, pd_attention_rows_kernel<true>, grid, 128, kPdAttnRowsSmemBytes, p, st));
      else CK(launch_pdl(h, pd_attention_kernel<true>, grid, 128, kPdAttnSmemBytes, p, st));
    } else {
      st.type = PD_ATTN_CROSS;
      // [crop][layer][K|V][head][197][64]
      std::vector<uint16_t> kv(static_cast<size_t>(n_rows) * 4 * kHeads * kEncTokens * kHeadDim, 0);
      for (int b = 0; b < n_rows; ++b)
        for (int j = 0; j < kEncTokens; ++j)
          for (int c = 0; c < kD; ++c) {
            const int hd = c / kHeadDim, d = c % kHeadDim;
            const size_t src = (static_cast<size_t>(b) * kEncTokens + j) * kD + c;
            const size_t base = static_cast<size_t>(b) * 4 * kHeads * kEncTokens * kHeadDim;
            kv[base + ((0 * kHeads + hd) * kEncTokens + j) * kHeadDim + d] = f32_to_bf16(k[src]);
            kv[base + ((1 * kHeads + hd) * kEncTokens + j) * kHeadDim + d] = f32_to_bf16(v[src]);
          }
      CK(cudaMemcpyAsync(h->crosskv, kv.data(), kv.size() * 2, cudaMemcpyHostToDevice, h->stream));
      if (mode == 2) {         // complete bf16 query rows (the large-batch program)
        const std::vector<uint16_t> qb = to_bf16(q, static_cast<size_t>(n_rows) * kD);
        CK(cudaMemcpyAsync(h->d_q, qb.data(), qb.size() * 2, cudaMemcpyHostToDevice, h->stream));
        st.A = h->d_q;
      } else {                 // fp32 split-K partials + bias: the query is spread over the partials as 1/2, 1/4, 1/4
        CK(cudaMalloc(reinterpret_cast<void**>(&d_zero_bias), kD * sizeof(float)));
        CK(cudaMemsetAsync(d_zero_bias, 0, kD * sizeof(float), h->stream));
        std::vector<float> parts(static_cast<size_t>(kPdSplit) * n_rows * kD);
        for (size_t i = 0; i < static_cast<size_t>(n_rows) * kD; ++i) {
          parts[i] = 0.5f * q[i];
          parts[static_cast<size_t>(n_rows) * kD + i] = 0.25f * q[i];
          parts[2 * static_cast<size_t>(n_rows) * kD + i] = 0.25f * q[i];
        }
        CK(cudaMemcpyAsync(h->d_yq, parts.data(), parts.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        st.src = h->d_yq;
        st.parts = kPdSplit;
        st.bias = d_zero_bias;
      }
      CK(cudaStreamSynchronize(h->stream));
      if (rows_kernel) CK(launch_pdl(h, pd_attention_rows_kernel<false>, grid, 128, kPdAttnRowsSmemBytes, p, st));
      else CK(launch_pdl(h, pd_attention_kernel<false>, grid, 128, kPdAttnSmemBytes, p, st));
    }
    ++h->launches;
    std::vector<uint16_t> cb(static_cast<size_t>(n_rows) * kD);
    CK(cudaMemcpyAsync(cb.data(), h->d_ctx.p, cb.size() * 2, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    from_bf16(cb.data(), out_ctx, cb.size());
    if (self && out_k_row != nullptr && out_v_row != nullptr) {
      std::vector<uint16_t> row(kD);
      for (int b = 0; b < n_rows; ++b) {
        CK(cudaMemcpy(row.data(), h->self_k[0] + (static_cast<size_t>(b) * h->max_length + pos[b]) * kD, kD * 2, cudaMemcpyDeviceToHost));
        from_bf16(row.data(), out_k_row + static_cast<size_t>(b) * kD, kD);
        CK(cudaMemcpy(row.data(), h->self_v[0] + (static_cast<size_t>(b) * h->max_length + pos[b]) * kD, kD * 2, cudaMemcpyDeviceToHost));
        from_bf16(row.data(), out_v_row + static_cast<size_t>(b) * kD, kD);
      }
    }
    return MOCR_OK;
  };
  r = body();
  cudaStreamSynchronize(h->stream);
  if (d_zero_bias) cudaFree(d_zero_bias);
  return r;
}

// kind: 0 bf16 (QKV-like, 16-column tiles), 1 bf16 + GELU (FFN1-like, 32-column tiles), 2 split-K fp32 partials (returned summed, + bias),
// 3 vocabulary arg-max on the mma.sync kernel (out = logits), 4 projection + residual + LayerNorm in the cluster kernel (N = K = 768),
// 5 the same projection as split-K partials + the LayerNorm row stage.  resid / gamma / beta: kinds 4, 5 (resid may be NULL); gelu: kinds 4, 5.
int mocr_test_stage_gemm(mocr_handle_t* h, int kind, int n_rows, int N, int K, const float* A, const float* Wt, const float* bias, const float* resid,
                         const float* gamma, const float* beta, int gelu, float* out, int32_t* out_argmax) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (kind < 0 || kind > 5 || n_rows < 1 || n_rows > h->max_batch || A == nullptr || Wt == nullptr || bias == nullptr || out == nullptr)
    return fail(h, MOCR_ERR_INVALID, "bad test_stage_gemm argument");
  const int nt = kind == 1 ? 32 : (kind == 3 ? 48 : 16);
  const int ksplit = (kind == 2 || kind == 5) ? kPdSplit : 1;
  if (N % nt != 0 || K % (32 * kPdStageKS * ksplit * 2) != 0 || N > 3 * kFFN * 2 || K > kFFN)
    return fail(h, MOCR_ERR_INVALID, "unsupported stage GEMM shape N=%d K=%d for kind %d", N, K, kind);
  if ((kind == 4 || kind == 5) && (N != kD || (kind == 4 && K != kD) || gamma == nullptr || beta == nullptr))
    return fail(h, MOCR_ERR_INVALID, "the LayerNorm kinds need N = 768 (K = 768 for the cluster kernel), gamma and beta");
  if (kind == 3 && N > kVocab) return fail(h, MOCR_ERR_INVALID, "arg-max stage: N <= %d", kVocab);
  h->pre_ok = h->enc_ok = h->dec_ok = false;
  const size_t mn = static_cast<size_t>(n_rows) * N;
  __nv_bfloat16 *d_a = nullptr, *d_w = nullptr, *d_ob = nullptr;
  float *d_bias = nullptr, *d_of = nullptr, *d_res = nullptr, *d_g = nullptr, *d_b = nullptr, *d_x = nullptr;
  int r = MOCR_OK;
  static bool done[16] = {};
  auto body = [&]() -> int {
    if (!done[h->device & 15]) {
      CK(cudaFuncSetAttribute(pd_gemm_kernel<48, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, pd_gemm_smem_bytes(48)));
      CK(cudaFuncSetAttribute(pd_proj_ln_kernel<8, 3, 1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      done[h->device & 15] = true;
    }
    const std::vector<uint16_t> ab = to_bf16(A, static_cast<size_t>(n_rows) * K), wb = to_bf16(Wt, static_cast<size_t>(N) * K);
    CK(cudaMalloc(reinterpret_cast<void**>(&d_a), ab.size() * 2));
    CK(cudaMalloc(reinterpret_cast<void**>(&d_w), wb.size() * 2));
    CK(cudaMalloc(reinterpret_cast<void**>(&d_bias), N * sizeof(float)));
    CK(cudaMalloc(reinterpret_cast<void**>(&d_ob), mn * 2));
    CK(cudaMalloc(reinterpret_cast<void**>(&d_of), mn * sizeof(float) * kPdSplit));
    CK(cudaMemcpy(d_a, ab.data(), ab.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_w, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_bias, bias, N * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_of, 0, mn * sizeof(float) * kPdSplit));
    if (kind >= 4) {
      CK(cudaMalloc(reinterpret_cast<void**>(&d_g), kD * sizeof(float)));
      CK(cudaMalloc(reinterpret_cast<void**>(&d_b), kD * sizeof(float)));
      CK(cudaMalloc(reinterpret_cast<void**>(&d_x), mn * sizeof(float)));
      CK(cudaMemcpy(d_g, gamma, kD * sizeof(float), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_b, beta, kD * sizeof(float), cudaMemcpyHostToDevice));
      if (resid != nullptr) {
        CK(cudaMalloc(reinterpret_cast<void**>(&d_res), mn * sizeof(float)));
        CK(cudaMemcpy(d_res, resid, mn * sizeof(float), cudaMemcpyHostToDevice));
      }
    }
    PdParams p = make_pd_params(h, n_rows, h->max_length, false, false);
    p.n_partials = kPdVocabTiles;
    if (kind == 3) {
      std::vector<int> zeros(n_rows, 0);
      CK(cudaMemcpy(h->d_pos, zeros.data(), sizeof(int) * n_rows, cudaMemcpyHostToDevice));
      p.logits = d_of;
      p.logits_cur = 1;
      if (N != kVocab) return fail(h, MOCR_ERR_INVALID, "arg-max stage test runs at N = %d", kVocab);
    }
    CK(cudaDeviceSynchronize());      // (legacy-stream set-up above vs the handle's non-blocking stream)
    const PdLinear lin{d_w, d_bias};
    const PdLn ln{d_g, d_b};
    if (kind == 4) {
      const PdStage st = pd_proj_ln_desc(d_a, K, lin, gelu, d_res, ln, d_x, d_ob);
      CK(launch_pdl(h, pd_proj_ln_kernel<8, 3, 1>, ((n_rows + kPlRows - 1) / kPlRows) * kPlCluster, 256, pd_proj_ln_smem_bytes(8), p, st));
    } else {
      const int epi = kind == 0 ? PD_BF16 : (kind == 1 ? PD_BF16_GELU : (kind == 3 ? PD_ARGMAX : PD_F32_PARTIAL));
      const PdStage st = pd_gemm_desc(kind == 1 ? PD_GEMM32 : (kind == 3 ? PD_GEMM48 : PD_GEMM16), epi, d_a, K, lin, N, ksplit, d_ob, d_of, N);
      const int grid = (N / nt) * ksplit;
      if (nt == 16) CK(launch_pdl(h, pd_gemm_kernel<16, 2>, grid, 128 * kPdStageKS, pd_gemm_smem_bytes(16), p, st));
      else if (nt == 32) CK(launch_pdl(h, pd_gemm_kernel<32, 2>, grid, 128 * kPdStageKS, pd_gemm_smem_bytes(32), p, st));
      else CK(launch_pdl(h, pd_gemm_kernel<48, 1>, grid, 128 * kPdStageKS, pd_gemm_smem_bytes(48), p, st));
      if (kind == 5) {
        const PdStage lst = pd_ln_desc(d_of, kPdSplit, d_bias, gelu, d_res, ln, d_x, d_ob);
        CK(launch_pdl(h, pd_ln_kernel, (n_rows + 1) / 2, 64, 0, p, lst));
        ++h->launches;
      }
    }
    ++h->launches;
    CK(cudaStreamSynchronize(h->stream));
    if (kind == 0 || kind == 1) {
      std::vector<uint16_t> ob(mn);
      CK(cudaMemcpy(ob.data(), d_ob, mn * 2, cudaMemcpyDeviceToHost));
      from_bf16(ob.data(), out, mn);
    } else if (kind == 2) {
      std::vector<float> parts(mn * kPdSplit);
      CK(cudaMemcpy(parts.data(), d_of, parts.size() * sizeof(float), cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < mn; ++i) out[i] = bias[i % N] + parts[i] + parts[mn + i] + parts[2 * mn + i];
    } else if (kind == 3) {
      CK(cudaMemcpy(out, d_of, mn * sizeof(float), cudaMemcpyDeviceToHost));
      if (out_argmax != nullptr) {
        std::vector<float> pm(static_cast<size_t>(n_rows) * kPdVocabTiles);
        std::vector<int> pi(pm.size());
        CK(cudaMemcpy(pm.data(), h->part_max, pm.size() * sizeof(float), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(pi.data(), h->part_idx, pi.size() * sizeof(int), cudaMemcpyDeviceToHost));
        for (int m = 0; m < n_rows; ++m) {
          float bv = -INFINITY;
          int bi = 0x7fffffff;
          for (int t = 0; t < kPdVocabTiles; ++t) {
            const float vv = pm[static_cast<size_t>(m) * kPdVocabTiles + t];
            const int ix = pi[static_cast<size_t>(m) * kPdVocabTiles + t];
            if (vv > bv || (vv == bv && ix < bi)) { bv = vv; bi = ix; }
          }
          out_argmax[m] = bi;
        }
      }
    } else {
      CK(cudaMemcpy(out, d_x, mn * sizeof(float), cudaMemcpyDeviceToHost));
    }
    return MOCR_OK;
  };
  r = body();
  cudaStreamSynchronize(h->stream);
  cudaFree(d_a); cudaFree(d_w); cudaFree(d_bias); cudaFree(d_ob); cudaFree(d_of); cudaFree(d_res); cudaFree(d_g); cudaFree(d_b); cudaFree(d_x);
  return r;
}

// Host-only: the Pillow-exact coefficient table the preprocess kernel reads for one input
// extent (xmin[224] | count[224] | k[224*ksize], int32).  No CUDA call; usable without a GPU.
int mocr_resample_table(int in_size, int32_t* ksize, int32_t* out, int capacity) {
  if (in_size < 1 || in_size > 32768 || ksize == nullptr) return MOCR_ERR_INVALID;
  ResampleTable t = make_resample_table(in_size);
  *ksize = t.ksize;
  if (out != nullptr) {
    if (capacity < static_cast<int>(t.data.size())) return MOCR_ERR_CAPACITY;
    memcpy(out, t.data.data(), t.data.size() * sizeof(int));
  }
  return static_cast<int>(t.data.size());
}

// Debug / tuning: clock64 timeline of CTA 0 of the last persistent decode (option decode_prof=1).
int mocr_get_decode_profile(mocr_handle_t* h, int64_t* out, int n) {
  TRY(check_handle(h));
  std::lock_guard<std::mutex> lock(h->mu);
  if (out == nullptr || n < 1 || n > 4096) return fail(h, MOCR_ERR_INVALID, "bad argument");
  return d2h(h, out, h->d_prof, static_cast<size_t>(n) * sizeof(long long));
}

const char* mocr_last_error(mocr_handle_t* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

}  // extern "C"
