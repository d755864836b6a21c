/* mocr_b200.h - C ABI of the B200-native Manga-OCR recognition engine.
 *
 * This is the drop-in boundary for ONE hot path of irazawa/Manga-OCR: the call
 *     raw_text = self.manga_ocr_reader(pil_img)         reference/src/ui/main_window.py:9801
 * where manga_ocr_reader = MangaOcr()                    reference/src/ui/main_window.py:3394
 * and MangaOcr comes from `from manga_ocr import MangaOcr`   reference/src/core/config.py:433.
 * The reference has no FFI of its own (it is pure Python); the entry points below are what a
 * ctypes binding for that call needs (INTEGRATION.md shows the binding).  Plain pointers and
 * sizes only; no C++ or torch types; every function returns 0 or a negative mocr_status and
 * never throws or aborts; mocr_last_error() gives the message of the last failure.
 *
 * Threading: one handle may be used from many host threads (the app calls the engine from up
 * to 50 worker threads, reference/src/ui/main_window.py:608-611,4317-4327); calls on the same
 * handle are serialised internally.
 */
#ifndef MOCR_B200_H_
#define MOCR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOCR_ABI_VERSION 1

typedef struct mocr_handle mocr_handle_t;
typedef struct mocr_beam mocr_beam_t;   /* beam-search bookkeeping (host) */

enum mocr_status {
  MOCR_OK = 0,
  MOCR_ERR_INVALID = -1,   /* bad argument / bad state                       */
  MOCR_ERR_CUDA = -2,      /* a CUDA call or kernel failed                   */
  MOCR_ERR_WEIGHTS = -3,   /* missing / mis-shaped tensor                    */
  MOCR_ERR_CAPACITY = -4,  /* batch, length or crop size beyond the handle's */
  MOCR_ERR_NO_DEVICE = -5  /* no usable sm_100 device                        */
};

/* One text-bubble crop as the app hands it over: interleaved uint8 pixels in HOST memory.
 * Replaces the PIL.Image argument of MangaOcr.__call__ (reference/src/ui/main_window.py:9800:
 * Image.fromarray(cv2.cvtColor(img, cv2.COLOR_BGR2RGB)) - a contiguous uint8 [H,W,3] array). */
typedef struct {
  const uint8_t* data; /* pixel (0,0)                                      */
  int32_t height;      /* >= 1                                             */
  int32_t width;       /* >= 1                                             */
  int32_t stride;      /* bytes between rows (>= width * channels)         */
  int32_t channels;    /* 1 = luma, 3 = RGB/BGR, 4 = RGBA/BGRA (alpha ignored) */
} mocr_crop_t;

enum mocr_channel_order { MOCR_RGB = 0, MOCR_BGR = 1 };

/* One selection on a page, as the reference stages it before the engine (SURVEY.md section 8f N2;
 * reference/src/ui/main_window.py:6497-6506 polygon selection, :6429-6430 rectangle selection,
 * :9789-9795 rotation by text orientation):
 *   crop    = page.crop((left, top, right, bottom))          PIL semantics: right/bottom exclusive,
 *                                                            pixels outside the page are 0
 *   polygon : mask = cv2.fillPoly(zeros, [polygon - (left, top)], 255); crop = mask ? crop : 255
 *             (points in PAGE coordinates; NULL / 0 points = no mask)
 *   rotate  : cv2.rotate(crop, ROTATE_90_CLOCKWISE | ROTATE_90_COUNTERCLOCKWISE)
 * NB the reference passes QRect.right()/bottom() (= x + w - 1) as the exclusive end, so its crops are
 * one pixel short of the polygon's bounding box; pass the same numbers to reproduce it. */
enum mocr_rotation { MOCR_ROT_NONE = 0, MOCR_ROT_CW = 1, MOCR_ROT_CCW = 2 };
typedef struct {
  int32_t left, top, right, bottom;
  const int32_t* polygon; /* n_points (x, y) pairs, page coordinates */
  int32_t n_points;       /* 0 .. 1024 */
  int32_t rotate;         /* enum mocr_rotation */
} mocr_region_t;

/* ---- life cycle -------------------------------------------------------------------- */

int mocr_abi_version(void);

/* Replaces MangaOcr.__init__ (model construction + .cuda()).  max_batch = crops decoded
 * together; max_length = decode cap (the reference hard-codes 300, <= 512). */
int mocr_create(int device, int max_batch, int max_length, mocr_handle_t** out);
int mocr_destroy(mocr_handle_t* h);

/* Weights by their reference state_dict name, fp32, host memory,
 * row-major.  Replaces VisionEncoderDecoderModel.from_pretrained.  After the last tensor call
 * mocr_finalize_weights once: it folds / converts / uploads (bf16 matrices, fp32 vectors). */
int mocr_set_weight(mocr_handle_t* h, const char* name, const float* data, const int64_t* shape, int ndim);
int mocr_finalize_weights(mocr_handle_t* h);

/* ---- the path, all in one ------------------------------------------------------------- */

/* n crops (any n; processed in chunks of max_batch) -> greedy token ids.
 * Replaces processor(...) + model.generate(x[None], max_length=300)[0] of MangaOcr.__call__.
 * out_ids: [n, max_length] int32, row i = [CLS] t1 t2 ... (EOS included when produced), padded
 * with PAD(0); out_lens[i] = number of valid ids in row i. */
int mocr_recognize(mocr_handle_t* h, const mocr_crop_t* crops, int n, int channel_order, int max_length,
                   int32_t* out_ids, int32_t* out_lens);

/* The same for n selections of ONE page: the page is uploaded once and every selection's crop,
 * polygon composite on white and rotation happen on the device, inside the preprocess reads.
 * Replaces the per-selection PIL crop / cv2 composite / rotate / colour round trips in front of
 * MangaOcr.__call__ (reference/src/ui/main_window.py:6497-6506, 9789-9800). */
int mocr_recognize_regions(mocr_handle_t* h, const mocr_crop_t* page, const mocr_region_t* regions, int n, int channel_order,
                           int max_length, int32_t* out_ids, int32_t* out_lens);

/* ---- the path, stage by stage (n <= max_batch) ------------------------------------------ */

/* Host crops -> pinned staging -> device arena (async on the handle's stream). */
int mocr_stage_crops(mocr_handle_t* h, const mocr_crop_t* crops, int n, int channel_order);
/* Page + selections -> device arena, polygon masks rasterised on the device (n <= max_batch);
 * the alternative to mocr_stage_crops in front of mocr_preprocess. */
int mocr_stage_regions(mocr_handle_t* h, const mocr_crop_t* page, const mocr_region_t* regions, int n, int channel_order);
/* Fused luma + Pillow-exact bilinear 224x224 + patch rows, on the staged crops.
 * Replaces img.convert("L").convert("RGB") and ViTImageProcessor (resize, rescale, normalize). */
int mocr_preprocess(mocr_handle_t* h);
/* ViT-base encoder on the patch rows + the once-per-crop cross-attention K/V projection.
 * Replaces ViTModel.forward and BertCrossAttention's K/V Linear of decode step 0. */
int mocr_encode(mocr_handle_t* h);
/* Batched greedy decode of the encoded crops; ids stay on the device until mocr_fetch_ids.
 * forced_ids (host, [n, max_length], may be NULL): teacher forcing for parity tests - the
 * input token of step t is forced_ids[i][t] while the arg-max of every step is still recorded.
 * Replaces GenerationMixin._sample (greedy) over BertLMHeadModel. */
int mocr_decode_greedy(mocr_handle_t* h, int max_length, const int32_t* forced_ids);
int mocr_fetch_ids(mocr_handle_t* h, int32_t* out_ids /*[n,max_length]*/, int32_t* out_lens /*[n]*/);
/* stage_crops must have been called; runs preprocess + encode + decode with no host copies. */
int mocr_run_resident(mocr_handle_t* h, int max_length);

/* ---- beam search (the shipped checkpoint's generation config is believed to be
 *      num_beams=4, no_repeat_ngram_size=3, length_penalty=2.0) -------------------------------------------
 * Replaces GenerationMixin._beam_search (transformers generation/utils.py:3076-3370) with the
 * NoRepeatNGramLogitsProcessor (logits_process.py:1012-1136) and the MaxLength / EOS stopping criteria.
 * mocr_decode_beam runs on the encoded crops (n * num_beams <= max_batch): every step is the greedy path's
 * stage kernels over n * num_beams rows that share their crop's cross-attention K/V, a device kernel that turns
 * the step's logits into each row's 2 * num_beams best continuations (log-softmax, n-gram ban), and the host
 * bookkeeping below; the self-attention cache follows the surviving beams.
 * early_stopping: 0 = False (heuristic), 1 = True, 2 = "never".  out_ids [n, max_length] (best hypothesis per crop,
 * filled with EOS past its end exactly as the reference does when the PAD id is 0), out_lens [n], out_scores [n] (sum of log-probabilities / length^length_penalty). */
int mocr_decode_beam(mocr_handle_t* h, int num_beams, int max_length, int no_repeat_ngram_size, float length_penalty, int early_stopping,
                     int32_t* out_ids, int32_t* out_lens, float* out_scores);
/* Crops (or the selections of one page) -> beam-search hypotheses in ONE call, any n (chunks of max_batch / num_beams
 * crops).  This is what generate() does when the checkpoint's generation config carries num_beams > 1, and the
 * handle stays locked from staging to the result: concurrent callers cannot interleave between the stages. */
int mocr_recognize_beam(mocr_handle_t* h, const mocr_crop_t* crops, int n, int channel_order, int max_length, int num_beams,
                        int no_repeat_ngram_size, float length_penalty, int early_stopping, int32_t* out_ids, int32_t* out_lens,
                        float* out_scores);
int mocr_recognize_regions_beam(mocr_handle_t* h, const mocr_crop_t* page, const mocr_region_t* regions, int n, int channel_order,
                                int max_length, int num_beams, int no_repeat_ngram_size, float length_penalty, int early_stopping,
                                int32_t* out_ids, int32_t* out_lens, float* out_scores);
/* The bookkeeping alone (no device needed): n crops, rows = n * num_beams, K = 2 * num_beams candidates per row. */
int mocr_beam_create(int n, int num_beams, int max_length, int no_repeat_ngram_size, float length_penalty, int early_stopping,
                     mocr_beam_t** out);
int mocr_beam_destroy(mocr_beam_t* b);
/* Tokens row `row` may not produce next (n-gram ban); returns their number (writes at most cap). */
int mocr_beam_banned(const mocr_beam_t* b, int row, int32_t* out, int cap);
/* cand_logprob / cand_token: [rows, K], every row sorted by descending log-probability (banned tokens excluded).
 * next_tokens / parents: [rows] the token each running row consumes next and the row it continues.
 * Returns 1 while the search goes on, 0 when it is finished, < 0 on error. */
int mocr_beam_step(mocr_beam_t* b, const float* cand_logprob, const int32_t* cand_token, int32_t* next_tokens, int32_t* parents);
int mocr_beam_result(const mocr_beam_t* b, int32_t* out_ids /*[n,max_length]*/, int32_t* out_lens /*[n]*/, float* out_scores /*[n]*/);

/* ---- parity taps (tests only; enable before the stage they tap) ------------------------ */

enum mocr_tap { MOCR_TAP_PIXELS = 1, MOCR_TAP_ENCODER = 2, MOCR_TAP_LOGITS = 4 };
int mocr_set_taps(mocr_handle_t* h, int taps);
int mocr_get_pixels_u8(mocr_handle_t* h, uint8_t* out /*[n,224,224]*/);
int mocr_get_pixel_values(mocr_handle_t* h, float* out /*[n,224,224] (the 3 planes are equal)*/);
int mocr_get_encoder_hidden(mocr_handle_t* h, float* out /*[n,197,768]*/);
/* Polygon mask of staged region `index` as rasterised on the device: [bottom-top, right-left] uint8 (0 / 255). */
int mocr_get_region_mask(mocr_handle_t* h, int index, uint8_t* out);
int mocr_get_step_logits(mocr_handle_t* h, float* out /*[n,max_length-1,6144]*/);

/* Kernel-level unit hooks (tests only): run ONE product kernel on caller-supplied host data.
 * epi: 0 bf16, 1 bf16+GELU, 2 f32+residual, 4 arg-max (out = logits), 5 f32+GELU, 7 in-place f32 accumulate (resid = initial out;
 * bn = 0 selects the cluster K-split kernel of the large-batch decoder program).
 * out = epilogue(A[M,K] * Wt[N,K]^T + bias), inputs rounded to bf16 exactly as the engine stores them.
 * Note the attention hook expects the 1/sqrt(64) scale already folded into q. */
int mocr_test_gemm(mocr_handle_t* h, int epi, int bn, int M, int N, int K, const float* A, const float* Wt, const float* bias,
                   const float* resid, float* out, int32_t* out_argmax);
int mocr_test_encoder_attention(mocr_handle_t* h, int n, const float* qkv /*[n*197,2304]*/, float* out /*[n*197,768]*/);

/* Decoder stage kernels on caller-supplied host data (n_rows <= max_batch), inputs rounded to bf16 as the engine stores them.
 * mocr_test_decode_attention: one query per row against its keys/values, ctx [n_rows, 768] out (q pre-scaled by 1/8).
 *   mode 1 = self-attention (modeling_bert.py:143-207): k/v [n_rows, n_ctx, 768] is the row's cache, row b attends to its first
 *            pos[b] cached keys plus this step's new_k/new_v [n_rows, 768], which the kernel also appends at index pos[b]
 *            (returned in out_k_row / out_v_row when not NULL); n_ctx <= max_length
 *   mode 0 = cross-attention (modeling_bert.py:210-284) over k/v [n_rows, 197, 768], query as fp32 split-K partials
 *   mode 2 = the same with complete bf16 query rows (the large-batch program)
 *   mode 3 / 4 = mode 1 / mode 2 through the warp-per-unit attention kernel of the large-batch program
 * mocr_test_stage_gemm: out[n_rows, N] of one small-M GEMM stage; kind 0 bf16, 1 bf16 + GELU, 2 split-K partials (summed + bias),
 *   3 vocabulary arg-max (out = logits, N = 6144), 4 projection + residual + LayerNorm fused in the 16-CTA cluster kernel
 *   (N = K = 768), 5 the same as split-K partials + the LayerNorm row stage. */
int mocr_test_decode_attention(mocr_handle_t* h, int mode, int n_rows, int n_ctx, const int32_t* pos, const float* q, const float* k, const float* v,
                               const float* new_k, const float* new_v, float* out_ctx, float* out_k_row, float* out_v_row);
int mocr_test_stage_gemm(mocr_handle_t* h, int kind, int n_rows, int N, int K, const float* A, const float* Wt, const float* bias, const float* resid,
                         const float* gamma, const float* beta, int gelu, float* out, int32_t* out_argmax);

/* Host-only (no CUDA call): the Pillow-exact resampling table for one input extent, as the
 * preprocess kernel reads it: xmin[224] | count[224] | k[224*ksize] (int32).  Returns the number
 * of int32 written (or needed when out is NULL), negative on error. */
int mocr_resample_table(int in_size, int32_t* ksize, int32_t* out, int capacity);

/* ---- admission into a running decode ------------------------------------------------------
 * The reference decodes one crop per call (reference/src/ui/main_window.py:9801) from up to 50 worker threads
 * (:608-611, :4317-4327); a batch API makes late callers wait for the whole batch in flight.  A session keeps `rows`
 * decoder rows stepping and admits crops while it runs: mocr_session_add stages, preprocesses, encodes and publishes
 * n crops (n <= free slots; out_slots[n] = the slot of each, 0 <= slot < max_batch, not necessarily adjacent: the crops of one
 * call are encoded in one pass whichever slots are free), an idle row picks each of them
 * up (the admission runs on a second stream: decode steps already launched keep running meanwhile);
 * mocr_session_run launches `steps` greedy steps followed by a snapshot of the slots' lengths (at most two snapshots may
 * be pending) and, when out_lens is not NULL, waits for the OLDEST pending snapshot and returns it as
 * out_lens[max_batch]: > 0 for a slot whose crop has finished (its id count), 0 for running or unused slots.  out_lens
 * NULL = launch only; steps 0 = read only: "launch, admit, launch, read" keeps one chunk queued while the host works.
 * mocr_session_fetch copies the id rows of finished slots
 * (out_ids [n, max_length], PAD-filled) and, with release != 0, frees the slots for reuse.  ids per crop equal
 * mocr_recognize's.  Other compute entry points of the handle fail while a session is active. */
int mocr_session_begin(mocr_handle_t* h, int channel_order, int max_length, int rows);
int mocr_session_add(mocr_handle_t* h, const mocr_crop_t* crops, int n, int32_t* out_slots);
int mocr_session_run(mocr_handle_t* h, int steps, int32_t* out_lens);
int mocr_session_fetch(mocr_handle_t* h, const int32_t* slots, int n, int32_t* out_ids, int release);
/* The chunks launched from here on step only the first `rows` decoder rows (rounded up to a row count the session has a step
 * program for: 16, or the session's own; that count is returned).  A lightly loaded session steps faster on fewer rows (92 vs
 * 119 us per step).  Growing is always allowed - rows above the current count are idle by construction; shrinking only while no
 * slot is in use.  A session begins on all its rows. */
int mocr_session_rows(mocr_handle_t* h, int rows);
int mocr_session_end(mocr_handle_t* h);

/* ---- plumbing --------------------------------------------------------------------------- */

void* mocr_stream(mocr_handle_t* h);            /* the cudaStream_t all work is launched on   */
int mocr_sync(mocr_handle_t* h);
int64_t mocr_launch_count(mocr_handle_t* h);    /* kernels launched by this handle so far     */
int mocr_last_steps(mocr_handle_t* h);          /* decode steps executed by the last decode   */
int mocr_set_option(mocr_handle_t* h, const char* key, int value);   /* tuning switches; fails while a session is active */
/* Times `iters` back-to-back launches of one named kernel of the path on the handle's stream
 * with CUDA events, on the current batch state (used by bench.py for the roofline line). */
int mocr_time_kernel(mocr_handle_t* h, const char* kernel, int iters, float* ms_per_launch, double* algo_bytes,
                     double* algo_flops);
/* Debug / tuning: SM-clock timeline (wait-exit / arrive of every stage, CTA 0) of the last
 * persistent decode, recorded when option "decode_prof" is 1. */
int mocr_get_decode_profile(mocr_handle_t* h, int64_t* out, int n);
const char* mocr_last_error(mocr_handle_t* h);  /* h may be NULL: error of the last failed create */

#ifdef __cplusplus
}
#endif
#endif /* MOCR_B200_H_ */
