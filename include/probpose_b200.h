/*
 * probpose_b200.h -- C ABI of the B200-native ProbPose heatmap hot path.
 *
 * The reference (zir-vision/ProbPose_pytorch) is pure Python and has no FFI;
 * the "interface each entry point replaces" is therefore the Python call the
 * reference makes at the cited file:line.  The Python shims in
 * probpose_pytorch_b200/ keep those call signatures and bind these symbols
 * with ctypes (see INTEGRATION.md for the binding a reference maintainer adds).
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer
 *     unless its name ends in _host;
 *   - the library never allocates, frees or owns caller memory, never
 *     synchronises the stream and keeps no mutable global state besides a
 *     thread-local last-error string and per-function attribute caching;
 *   - every function returns pp_status (0 = ok, negative = error) and never
 *     throws; pp_last_error_string() describes the last failure of the
 *     calling thread;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - heatmaps are dense row-major (B, K, H, W); "N" below means B*K heatmaps.
 */
#ifndef PROBPOSE_B200_H_
#define PROBPOSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_ABI_VERSION 2
#define PP_MAX_OKS_RADIUS 9   /* ceil(3 * 3.0): s is clipped to [0.55, 3.0] (heatmap.py:178-179) */
#define PP_OKS_TAPS (2 * PP_MAX_OKS_RADIUS + 1)
#define PP_MAX_BLUR_KSIZE 31

#if defined(__GNUC__)
#define PP_API __attribute__((visibility("default")))
#else
#define PP_API
#endif

typedef void* pp_stream_t;

typedef enum pp_status {
  PP_OK = 0,
  PP_ERR_INVALID_ARG = -1,
  PP_ERR_UNSUPPORTED_SHAPE = -2,
  PP_ERR_CUDA = -3,
  PP_ERR_SCRATCH = -4
} pp_status;

typedef enum pp_dtype { PP_F32 = 0, PP_BF16 = 1, PP_F64 = 2 } pp_dtype;

struct pp_mailbox;   /* peer-memory exchange of a step's results between GPUs, defined with pp_pack_records below */

/* ---- library ---------------------------------------------------------- */
PP_API int pp_version(void);
/* hash of the CUDA sources this library was compiled from (the build passes it in); lets a binding detect a stale build */
PP_API const char* pp_source_hash(void);
PP_API const char* pp_last_error_string(void);
/* number of SMs and opt-in shared memory per block of the current device */
PP_API int pp_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* smem_optin_bytes);

/* ---- encode: generate_probmaps (codec.py:11-70) + ProbMap.encode /
 *      ArgMaxProbMap.encode flags (codec.py:138-212, 443-513) ------------- */
typedef struct pp_encode_params {
  int32_t B, K, H, W;
  int32_t heatmap_dtype;   /* PP_F32 | PP_BF16: dtype of the maps written */
  int32_t keypoint_dtype;  /* PP_F32 | PP_F64: dtype of `keypoints`; the division by the codec
                              scale factor is done in that dtype, as NumPy does (codec.py:180) */
  int32_t keypoint_dim;    /* trailing dimension D of keypoints (>= 2) */
  float scale_x, scale_y;  /* codec scale_factor = (input-1)/(heatmap-1), float32 (codec.py:131-133);
                              1.0f when keypoints are already in heatmap space */
  float input_w, input_h;  /* for in_image (codec.py:189-200) */
} pp_encode_params;

PP_API int pp_encode(const pp_encode_params* p,
              const void* keypoints,     /* (B, K, D) input-image space */
              const float* visible,      /* (B, K) keypoints_visible; NULL = all ones */
              const double* two_s,       /* (K) 2*s per keypoint: the divisor of codec.py:65 */
              void* heatmaps,            /* out (B, K, H, W) */
              float* keypoint_weights,   /* out (B, K) or NULL (codec.py:46,68) */
              uint8_t* in_image,         /* out (B, K) or NULL */
              uint8_t* annotated,        /* out (B, K) or NULL (codec.py:187) */
              pp_stream_t stream);

/* ---- per-codec constant table for the expected-OKS decoder:
 *      _prepare_oks_kernels (heatmap.py:170-194), built once on the host ---- */
typedef struct pp_oks_table {
  const int32_t* radius;    /* (K) ceil(3 s) in [1, PP_MAX_OKS_RADIUS] */
  const float* taps_f32;    /* (K, PP_OKS_TAPS) normalised 1-D taps, tap j at column j (j < 2r+1) */
  const double* kernel2d;   /* (K, PP_OKS_TAPS*PP_OKS_TAPS) the reference's normalised d x d kernel,
                               packed row-major with row stride d = 2r+1 */
  const int32_t* order;     /* (K) channels sorted by decreasing radius (longest jobs first in the kernel's
                               work queue), or NULL for 0..K-1 */
  /* Optional (both NULL = not available): the 1-D taps of every distinct channel folded with the 'reflect' boundary
   * into banded Toeplitz matrices, float16, in the register-fragment order of mma.sync.m16n8k16 -- the operand tables
   * of the tensor-core prefilter (csrc/pp_decode_mma.cuh).  Built on the device by pp_oks_mma_table_build for one
   * (H, W); channels with identical taps share a table through mma_index. */
  const void* mma_tables;   /* pp_oks_mma_table_bytes(U, H, W) bytes, 16-byte aligned */
  const int32_t* mma_index; /* (K) table of channel k, in [0, U) */
  int32_t mma_H, mma_W;     /* the heatmap shape the tables were built for (they are ignored for any other shape) */
} pp_oks_table;

/* Bytes of the tensor-core prefilter tables for U distinct channels of an H x W heatmap; 0 when the shape has no
 * tensor-core kernel (then leave pp_oks_table.mma_tables NULL). */
PP_API int64_t pp_oks_mma_table_bytes(int32_t U, int32_t H, int32_t W);
/* Fill `out` (device) from the U distinct channels' taps_f32 rows (U, PP_OKS_TAPS) and radii (U), device pointers. */
PP_API int pp_oks_mma_table_build(const float* taps_f32, const int32_t* radius, int32_t U, int32_t H, int32_t W,
                                  void* out, pp_stream_t stream);

typedef struct pp_decode_params {
  int32_t B, K, H, W;
  int32_t heatmap_dtype;   /* PP_F32 | PP_BF16 */
  int32_t apply_tail;      /* 1: decode clamp(x / temperature, 0, 1) (head.py:526-532) instead of x */
  float temperature;
  double input_w, input_h; /* keypoints = locs / [W-1, H-1] * input_size (codec.py:237, 541) */
} pp_decode_params;

/* get_heatmap_expected_value (heatmap.py:291-395) + ProbMap.decode scaling (codec.py:214-239).
 * conv_out != NULL additionally returns the OKS-convolved maps (return_heatmap=True).
 * scratch: pp_decode_expected_scratch_bytes() bytes of device memory (contents undefined) that hold the
 * kernel's work-queue counter; NULL is allowed (the heatmaps are then split statically, ~25 % slower on
 * mixed inputs). */
PP_API int64_t pp_decode_expected_scratch_bytes(void);
/* Scratch that additionally lets pp_decode_expected use the tensor-core prefilter kernel: 16 bytes + one int32 per
 * heatmap (the list of heatmaps -- exact plateaus, maps without float32 dynamic range -- that it hands on to the
 * general kernel).  With less scratch than this the general kernels run. */
PP_API int64_t pp_decode_expected_scratch_bytes_for(const pp_decode_params* p);
PP_API int pp_decode_expected(const pp_decode_params* p, const pp_oks_table* table,
                       const void* heatmaps,
                       float* locs,        /* out (N, 2) sub-pixel argmax, heatmap px */
                       float* vals,        /* out (N) unconvolved map at the integer argmax */
                       int32_t* argmax,    /* out (N) flat index y*W+x of the convolved maximum, or NULL */
                       double* keypoints,  /* out (N, 2) input-space coordinates, or NULL */
                       float* conv_out,    /* out (N, H, W) or NULL */
                       void* scratch, int64_t scratch_bytes,
                       pp_stream_t stream);

/* Which kernel this thread's most recent pp_decode_expected launched (-1 before the first call).  The choice
 * depends on shape, alignment and batch size (heatmap.py:291-395 has a single code path; this is for benchmarks
 * and profiles, which must name the kernel they time). */
#define PP_DECODE_KERNEL_EXACT 0    /* full-map double-precision convolution (return_heatmap / maps beyond shared memory) */
#define PP_DECODE_KERNEL_CTA 1      /* one CTA per heatmap, pruned prefilter (pp_decode_fast.cuh) */
#define PP_DECODE_KERNEL_TEAM 2     /* one warp (team) per heatmap, band-wise prefilter (pp_decode_warp.cuh) */
#define PP_DECODE_KERNEL_DENSE 3    /* unpruned per-radius variant, PP_DECODE_DENSE=1 (pp_decode_dense.cuh) */
#define PP_DECODE_KERNEL_GENERIC 4  /* unaligned / odd shapes */
#define PP_DECODE_KERNEL_MMA 5      /* one warp per heatmap, tensor-core (mma.sync) Toeplitz prefilter (pp_decode_mma.cuh) */
PP_API int pp_decode_expected_last_kernel(void);

/* Number of floats of `conv_out` work space pp_decode_expected needs for this shape even when the
 * caller does not want the convolved maps (maps too large for the shared-memory kernel); 0 otherwise. */
PP_API int64_t pp_decode_expected_workspace_floats(const pp_decode_params* p);

/* get_heatmap_maximum (heatmap.py:13-52) */
PP_API int pp_heatmap_maximum(const void* heatmaps, int32_t heatmap_dtype, int64_t N, int32_t H, int32_t W,
                       float* locs,       /* out (N, 2); (-1,-1) where max <= 0 */
                       float* vals,       /* out (N) */
                       int32_t* argmax,   /* out (N) or NULL */
                       pp_stream_t stream);

/* ArgMaxProbMap.decode (codec.py:515-543): argmax + gaussian_blur (codec.py:284-313) +
 * refine_keypoints_dark_udp (codec.py:315-375) + scaling. */
PP_API int pp_decode_argmax_dark(const pp_decode_params* p,
                          const float* blur_taps, int32_t blur_ksize, /* (ksize) float32 taps, odd ksize */
                          const void* blur_mma_table, /* pp_blur_mma_table_build for this (H, W, ksize) -- enables the
                                                         tensor-core kernel (then scratch must hold
                                                         pp_decode_expected_scratch_bytes_for(p) bytes) -- or NULL */
                          const void* heatmaps,
                          float* peaks,       /* out (N, 2) integer peaks (or -1) or NULL */
                          float* scores,      /* out (N) raw maxima */
                          float* refined,     /* out (N, 2) refined heatmap-space coordinates */
                          double* keypoints,  /* out (N, 2) input-space coordinates, or NULL */
                          void* scratch, int64_t scratch_bytes, /* as for pp_decode_expected (may be NULL) */
                          pp_stream_t stream);

/* Operand table of the DARK decoder's tensor-core kernel: the zero-padded ksize-tap blur (codec.py:303-310) as banded
 * Toeplitz matrices in mma.sync fragment order, pp_oks_mma_table_bytes(1, H, W) bytes; ksize <= 15. */
PP_API int pp_blur_mma_table_build(const float* blur_taps, int32_t blur_ksize, int32_t H, int32_t W, void* out, pp_stream_t stream);
/* Which kernel this thread's most recent pp_decode_argmax_dark launched: 5 tensor-core, 1 CTA per heatmap, 0 other. */
PP_API int pp_decode_argmax_dark_last_kernel(void);

/* head tail (head.py:526-532, normalize=None): y = clamp(x / temperature, 0, 1) */
PP_API int pp_heatmap_tail(const void* x, void* y, int32_t dtype, int64_t numel, float temperature, pp_stream_t stream);

/* backward of the head tail: grad_x = grad_y / temperature where 0 <= x / temperature <= 1, else 0
 * (what autograd derives for torch.clamp(x / t, 0, 1), head.py:528-531) */
PP_API int pp_heatmap_tail_backward(const void* x, const void* grad_y, void* grad_x, int32_t dtype, int64_t numel,
                                    float temperature, pp_stream_t stream);

/* Sparsemax-normalised head tail (head.py:237-245, 526-532 with normalize != None; the projection itself is the
 * PyPI package sparsemax==0.1.9, `Sparsemax(dim=-1)` over the H*W pixels of each heatmap):
 *   y = clamp(sparsemax(x / temperature) * normalize, 0, 1).
 * aux (n_heatmaps, 2) float32 receives (max, tau) per heatmap for the backward; may be NULL at inference. */
PP_API int pp_sparsemax_tail(const void* x, void* y, float* aux, int32_t dtype, int64_t n_heatmaps, int64_t hw,
                             float temperature, float normalize, pp_stream_t stream);

/* backward of pp_sparsemax_tail: through the clamp (inclusive), the scale, the projection
 * (g_z = [p != 0] (g_p - mean over the support of g_p)) and the division by the temperature */
PP_API int pp_sparsemax_tail_backward(const void* x, const void* grad_y, const float* aux, void* grad_x, int32_t dtype,
                                      int64_t n_heatmaps, int64_t hw, float temperature, float normalize,
                                      pp_stream_t stream);

/* ---- training targets from decoded keypoints: ProbPoseLoss._oks_from_heatmaps (loss.py:550-640, with
 *      compute_oks(use_area=False, per_kpt=True), loss.py:715-764) and _error_from_heatmaps (loss.py:512-548).
 *      gt/dt keypoints are the (B, K, 2) float64 outputs of pp_decode_argmax_dark on the target / predicted maps. */
PP_API int pp_pose_targets(const double* gt_keypoints, const double* dt_keypoints,
                           const float* weight,    /* (B, K) in_image & annotated (loss.py:394) */
                           const double* sigmas,   /* (K) */
                           int32_t B, int32_t K, double heatmap_w, double heatmap_h,
                           float* oks,             /* out (B, K) or NULL */
                           float* oks_weight,      /* out (B) or NULL */
                           double* error,          /* out (B, K) Euclidean error or NULL */
                           pp_stream_t stream);

/* ---- OKSHeatmapLoss.forward (loss.py:55-143) and its backward ---------- */
typedef enum pp_loss_mode {
  PP_LOSS_PIXEL_MEAN = 0, /* mean over B*K*H*W of the per-pixel loss: what ProbPoseLoss uses (loss.py:428-431) */
  PP_LOSS_PER_PIXEL = 1,  /* (B, K, H, W) map (loss.py:122-127) */
  PP_LOSS_PER_KEYPOINT = 2 /* (B, K) (loss.py:128-134); the default mode is its mean (loss.py:135-141) */
} pp_loss_mode;

typedef struct pp_loss_params {
  int32_t B, K, H, W;
  int32_t dtype;               /* PP_F32 | PP_BF16: output, target, pixel weights, mask, loss map, grad */
  int32_t mode;                /* pp_loss_mode */
  int32_t oks_type;            /* 0 minus, 1 plus, 2 both (loss.py:92-99) */
  int32_t skip_empty_channel;  /* loss.py:180-189 */
  double smoothing_weight, gaussian_weight, loss_weight; /* the reference's Python floats */
  int64_t mask_stride_b, mask_stride_k; /* element strides of `mask` over B and K (0 = broadcast) */
} pp_loss_params;

/* bytes of caller-provided scratch needed by the loss entry points */
PP_API int64_t pp_oks_loss_scratch_bytes(const pp_loss_params* p);

/* Forward.  Outputs by mode:
 *   PIXEL_MEAN   : loss_scalar[0] (float32).  If grad != NULL the backward is fused into the same
 *                  pass: grad = d loss / d output * grad_scale (grad_scale is a host float, 1.0f for
 *                  a plain backward()).
 *   PER_PIXEL    : loss_map (B,K,H,W) in `dtype`.
 *   PER_KEYPOINT : loss_kpt (B,K) float32, loss_scalar[0] = mean(loss_kpt), peak_index (B,K) int32 =
 *                  flat index of the maximum masked Sobel energy (needed by the backward).
 * target_out_of_range (device int32, or NULL) is set to 1 when some target value is outside [0, 1]
 * -- the condition the reference asserts on (loss.py:85-86) -- and to 0 otherwise, in the same pass.
 */
PP_API int pp_oks_loss_forward(const pp_loss_params* p,
                        const void* output, const void* target,
                        const float* keypoint_weights, /* (B,K) or NULL */
                        const void* pixel_weights,     /* (B,K,H,W) in `dtype` or NULL */
                        const void* mask,              /* strided (B,K|1,H,W) in `dtype` or NULL */
                        void* loss_map, float* loss_kpt, float* loss_scalar, int32_t* peak_index,
                        void* grad, float grad_scale,
                        int32_t* target_out_of_range,
                        void* scratch, int64_t scratch_bytes,
                        const struct pp_mailbox* publish, /* or NULL; PIXEL_MEAN only: the kernel that finishes the loss also
                                                             stores it into the mailbox slot (the loss party, see
                                                             pp_mailbox_commit) */
                        pp_stream_t stream);

/* Encode-inside-loss: pp_encode (generate_probmaps, codec.py:11-70 + the flags of codec.py:187-200) followed by
 * pp_oks_loss_forward(PP_LOSS_PIXEL_MEAN) with the fused gradient (loss.py:428-431), as ONE pass that never
 * materialises the target: the kernel forms target[y][x] from the keypoint's separable float64 factors.  Per heatmap
 * it reads `output` and writes `grad`: 2 H W e bytes instead of 1 (encode) + 3 (loss).
 *   ep               : the encoder's parameters for the same (B, K, H, W); ep->heatmap_dtype is ignored (= p->dtype)
 *   keypoints/visible/two_s : as for pp_encode
 *   keypoint_weights : (B, K) weights of the loss, or NULL = the weights the encoder would return (codec.py:46,68)
 *   weights_out, in_image, annotated : the encoder's (B, K) outputs, each may be NULL
 * Shapes the fused kernel does not cover return PP_ERR_UNSUPPORTED_SHAPE (use pp_encode + pp_oks_loss_forward). */
PP_API int pp_oks_loss_forward_encoded(const pp_loss_params* p, const pp_encode_params* ep, const void* output,
                                       const void* keypoints, const float* visible, const double* two_s,
                                       const float* keypoint_weights, float* loss_scalar, void* grad, float grad_scale,
                                       float* weights_out, uint8_t* in_image, uint8_t* annotated, void* scratch,
                                       int64_t scratch_bytes, const struct pp_mailbox* publish, pp_stream_t stream);

typedef enum pp_upstream_kind {
  PP_UPSTREAM_SCALAR = 0, /* one float32 on the device, broadcast over the forward's output
                             (d L / d loss for PIXEL_MEAN; an expanded gradient for the other modes) */
  PP_UPSTREAM_FULL = 1    /* PER_PIXEL: (B,K,H,W) in `dtype`; PER_KEYPOINT: (B,K) float32 */
} pp_upstream_kind;

/* Backward for an arbitrary upstream gradient; peak_index (from the forward) is required for
 * PER_KEYPOINT. */
PP_API int pp_oks_loss_backward(const pp_loss_params* p,
                         const void* output, const void* target,
                         const float* keypoint_weights, const void* pixel_weights, const void* mask,
                         const void* upstream, int32_t upstream_kind, const int32_t* peak_index,
                         void* grad,
                         void* scratch, int64_t scratch_bytes, pp_stream_t stream);

/* grad *= scale[0] unless scale[0] == 1 (device scalar): lets the fused PIXEL_MEAN gradient be
 * reused by autograd without a host sync; a no-op launch in the usual case. */
PP_API int pp_scale_inplace(void* data, int32_t dtype, int64_t numel, const float* scale_dev, pp_stream_t stream);

/* ---- validation metrics on (N, K)-sized arrays; one small launch each ---- */

/* keypoint_pck_accuracy (loss.py:825-866) over _calc_distances / _distance_acc (heatmap.py:55-111).
 * pred / gt: (N, K, 2) float32 coordinates (pp_heatmap_maximum's `locs`); mask: (N, K) bytes;
 * norm_factor: (N, 2) float32 or float64 (norm_dtype = PP_F32 | PP_F64; the arithmetic runs in that type,
 * as NumPy's promotion does).  Out: acc (K) float64 (-1 = no valid instance), avg_acc, cnt; optionally the
 * (K, N) float32 distances (-1 = masked). */
PP_API int pp_pck_accuracy(const float* pred, const float* gt, const uint8_t* mask, const void* norm_factor,
                           int32_t norm_dtype, int32_t N, int32_t K, double thr, double* acc, double* avg_acc,
                           int32_t* cnt, float* distances /* or NULL */, pp_stream_t stream);

/* ---- decoded-keypoint records and their exchange between GPUs ("final keypoint gather" + loss) ----
 * Tail of Codec.decode (codec.py:249-263): one (x, y, score, probability, visibility, oks, error / diagonal) float64
 * record per keypoint.  With a mailbox the same kernel also stores the records and the step's local loss into every
 * rank's mailbox over NVLink peer memory; pp_mailbox_commit adds the loss and raises a per-source flag: the
 * all-gather of a multi-GPU loop without a collective call.  The mailbox is a symmetric allocation (same size on every rank, mapped
 * into every rank's address space, e.g. torch.distributed._symmetric_memory) of slots * world blocks of
 * pp_mailbox_block_bytes(N) bytes: [N * 7 doubles | pad to 16 | loss (double) | flag (uint32) | pad]; the loss sits
 * 16 bytes before the end of the block. */
typedef struct pp_mailbox {
  void* const* peer_bufs;   /* device array of `world` pointers: base of every rank's mailbox (own rank included) */
  uint32_t* state;          /* local device memory, pp_mailbox_state_words(slots) words, zero before the first call:
                               sequence number, arrival counter, finished-block counter and consumed sequence number of
                               each slot, then a status word (1 + rank of a consumer whose acknowledgement did not
                               arrive in time) */
  int32_t world, rank;
  int32_t slots, slot;      /* the block (slot, rank) of every mailbox is written */
  int64_t block_bytes;      /* pp_mailbox_block_bytes(N) */
  int32_t flow_control;     /* 1: a slot is not rewritten before every rank has acknowledged its previous publication
                               (pp_mailbox_ack); every rank must then consume every publication */
  int32_t reserved;
} pp_mailbox;

PP_API int64_t pp_mailbox_block_bytes(int64_t n_records);
/* size of one rank's mailbox (slots x world blocks + slots x world acknowledgement words) and of its local state */
PP_API int64_t pp_mailbox_bytes(int64_t n_records, int32_t world, int32_t slots);
PP_API int64_t pp_mailbox_state_words(int32_t slots);
PP_API int pp_pack_records(int64_t N,
                           const double* keypoints,      /* (N, 2) input-space coordinates (pp_decode_expected) */
                           const float* scores,          /* (N) */
                           const float* probabilities, const float* visibilities, const float* oks,
                           const float* errors,          /* (N) each: the four scalar heads */
                           float inv_diagonal,           /* 1 / sqrt(H^2 + W^2) in float32 (codec.py:261) */
                           double* records,              /* out (N, 7) local, or NULL when only the mailbox is wanted */
                           const pp_mailbox* mailbox,    /* or NULL: local records only */
                           pp_stream_t stream);
/* The loss party of the publication of `mailbox->slot`: stores the step's local loss (device scalar, or NULL for 0) into
 * the block on every rank.  A slot is published (its flag raised on every rank) by whichever of its two parties --
 * pp_pack_records and this call, or pp_oks_loss_forward[_encoded] with `publish` -- finishes last; they may run on
 * different streams without an ordering between them.  Exactly one of each per publication. */
PP_API int pp_mailbox_commit(const pp_mailbox* mailbox, int64_t n_records, const float* loss, pp_stream_t stream);
/* The same one step late, for a pipelined loop that keeps the NVLink round trips of the publication off the end of its
 * step: launched at the START of step j + 1 for the slot of step j (whose loss is final by then), it publishes only if
 * the slot's records party has arrived and is waiting, and does nothing otherwise (first step; slot already flushed). */
PP_API int pp_mailbox_commit_deferred(const pp_mailbox* mailbox, int64_t n_records, const float* loss, pp_stream_t stream);
/* Consumer side, after the blocks of `mailbox->slot` have been used: tell every producer that this rank is done with
 * sequence number `seq` of the slot (required with flow_control). */
PP_API int pp_mailbox_ack(const pp_mailbox* mailbox, uint32_t seq, pp_stream_t stream);
/* The whole consumer side in ONE kernel, with the sequence numbers kept on the device (state words 3 * slots ..): wait
 * (bounded, *status as for pp_mailbox_wait) for the oldest publication of `mailbox->slot` this rank has not consumed yet,
 * copy it into compact private buffers -- records_out (world * n_records * 7 doubles, rank order) and losses_out (world
 * doubles); both may be NULL to skip the copy -- and acknowledge it to every producer.  Returns at once, touching nothing,
 * when this rank has not published such a step itself.  No host-side values: capturable into the CUDA graph of the step,
 * e.g. step j publishes slot j % slots and consumes slot (j - 2) % slots. */
PP_API int pp_mailbox_consume(const pp_mailbox* mailbox, int64_t n_records, double* records_out, double* losses_out,
                              int64_t timeout_us, int32_t* status, pp_stream_t stream);
/* Wait until all `world` sources have published sequence number expected_seq into `slot` of this rank's mailbox.
 * *status (device int, zero before the call) becomes 1 + source rank if a source did not arrive within timeout_us and
 * -(1 + source rank) if a source has already published a LATER sequence number (the block was overwritten: only
 * possible without flow control). */
PP_API int pp_mailbox_wait(const void* local_mailbox, int32_t world, int32_t slot, int64_t n_records, uint32_t expected_seq,
                           int64_t timeout_us, int32_t* status, pp_stream_t stream);

/* ProbPoseLoss.get_binary_accuracy, force_balanced=False (loss.py:653-697): best accuracy over the given
 * thresholds among the mask-selected entries.  out = (best_acc, best_threshold); counts (n_thresholds + 1, the
 * last entry is the number of selected samples) may be NULL. */
PP_API int pp_binary_accuracy(const float* dt, const float* gt, const uint8_t* mask, int64_t n, const double* thresholds,
                              int32_t n_thresholds, float* out, int64_t* counts, pp_stream_t stream);

/* ProbPoseLoss.get_mae (loss.py:699-712): mean |dt - gt| over the mask-selected entries */
PP_API int pp_masked_mae(const float* dt, const float* gt, const uint8_t* mask, int64_t n, float* out, pp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PROBPOSE_B200_H_ */
