/*
 * vqa_b200.h -- C ABI of the B200-native per-frame scene-complexity + PSNR/SSIM path.
 *
 * The reference (zaki699/Real-Time-Video-Quality-Analysis) has no FFI of its own: the hot path
 * sits behind Python module-level functions that call OpenCV / the FFmpeg CLI.  Each entry
 * point below names the reference interface it replaces (file:line in /root/reference); the
 * ctypes binding a maintainer would add is shown in INTEGRATION.md and shipped in
 * real-time-video-quality-analysis_b200/_native.py.
 *
 * Conventions: plain pointers and sizes only; return 0 on success or a negative VQA_E_* code
 * (never an exception); the caller owns every in/out buffer; a context owns its scratch arena
 * and CUDA stream; one context per GPU; a context is not thread-safe.  `on_device != 0` means
 * the input pointers are device pointers on the context's GPU, otherwise host pointers (the
 * library stages them through its own pinned ring).  Outputs are always host memory and are
 * valid when the call returns.
 */
#ifndef VQA_B200_H
#define VQA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQA_ABI_VERSION 3

#if defined(__GNUC__)
#define VQA_API __attribute__((visibility("default")))
#else
#define VQA_API
#endif

enum {
    VQA_OK = 0,
    VQA_E_INVALID = -1,   /* bad argument */
    VQA_E_CUDA = -2,      /* CUDA runtime / driver error: see vqa_last_error() */
    VQA_E_NOMEM = -3,     /* device or pinned allocation failed */
    VQA_E_UNSUPPORTED = -4
};

/* metrics_mask bits of vqa_cfg (one per reference operator) */
enum {
    VQA_M_HIST = 1 << 0,      /* process_histogram_frame        complexity_metrics.py:392-416 */
    VQA_M_COLOR = 1 << 1,     /* process_color_histogram_frame  complexity_metrics.py:418-475 */
    VQA_M_EDGE = 1 << 2,      /* process_edge_frame             complexity_metrics.py:477-504 */
    VQA_M_DCT = 1 << 3,       /* process_dct_frame              complexity_metrics.py:346-364 */
    VQA_M_ORB = 1 << 4,       /* process_orb_frame_for_parallel complexity_metrics.py:367-389 */
    VQA_M_MOTION = 1 << 5,    /* process_frame_complexity       complexity_metrics.py:313-343 */
    VQA_M_TDCT = 1 << 6,      /* process_temporal_dct_frame     complexity_metrics.py:543-579 */
    VQA_M_ALL = 0x7f
};

typedef struct vqa_ctx vqa_ctx;

typedef struct vqa_cfg {
    int32_t resize_width;     /* config.json resize_width  (video_processing.py:186) */
    int32_t resize_height;    /* config.json resize_height (video_processing.py:187) */
    uint32_t metrics_mask;    /* VQA_M_* */
    int32_t dct_impl;         /* 0 = auto (tcgen05 tensor-core contraction), 1 = force fp32 SIMT check kernel */
    int32_t orb_width;        /* ORB input size: 0 = the reference's hard-wired 64x64 (complexity_metrics.py:386); */
    int32_t orb_height;       /* otherwise gray(resize(frame, (orb_width, orb_height))) through the full ORB pipeline */
} vqa_cfg;

/* cv2.ORB_create(nfeatures, scaleFactor, nlevels, edgeThreshold, ..., fastThreshold) parameters
 * (complexity_metrics.py:385 uses the defaults: 500, 1.2, 8, 31, 20; HARRIS score, patch 31) */
typedef struct vqa_orb_cfg {
    int32_t nfeatures;
    int32_t nlevels;          /* 1..16 */
    int32_t edge_threshold;   /* >= 4 */
    int32_t fast_threshold;
    float scale_factor;       /* > 1 */
} vqa_orb_cfg;

/* One detected keypoint (cv2.KeyPoint fields that exist before orientation / descriptors). */
typedef struct vqa_keypoint {
    float x, y;               /* cv2.KeyPoint.pt: level coordinates * (float)pow(scale_factor, octave) */
    float size;               /* cv2.KeyPoint.size: patch size 31 * level scale */
    float angle;              /* cv2.KeyPoint.angle: intensity-centroid orientation in degrees (orb.cpp ICAngles, fastAtan2) */
    float response;           /* Harris response (orb.cpp HarrisResponses), bit-exact */
    int32_t octave;           /* pyramid level */
    int32_t lx, ly;           /* integer coordinates inside the level image */
    int32_t fast_score;       /* FAST-9/16 corner score (strength - 1) */
} vqa_keypoint;

/* One row per analysed frame.  Field <- reference return value. */
typedef struct vqa_frame_metrics {
    float hist_entropy;       /* process_histogram_frame        -> np.float32 */
    float color_entropy;      /* process_color_histogram_frame  -> np.float32 (NaN on empty hist) */
    float dct_energy;         /* process_dct_frame              -> np.float32 sum(dct^2) */
    float motion;             /* process_frame_complexity((this, previous)); NaN if no previous */
    float temporal_dct;       /* process_temporal_dct_frame(previous, this);  NaN if no previous */
    int32_t orb_count;        /* process_orb_frame_for_parallel -> int */
    int64_t edge_count;       /* process_edge_frame             -> np.int64 */
    uint64_t gray_sq_sum;     /* exact sum of x^2 of the DCT input (Parseval cross-check of dct_energy) */
} vqa_frame_metrics;

/* One row per frame pair of the full-reference half. */
typedef struct vqa_fr_metrics {
    uint64_t sse[3];          /* per-plane sum of squared differences (vf_psnr.c compute_images_mse) */
    double mse[3];            /* sse / (w_c*h_c) */
    double mse_avg;           /* area-weighted (2/3, 1/6, 1/6 for yuv420p) */
    double psnr[3];           /* 10 log10(255^2 / mse_c), +inf when mse == 0 */
    double psnr_avg;          /* the `psnr_avg:` field parsed at video_processing.py:160 */
    double ssim[3];           /* vf_ssim.c ssim_plane per plane */
    double ssim_all;          /* the `All:` field parsed at video_processing.py:166 */
} vqa_fr_metrics;

/* ---- lifetime ------------------------------------------------------------------------- */
VQA_API int vqa_abi_version(void);
VQA_API int vqa_init(int device, vqa_ctx **out);
VQA_API void vqa_destroy(vqa_ctx *ctx);
VQA_API const char *vqa_last_error(const vqa_ctx *ctx);       /* ctx may be NULL: last init error */
VQA_API int vqa_set_stream(vqa_ctx *ctx, void *cuda_stream);  /* borrow a cudaStream_t (e.g. torch's) */
VQA_API int vqa_sync(vqa_ctx *ctx);
VQA_API uint64_t vqa_kernel_launches(const vqa_ctx *ctx);     /* kernels launched by this context so far */
/* device time (ms, CUDA events on the context's stream) of the named stage accumulated since
 * the last vqa_reset_timers; stage in {"ingest","canny","dct","orb","motion","frscore","all"} */
VQA_API int vqa_stage_ms(vqa_ctx *ctx, const char *stage, double *ms, uint64_t *launches);
VQA_API int vqa_reset_timers(vqa_ctx *ctx, int enable);
/* per-kernel profile (bench.py roofline leg): when enabled every launch is bracketed by CUDA events
 * on the context's stream; the report has one line per kernel: "name launches total_ms
 * algorithmic_bytes algorithmic_flops" (the byte/flop model is stated in DESIGN.md). */
VQA_API int vqa_kernel_profile(vqa_ctx *ctx, int enable);
VQA_API int vqa_kernel_report(vqa_ctx *ctx, char *buf, size_t cap);

/* ---- a1-a8: per-frame and pair complexity metrics ----------------------------------------
 * Replaces the bodies of process_frame_complexity / process_dct_frame /
 * process_orb_frame_for_parallel / process_histogram_frame / process_color_histogram_frame /
 * process_edge_frame / process_temporal_dct_frame (complexity_metrics.py:313-504,543-579) as
 * driven by process_in_batches (:128-148) over the sampled frames of one clip.
 *   bgr          n frames, uint8 HWC in B,G,R order, `frame_stride` bytes apart
 *   halo         optional previous sampled frame (same h,w) or NULL: pair metrics of frame 0
 *                are computed against it (frame-range sharding)
 *   out          n rows, host memory
 */
VQA_API int vqa_complexity_frames(vqa_ctx *ctx, const uint8_t *bgr, int n, int h, int w, size_t frame_stride,
                          const uint8_t *halo, int on_device, const vqa_cfg *cfg, vqa_frame_metrics *out);

/* ---- a8 at any size: ORB keypoint detection and scoring -------------------
 * cv2.ORB_create(...).detectAndCompute(gray, None) up to the keypoint list, on n one-channel frames
 * (rows of w bytes, frames `frame_stride` bytes apart).  counts[i] = len(keypoints) of frame i;
 * level_counts (optional) is [n][16]; kps (optional) receives up to kp_cap keypoints per frame,
 * grouped by octave, unordered inside an octave.  cfg NULL = ORB_create() defaults.
 * vqa_orb_describe is host-only (no context): pyramid level sizes and per-level feature quotas;
 * returns the number of levels. */
VQA_API void vqa_orb_default_cfg(vqa_orb_cfg *cfg);
VQA_API int vqa_orb_describe(const vqa_orb_cfg *cfg, int h, int w, int32_t *level_w, int32_t *level_h, int32_t *quota);
VQA_API int vqa_orb_detect(vqa_ctx *ctx, const uint8_t *gray, int n, int h, int w, size_t frame_stride, int on_device,
                           const vqa_orb_cfg *cfg, int32_t *counts, int32_t *level_counts, vqa_keypoint *kps, int kp_cap);

/* ---- a13: PSNR + SSIM of yuv420p (or any 3-plane 8-bit) frame pairs -------------------------
 * Replaces the psnr= and ssim= filter graphs run_ffmpeg_metrics builds
 * (video_processing.py:270-297); `main` is the distorted input [0:v], `ref` the reference
 * [1:v].  Planes are dense stacks: plane c of frame i starts at plane[c] + i*plane_h[c]*stride[c].
 */
VQA_API int vqa_psnr_ssim_planar(vqa_ctx *ctx, const uint8_t *const main_planes[3], const uint8_t *const ref_planes[3],
                         const int32_t plane_w[3], const int32_t plane_h[3], const int32_t stride[3],
                         int n, int on_device, vqa_fr_metrics *out);

/* ---- a1-a8 + a13 for one clip with HOST buffers: what process_video_and_extract_metrics does per
 * clip (video_processing.py:216 run_ffmpeg_metrics, :242 calculate_average_scene_complexity), as one
 * call with one interleaved upload schedule -- the planes of the full-reference half ride the copy
 * stream between the complexity chunks, so their transfer hides behind the Farneback compute.
 * Results are identical to calling vqa_complexity_frames and vqa_psnr_ssim_planar in turn. */
VQA_API int vqa_analyze_clip(vqa_ctx *ctx, const uint8_t *bgr, int n, int h, int w, size_t frame_stride,
                             const vqa_cfg *cfg, vqa_frame_metrics *rows_out,
                             const uint8_t *const main_planes[3], const uint8_t *const ref_planes[3],
                             const int32_t plane_w[3], const int32_t plane_h[3], const int32_t stride[3],
                             int n_pairs, vqa_fr_metrics *fr_out);

/* ---- f4: both halves of one clip from ONE upload of yuv420p planes ----------------------------------
 * The reference scores the source against its encode (run_ffmpeg_metrics(input, encoded),
 * video_processing.py:216) and then analyses the ENCODED file (calculate_average_scene_complexity,
 * video_processing.py:242), whose BGR frames are what cv2.VideoCapture.read (complexity_metrics.py:38-111)
 * makes of the encode's yuv420p planes: libswscale's unscaled yuv420p -> bgr24 converter.  That conversion
 * is reproduced bit-exactly on the device, so the planes are uploaded once (3 bytes per pixel and frame pair
 * instead of 6) and feed both PSNR/SSIM and the seven complexity metrics.
 *   main_planes  Y, U, V stacks of the encoded clip ([0:v] of the filter graphs); n frames of h x w (both
 *                even); plane p of frame i at main_planes[p] + i * plane_h[p] * stride[p]
 *   ref_planes   the source clip's stacks ([1:v]) or NULL: complexity rows only, fr_out untouched
 *   halo_planes  optional Y, U, V of the previous sampled frame of the encoded clip (frame-range sharding)
 *   on_device    planes (all of them) are device pointers on the context's GPU
 */
VQA_API int vqa_analyze_clip_yuv420(vqa_ctx *ctx, const uint8_t *const main_planes[3], const uint8_t *const ref_planes[3],
                                    const int32_t stride[3], int n, int h, int w, const uint8_t *const halo_planes[3],
                                    int on_device, const vqa_cfg *cfg, vqa_frame_metrics *rows_out, vqa_fr_metrics *fr_out);

/* ---- a9/a10: framerate variation + EWM-smoothed mean ----------------------------------------
 * vqa_framerate_series: process_frame_interval_for_parallel over consecutive timestamps
 * (complexity_metrics.py:150-165, driven at :296-298): fps[k] = 1000/(t[k+1]-t[k]) or 0.
 * vqa_ewm_partial: np.mean(pd.Series(x).ewm(alpha, adjust=True).mean()) (:114-125, :301-310)
 * written as a shard-summable weighted sum: returns sum_i c_{offset+i} x_i where c are the
 * closed-form coefficients for a series of `total` elements.  Summing the
 * partials of all shards gives the reference's smoothed mean.
 */
VQA_API int vqa_framerate_series(vqa_ctx *ctx, const double *timestamps_ms, int n, double *fps_out /* n-1 */);
VQA_API int vqa_ewm_partial(vqa_ctx *ctx, const double *x, int n_local, int64_t offset, int64_t total,
                    double alpha, double *partial_out);

/* ---- e: multi-GPU close of a clip (frame-range / clip sharding, SURVEY.md 8 e) -----------------------
 * One process per GPU.  Each rank analyses its frame range (vqa_complexity_frames / vqa_analyze_clip_yuv420
 * with the previous rank's last frame as halo), forms per-clip partial weighted sums with vqa_ewm_partial,
 * and vqa_clip_reduce sums them over all ranks: one fused buffer [n_f64 doubles | n_i64 integers], ONE
 * ncclAllReduce(sum) on the context's stream, results in place on every rank.  This stands where the
 * reference gathers the per-frame lists of its process pool (complexity_metrics.py:128-148) and takes
 * np.mean at :301-310.  Integers are carried exactly (the call fails rather than round above 2^53).
 *   nccl_comm   an ncclComm_t the caller owns whose rank's device is the context's GPU, or NULL to use
 *               the communicator created by vqa_comm_init
 * vqa_comm_unique_id (rank 0; host-only) + vqa_comm_init (every rank, with the id rank 0 distributed
 * out of band) wrap ncclGetUniqueId / ncclCommInitRank for hosts without an NCCL binding of their own.
 * vqa_comm_halo_exchange: rank r sends `send` (its last frame, device memory) to r+1 and receives
 * rank r-1's into `recv` over NVLink; NULL on the open ends.  NCCL is dlopen'ed on first use.
 */
#define VQA_COMM_ID_BYTES 128
VQA_API int vqa_comm_unique_id(uint8_t *id_out /* VQA_COMM_ID_BYTES */);
VQA_API int vqa_comm_init(vqa_ctx *ctx, const uint8_t *id /* VQA_COMM_ID_BYTES */, int rank, int world);
VQA_API int vqa_comm_destroy(vqa_ctx *ctx);
VQA_API int vqa_clip_reduce(vqa_ctx *ctx, void *nccl_comm, double *partials, int n_f64, int64_t *ints, int n_i64);
VQA_API int vqa_comm_halo_exchange(vqa_ctx *ctx, void *nccl_comm, const uint8_t *send, uint8_t *recv, size_t bytes);

/* ---- debug / stage-level parity taps (tests only; device work, host results) ------------ */
VQA_API int vqa_debug_gray(vqa_ctx *ctx, const uint8_t *bgr, int h, int w, uint8_t *gray_out);
VQA_API int vqa_debug_resize(vqa_ctx *ctx, const uint8_t *src, int h, int w, int channels, int rw, int rh, uint8_t *dst);
VQA_API int vqa_debug_hist(vqa_ctx *ctx, const uint8_t *bgr, int h, int w, int rw, int rh, uint32_t *hist_out /* 4*256: B,G,R,gray */);
/* out117: 10x10 live window of the 64x64 gray, 4x4 FAST scores, keypoint count */
VQA_API int vqa_debug_orb(vqa_ctx *ctx, const uint8_t *bgr, int h, int w, int32_t *out117);
VQA_API int vqa_debug_canny(vqa_ctx *ctx, const uint8_t *gray, int h, int w, uint8_t *edges_out /* 0/255 */);
VQA_API int vqa_debug_flow(vqa_ctx *ctx, const uint8_t *prev_gray, const uint8_t *next_gray, int h, int w, float *flow_out /* h*w*2 */);
/* level `level` (>= 1) of the ORB pyramid of one gray frame (INTER_LINEAR_EXACT chain), dense rows */
VQA_API int vqa_debug_orb_pyramid(vqa_ctx *ctx, const uint8_t *gray, int h, int w, const vqa_orb_cfg *cfg, int level,
                                  uint8_t *level_out);
/* host-only: packed (offset << 16 | weight of the right tap, 8.8 fixed point) taps of INTER_LINEAR_EXACT */
VQA_API int vqa_debug_exact_taps(int src_len, int dst_len, uint32_t *taps_out);
VQA_API int vqa_debug_dct(vqa_ctx *ctx, const uint8_t *gray, int h, int w, int impl, float *coef_out /* h*w */);
/* the yuv420p -> BGR conversion of vqa_analyze_clip_yuv420 on one frame (h, w even; dense planes) */
VQA_API int vqa_debug_yuv2bgr(vqa_ctx *ctx, const uint8_t *y, const uint8_t *u, const uint8_t *v, int h, int w,
                              uint8_t *bgr_out /* h*w*3 */);

#ifdef __cplusplus
}
#endif
#endif /* VQA_B200_H */
