// Shared infrastructure of libvqa_b200.so: context, scratch arena, launch accounting, warp helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/vqa_b200.h"

namespace vqa {

struct Buf {
    void *p = nullptr;
    size_t cap = 0;
};

struct KRec {                                    // per-kernel-name profile (vqa_kernel_profile)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
    size_t used = 0;
    double bytes = 0;                            // algorithmic bytes declared with VQA_BYTES
    double flops = 0;
};

struct StageTimer {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
    size_t used = 0;
    double ms = 0;
    uint64_t launches = 0;
};

}  // namespace vqa

struct vqa_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t side_stream = nullptr;          // Canny / ORB / DCT chain of a chunk, concurrent with the Farneback chain
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t fb_stream = nullptr;            // second half of the pairs of a Farneback level, one kernel behind the first half
    cudaEvent_t ev_fb_fork = nullptr, ev_fb_stagger = nullptr, ev_fb_join = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    char err[512] = {0};
    uint64_t launches = 0;
    std::map<std::string, vqa::Buf> bufs;         // named grow-only device scratch
    std::map<std::string, vqa::Buf> pinned;       // named grow-only pinned host staging
    bool timing = false;
    std::map<std::string, vqa::StageTimer> timers;
    cudaEvent_t ev_copy[3] = {nullptr, nullptr, nullptr};   // H2D of staging slot s finished
    cudaEvent_t ev_done[3] = {nullptr, nullptr, nullptr};   // compute that read staging slot s finished
    cudaEvent_t ev_sync = nullptr;                          // blocking-sync event for end-of-call waits
    bool ktiming = false;                         // per-kernel CUDA-event timing (bench roofline leg)
    std::map<std::string, vqa::KRec> krec;
    double cur_bytes = 0, cur_flops = 0;          // algorithmic traffic of the NEXT launch
    // cudaMemGetInfo goes through the kernel driver's resource-manager lock (which monitoring agents
    // polling NVML hold for tens of ms at a time): asked once per allocation epoch, not once per call
    uint64_t alloc_epoch = 1, free_epoch = 0;
    size_t free_cached = 0;
    void *umma = nullptr;                         // tensor-map cache of the tcgen05 DCT (dct_umma.cu)
    void *orb = nullptr;                          // pyramid geometry cache of the general-size ORB (orb.cu)
    void *comm = nullptr;                         // ncclComm_t owned by the context (vqa_comm_init, comm.cu)
    int comm_rank = 0, comm_world = 0;
};

namespace vqa {

int set_err(vqa_ctx *c, int code, const char *fmt, ...);
size_t free_device_memory(vqa_ctx *c);
void *dev_buf(vqa_ctx *c, const char *name, size_t bytes);     // nullptr on failure (error set)
void *pinned_buf(vqa_ctx *c, const char *name, size_t bytes);
void stage_begin(vqa_ctx *c, const char *stage);
void stage_end(vqa_ctx *c, const char *stage);
cudaError_t wait_stream(vqa_ctx *c);                           // yielding wait for c->stream (no driver-lock spinning)

#define VQA_CUDA(c, call)                                                                       \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return vqa::set_err((c), VQA_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,  \
                                cudaGetErrorString(e__));                                       \
    } while (0)

struct KTimer {
    vqa_ctx *c;
    KRec *r = nullptr;
    KTimer(vqa_ctx *c_, const char *name) : c(c_)
    {
        if (!c->ktiming) { c->cur_bytes = c->cur_flops = 0; return; }
        r = &c->krec[name];
        if (r->used == r->ev.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            r->ev.push_back({a, b});
        }
        cudaEventRecord(r->ev[r->used].first, c->stream);
    }
    ~KTimer()
    {
        if (!r) return;
        cudaEventRecord(r->ev[r->used].second, c->stream);
        r->used++;
        r->bytes += c->cur_bytes;
        r->flops += c->cur_flops;
        c->cur_bytes = c->cur_flops = 0;
    }
};

// Declare the ALGORITHMIC bytes / flops of the next launch (DESIGN.md gives the per-unit model).
#define VQA_BYTES(c, b) ((c)->cur_bytes = (double)(b))
#define VQA_FLOPS(c, f) ((c)->cur_flops = (double)(f))

// Launch + account + check.  Every kernel of the library goes through this.
#define VQA_LAUNCH(c, kern, grid, block, smem, ...)                                             \
    do {                                                                                        \
        vqa::KTimer kt__((c), #kern);                                                           \
        kern<<<(grid), (block), (smem), (c)->stream>>>(__VA_ARGS__);                            \
        (c)->launches++;                                                                        \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess)                                                                 \
            return vqa::set_err((c), VQA_E_CUDA, "%s:%d launch %s -> %s", __FILE__, __LINE__,   \
                                #kern, cudaGetErrorString(e__));                                \
    } while (0)

#define VQA_BUF(c, var, type, name, count)                                                      \
    type *var = (type *)vqa::dev_buf((c), (name), sizeof(type) * (size_t)(count));              \
    if (!var) return VQA_E_NOMEM

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------- device helpers
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 128-bit streaming load (read-once data: bypass L1 allocation)
__device__ __forceinline__ uint4 ld_stream(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// OpenCV BGR->gray, 15-bit fixed point
__device__ __forceinline__ unsigned gray_of(unsigned b, unsigned g, unsigned r)
{
    return (3735u * b + 19235u * g + 9798u * r + (1u << 14)) >> 15;
}

// Histogram update of a warp-private shared-memory copy: plain shared-memory atomics, except that a warp whose lanes all
// hit the same bin (flat areas) issues one add.  On B200 this is ~10x faster than per-bin match.any aggregation for
// textured frames (profiles/r01_notes.md) and keeps the flat-frame worst case at 1 atomic.
__device__ __forceinline__ void hist_add_plain(unsigned *h, unsigned bin, bool valid, int lane)
{
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    const unsigned b0 = __shfl_sync(0xffffffffu, bin, __ffs(vm | 0x80000000u) - 1);
    if (__all_sync(0xffffffffu, !valid || bin == b0)) {
        if (vm && lane == __ffs(vm) - 1) atomicAdd(&h[bin], (unsigned)__popc(vm));
    } else if (valid) {
        atomicAdd(&h[bin], 1u);
    }
}

// The four histogram updates of one pixel (B, G, R, gray planes of a warp-private copy) behind ONE uniformity test: `vm` is the
// ballot of the lanes that carry a pixel (hoisted by the caller: it is the same for all pixels of a work item).  Three warp
// collectives per pixel instead of twelve (the per-plane test made the fused ingest kernel ALU-bound: SM 71 % busy).
__device__ __forceinline__ void hist_add4(unsigned *wh, unsigned B, unsigned G, unsigned R, unsigned Y, bool valid, unsigned vm, int lane)
{
    const unsigned key = B | (G << 8) | (R << 16) | (Y << 24);
    const int first = __ffs(vm | 0x80000000u) - 1;
    const unsigned k0 = __shfl_sync(0xffffffffu, key, first);
    if (__all_sync(0xffffffffu, !valid || key == k0)) {
        if (vm && lane == first) {
            const unsigned cnt = (unsigned)__popc(vm);
            atomicAdd(&wh[B], cnt);
            atomicAdd(&wh[256 + G], cnt);
            atomicAdd(&wh[512 + R], cnt);
            atomicAdd(&wh[768 + Y], cnt);
        }
    } else if (valid) {
        atomicAdd(&wh[B], 1u);
        atomicAdd(&wh[256 + G], 1u);
        atomicAdd(&wh[512 + R], 1u);
        atomicAdd(&wh[768 + Y], 1u);
    }
}

// libswscale's unscaled yuv420p -> bgr24 pixel (yuv.cu): luma term y, chroma terms of the 2x2 block
__device__ __forceinline__ void yuv_chroma(int U, int V, int &cb, int &cg, int &cr)
{
    const int u = (U << 3) - 1024, v = (V << 3) - 1024;
    cb = (u * 16525) >> 16;
    cg = ((u * -3209) >> 16) + ((v * -6660) >> 16);
    cr = (v * 13075) >> 16;
}
__device__ __forceinline__ void yuv_px(int Y, int cu_b, int cg, int cv_r, uint8_t &b, uint8_t &g, uint8_t &r)
{
    const int y = (((Y << 3) - 128) * 9539) >> 16;
    b = (uint8_t)min(max(y + cu_b, 0), 255);
    g = (uint8_t)min(max(y + cg, 0), 255);
    r = (uint8_t)min(max(y + cv_r, 0), 255);
}

// ---------------------------------------------------------------------------- stage launchers
// ingest.cu
int run_gray_hist(vqa_ctx *c, const uint8_t *bgr, int n, int h, int w, size_t frame_stride, uint8_t *gray,
                  uint32_t *hist /* [n][4][256] or nullptr */);
int run_resize_bgr_gray_hist(vqa_ctx *c, const uint8_t *bgr, int n, int h, int w, size_t frame_stride, int rw, int rh,
                             uint8_t *gray_small, uint32_t *hist);
int run_resize_u8(vqa_ctx *c, const uint8_t *src, int n, int h, int w, int cn, size_t frame_stride, int rw, int rh,
                  uint8_t *dst);
int run_entropy(vqa_ctx *c, const uint32_t *hist, int n, float *hist_entropy, float *color_entropy);
int run_sq_sum(vqa_ctx *c, const uint8_t *x, int n, long per_frame, unsigned long long *out);
int run_hist_moments(vqa_ctx *c, const uint32_t *hist /* [n][4][256] */, int n, unsigned long long *sum /* [n] */,
                     unsigned long long *sq_sum /* [n] */);
// canny.cu
int run_canny(vqa_ctx *c, const uint8_t *gray, int n, int h, int w, unsigned long long *counts /* [n] dev */,
              uint8_t *edges_out /* optional [n][h][w] 0/255 */);
// fast_orb.cu
int run_orb64(vqa_ctx *c, const uint8_t *bgr, int n, int h, int w, size_t frame_stride, int *counts /* [n] dev */,
              int *dbg = nullptr /* optional [116]: 10x10 window + 4x4 scores of frame 0 */);
int run_orb64_yuv(vqa_ctx *c, const uint8_t *const planes[3], const int stride[3], const size_t frame_stride[3], int n, int h,
                  int w, int *counts /* [n] dev */);
// orb.cu: general-size ORB (pyramid, FAST, Harris, retainBest); gray rows `pitch0` bytes apart
void orb_defaults(vqa_orb_cfg *cfg);
int run_orb_general(vqa_ctx *c, const uint8_t *gray, int n, int h, int w, size_t frame_stride, int pitch0,
                    const vqa_orb_cfg *cfg, int *counts /* [n] dev */, int *level_counts /* [n][16] dev or null */,
                    vqa_keypoint *kps /* [n][kp_cap] dev or null */, int kp_cap);
int orb_describe(const vqa_orb_cfg *cfg, int h, int w, int32_t *level_w, int32_t *level_h, int32_t *quota);
int orb_exact_taps(int sn, int dn, uint32_t *out);
int orb_level_view(vqa_ctx *c, int level, const uint8_t **ptr, int *pitch, int *lh, int *lw);   // frame 0 of the last call
void orb_release(vqa_ctx *c);
// dct.cu / dct_umma.cu
int run_dct(vqa_ctx *c, const uint8_t *x, int n, int h, int w, int impl, float *coef /* [n][h][w] dev */,
            double *energy /* [n] dev */, const unsigned long long *pixel_sums = nullptr /* [n] dev: sum of x per frame, if known */);
int run_abs_diff_sum(vqa_ctx *c, const float *a, const float *b, int n, long per_frame, size_t stride_a,
                     size_t stride_b, double *out /* [n] dev */);
int run_dct_umma(vqa_ctx *c, const uint8_t *x, int n, int h, int w, float *coef, double *energy,
                 const unsigned long long *pixel_sums = nullptr);
void dct_umma_release(vqa_ctx *c);
// farneback.cu
int run_farneback(vqa_ctx *c, const uint8_t *gray /* [n+1][h][w] */, int npairs, int h, int w,
                  double *mag_sum /* [npairs] dev: sum |flow| */, float *flow_out /* optional, level-0 flow of pair 0.. */,
                  const std::function<int()> *level0_hook = nullptr /* called once, before the first UpdateMatrices of level 0 */);
// psnr_ssim.cu
int run_psnr_ssim_plane(vqa_ctx *c, const uint8_t *a, const uint8_t *b, int n, int h, int w, int stride, size_t frame_stride,
                        unsigned long long *sse /* [n] */, double *ssim_sum /* [n] */);
int run_psnr_ssim_planes(vqa_ctx *c, const uint8_t *const a[3], const uint8_t *const b[3], int n, const int plane_h[3],
                         const int plane_w[3], const int stride[3], const size_t frame_stride[3],
                         unsigned long long *sse /* [3][out_pitch] */, double *ssim_sum /* [3][out_pitch] */, int out_pitch);
// yuv.cu
int run_yuv420_to_bgr(vqa_ctx *c, const uint8_t *const planes[3], const int stride[3], const size_t frame_stride[3],
                      int n, int h, int w, uint8_t *bgr /* [n][h][w][3] dense */);
bool yuv420_gray_hist_ok(const uint8_t *const planes[3], const int stride[3], const size_t frame_stride[3], int h, int w,
                         const uint8_t *gray);
int run_yuv420_gray_hist(vqa_ctx *c, const uint8_t *const planes[3], const int stride[3], const size_t frame_stride[3],
                         int n, int h, int w, uint8_t *gray /* [n][h][w] */, uint32_t *hist /* [n][4][256] or nullptr */);
// comm.cu
void comm_release(vqa_ctx *c);
// stats.cu
int run_framerate(vqa_ctx *c, const double *ts_dev, int n, double *fps_dev);
int run_ewm_partial(vqa_ctx *c, const double *x_dev, int n_local, long long offset, long long total, double alpha,
                    double *out_dev);

}  // namespace vqa
