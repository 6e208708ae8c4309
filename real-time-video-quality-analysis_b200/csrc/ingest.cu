// Ingest stage: BGR -> gray (15-bit fixed point), bit-exact INTER_LINEAR resize, the four 256-bin
// histograms (B, G, R, gray) in warp-private shared-memory copies, and the entropies.
//
// Replaces the cv2.cvtColor / cv2.resize / cv2.calcHist / np.log2 calls of
// complexity_metrics.py:327-328,358-359,386,404-414,430,455-473,490-493,530-531.
// Roofline: HBM.  Algorithmic bytes per analysed frame (identity resize): read 3*H*W, write H*W.
#include <stdlib.h>

#include "vqa_common.cuh"

namespace vqa {

constexpr int GH_THREADS = 256;
constexpr int GH_WARPS = GH_THREADS / 32;

// grid = (blocks_per_frame, n_frames).  Each thread converts groups of 16 pixels: three 128-bit
// loads of interleaved BGR, one 128-bit store of gray.
template <bool HIST>
__global__ void __launch_bounds__(GH_THREADS)
k_gray_hist(const uint8_t *__restrict__ bgr, size_t frame_stride, int P, uint8_t *__restrict__ gray,
            uint32_t *__restrict__ hist)
{
    __shared__ unsigned sh[HIST ? GH_WARPS * 1024 : 1];
    const int frame = blockIdx.y;
    const uint8_t *src = bgr + (size_t)frame * frame_stride;
    uint8_t *dst = gray + (size_t)frame * P;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned *wh = sh + (HIST ? warp * 1024 : 0);
    if (HIST) {
        for (int i = threadIdx.x; i < GH_WARPS * 1024; i += GH_THREADS) sh[i] = 0;
        __syncthreads();
    }
    const int ngroups = P / 16;
    const bool vec = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
    const int stride = gridDim.x * GH_THREADS;
    for (int base = blockIdx.x * GH_THREADS; base < ngroups; base += stride) {
        const int g = base + threadIdx.x;
        const bool valid = g < ngroups;
        unsigned w[12];
        if (valid) {
            if (vec) {
                const uint4 *p = reinterpret_cast<const uint4 *>(src + (size_t)g * 48);
                uint4 a = ld_stream(p), b = ld_stream(p + 1), c = ld_stream(p + 2);
                w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
                w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
                w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
            } else {
                const uint8_t *p = src + (size_t)g * 48;
#pragma unroll
                for (int i = 0; i < 12; i++)
                    w[i] = p[4 * i] | (p[4 * i + 1] << 8) | (p[4 * i + 2] << 16) | ((unsigned)p[4 * i + 3] << 24);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 12; i++) w[i] = 0;
        }
        unsigned out[4] = {0, 0, 0, 0};
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int px = 0; px < 16; px++) {
            const int o = px * 3;
            unsigned B = (w[o >> 2] >> ((o & 3) * 8)) & 255u;
            unsigned G = (w[(o + 1) >> 2] >> (((o + 1) & 3) * 8)) & 255u;
            unsigned R = (w[(o + 2) >> 2] >> (((o + 2) & 3) * 8)) & 255u;
            unsigned Y = gray_of(B, G, R);
            out[px >> 2] |= Y << ((px & 3) * 8);
            if (HIST) hist_add4(wh, B, G, R, Y, valid, vm, lane);
        }
        if (valid) {
            if (vec) {
                *reinterpret_cast<uint4 *>(dst + (size_t)g * 16) = make_uint4(out[0], out[1], out[2], out[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) dst[(size_t)g * 16 + i] = (uint8_t)(out[i >> 2] >> ((i & 3) * 8));
            }
        }
    }
    // tail pixels (P % 16) by the first warp of block 0
    if (blockIdx.x == 0 && warp == 0) {
        const int i = ngroups * 16 + lane;
        const bool valid = i < P;
        unsigned B = 0, G = 0, R = 0, Y = 0;
        if (valid) {
            B = src[(size_t)i * 3]; G = src[(size_t)i * 3 + 1]; R = src[(size_t)i * 3 + 2];
            Y = gray_of(B, G, R);
            dst[i] = (uint8_t)Y;
        }
        if (HIST) {
            hist_add_plain(wh, B, valid, lane);
            hist_add_plain(wh + 256, G, valid, lane);
            hist_add_plain(wh + 512, R, valid, lane);
            hist_add_plain(wh + 768, Y, valid, lane);
        }
    }
    if (HIST) {
        __syncthreads();
        uint32_t *gh = hist + (size_t)frame * 1024;
        for (int i = threadIdx.x; i < 1024; i += GH_THREADS) {
            unsigned s = 0;
#pragma unroll
            for (int k = 0; k < GH_WARPS; k++) s += sh[k * 1024 + i];
            if (s) atomicAdd(&gh[i], s);
        }
    }
}

// cv2.resize INTER_LINEAR taps for one destination index.  Horizontal taps clamp
// the fraction at the borders; vertical taps keep it and clip the row indices (resizeGeneric_).
__device__ __forceinline__ void linear_tap(int d, int sn, int dn, bool vertical, int &i0, int &i1, int &w0, int &w1)
{
    const double scale = (double)sn / (double)dn;
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int i = (int)floorf(f);
    float a = f - (float)i;
    if (!vertical) {
        if (i < 0) { i = 0; a = 0.f; }
        if (i >= sn - 1) { i = sn - 1; a = 0.f; }
    }
    i0 = clampi(i, 0, sn - 1);
    i1 = clampi(i + 1, 0, sn - 1);
    w1 = __float2int_rn(__fmul_rn(a, 2048.f));
    w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, a), 2048.f));
}

__device__ __forceinline__ unsigned bilin_u8(int p00, int p01, int p10, int p11, int a0, int a1, int b0, int b1)
{
    int t0 = p00 * a0 + p01 * a1, t1 = p10 * a0 + p11 * a1;
    return (unsigned)((((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2);
}

// resize(frame) -> gray, with B/G/R/gray histograms.  One thread per destination pixel.
__global__ void __launch_bounds__(256)
k_resize_bgr_gray_hist(const uint8_t *__restrict__ bgr, size_t frame_stride, int h, int w, int rw, int rh,
                       uint8_t *__restrict__ gray_small, uint32_t *__restrict__ hist)
{
    __shared__ unsigned sh[1024];
    const int frame = blockIdx.y, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += 256) sh[i] = 0;
    __syncthreads();
    const uint8_t *src = bgr + (size_t)frame * frame_stride;
    const int total = rw * rh;
    for (int base = blockIdx.x * 256; base < total; base += gridDim.x * 256) {
        const int idx = base + threadIdx.x;
        const bool valid = idx < total;
        unsigned B = 0, G = 0, R = 0, Y = 0;
        if (valid) {
            const int dy = idx / rw, dx = idx - dy * rw;
            int x0, x1, a0, a1, y0, y1, b0, b1;
            linear_tap(dx, w, rw, false, x0, x1, a0, a1);
            linear_tap(dy, h, rh, true, y0, y1, b0, b1);
            const uint8_t *r0 = src + (size_t)y0 * w * 3, *r1 = src + (size_t)y1 * w * 3;
            B = bilin_u8(r0[x0 * 3], r0[x1 * 3], r1[x0 * 3], r1[x1 * 3], a0, a1, b0, b1);
            G = bilin_u8(r0[x0 * 3 + 1], r0[x1 * 3 + 1], r1[x0 * 3 + 1], r1[x1 * 3 + 1], a0, a1, b0, b1);
            R = bilin_u8(r0[x0 * 3 + 2], r0[x1 * 3 + 2], r1[x0 * 3 + 2], r1[x1 * 3 + 2], a0, a1, b0, b1);
            Y = gray_of(B, G, R);
            gray_small[(size_t)frame * total + idx] = (uint8_t)Y;
        }
        hist_add_plain(sh, B, valid, lane);
        hist_add_plain(sh + 256, G, valid, lane);
        hist_add_plain(sh + 512, R, valid, lane);
        hist_add_plain(sh + 768, Y, valid, lane);
    }
    __syncthreads();
    uint32_t *gh = hist + (size_t)frame * 1024;
    for (int i = threadIdx.x; i < 1024; i += 256)
        if (sh[i]) atomicAdd(&gh[i], sh[i]);
}

// Generic bit-exact resize of cn interleaved uint8 channels; one thread per destination element.
__global__ void __launch_bounds__(256)
k_resize_u8(const uint8_t *__restrict__ src, size_t frame_stride, int h, int w, int cn, int rw, int rh,
            uint8_t *__restrict__ dst)
{
    const int frame = blockIdx.y;
    const uint8_t *s = src + (size_t)frame * frame_stride;
    const long total = (long)rw * rh * cn;
    for (long idx = (long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long)gridDim.x * 256) {
        const int c = (int)(idx % cn);
        const long pix = idx / cn;
        const int dy = (int)(pix / rw), dx = (int)(pix - (long)dy * rw);
        int x0, x1, a0, a1, y0, y1, b0, b1;
        linear_tap(dx, w, rw, false, x0, x1, a0, a1);
        linear_tap(dy, h, rh, true, y0, y1, b0, b1);
        const uint8_t *r0 = s + (size_t)y0 * w * cn, *r1 = s + (size_t)y1 * w * cn;
        dst[(size_t)frame * total + idx] =
            (uint8_t)bilin_u8(r0[x0 * cn + c], r0[x1 * cn + c], r1[x0 * cn + c], r1[x1 * cn + c], a0, a1, b0, b1);
    }
}

// Entropies from the histograms: one warp per frame.
//   gray:   -sum_{p>0} p log2 p            (complexity_metrics.py:412-414, float32 terms)
//   colour: -sum_c sum_k p log2(p + 1e-8)  (complexity_metrics.py:455-473), NaN on an empty histogram
__global__ void k_entropy(const uint32_t *__restrict__ hist, int n, float *__restrict__ hist_entropy,
                          float *__restrict__ color_entropy)
{
    const int frame = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (frame >= n) return;
    const uint32_t *h = hist + (size_t)frame * 1024;
    double ent[4];
    bool empty = false;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        unsigned long long tot = 0;
        for (int k = lane; k < 256; k += 32) tot += h[c * 256 + k];
        tot = warp_sum(tot);
        const float ftot = (float)tot;
        double acc = 0;
        for (int k = lane; k < 256; k += 32) {
            const float p = __fdiv_rn((float)h[c * 256 + k], ftot);
            if (c == 3) {
                if (p > 0.f) acc += (double)__fmul_rn(p, log2f(p));
            } else {
                acc += (double)__fmul_rn(p, log2f(__fadd_rn(p, 1e-8f)));
            }
        }
        ent[c] = warp_sum(acc);
        if (c < 3 && tot == 0) empty = true;
    }
    if (lane == 0) {
        hist_entropy[frame] = (float)(-ent[3]);
        color_entropy[frame] = empty ? __int_as_float(0x7fc00000) : (float)(-(ent[0] + ent[1] + ent[2]));
    }
}

// sum of squares of a uint8 plane: 128-bit loads + dp4a (4 u8*u8 products per instruction)
__global__ void __launch_bounds__(256)
k_sq_sum(const uint8_t *__restrict__ x, long per_frame, unsigned long long *__restrict__ out)
{
    const int frame = blockIdx.y;
    const uint8_t *p = x + (size_t)frame * per_frame;
    unsigned long long acc = 0;
    const long nvec = ((reinterpret_cast<uintptr_t>(p) & 15) == 0) ? per_frame / 16 : 0;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (long)gridDim.x * 256) {
        const uint4 v = ld_stream(reinterpret_cast<const uint4 *>(p) + i);
        unsigned s = __dp4a(v.x, v.x, 0u);
        s = __dp4a(v.y, v.y, s);
        s = __dp4a(v.z, v.z, s);
        s = __dp4a(v.w, v.w, s);
        acc += s;
    }
    for (long i = nvec * 16 + (long)blockIdx.x * 256 + threadIdx.x; i < per_frame; i += (long)gridDim.x * 256) {
        unsigned v = p[i];
        acc += v * v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&out[frame], acc);
}

// sum x and sum x^2 of a gray frame from its 256-bin histogram (bins of plane 3): what k_frame_sum (DCT mean) and
// k_sq_sum (Parseval check of the DCT energy) would re-read the frame for
__global__ void __launch_bounds__(256)
k_hist_moments(const uint32_t *__restrict__ hist, unsigned long long *__restrict__ sum, unsigned long long *__restrict__ sq)
{
    __shared__ unsigned long long r1[8], r2[8];
    const int frame = blockIdx.x, i = threadIdx.x;
    const unsigned long long cnt = hist[(size_t)frame * 1024 + 768 + i];
    unsigned long long s1 = warp_sum(cnt * (unsigned long long)i), s2 = warp_sum(cnt * (unsigned long long)(i * i));
    if ((i & 31) == 0) { r1[i >> 5] = s1; r2[i >> 5] = s2; }
    __syncthreads();
    if (i == 0) {
        s1 = s2 = 0;
        for (int k = 0; k < 8; k++) { s1 += r1[k]; s2 += r2[k]; }
        if (sum) sum[frame] = s1;
        if (sq) sq[frame] = s2;
    }
}

// ------------------------------------------------------------------------------- launchers
int run_hist_moments(vqa_ctx *c, const uint32_t *hist, int n, unsigned long long *sum, unsigned long long *sq_sum)
{
    if (n <= 0) return VQA_OK;
    VQA_LAUNCH(c, k_hist_moments, n, 256, 0, hist, sum, sq_sum);
    return VQA_OK;
}

int run_gray_hist(vqa_ctx *c, const uint8_t *bgr, int n, int h, int w, size_t frame_stride, uint8_t *gray,
                  uint32_t *hist)
{
    const int P = h * w;
    int bpf = cdiv(cdiv(P, 16), GH_THREADS * 8);
    if (bpf < 1) bpf = 1;
    dim3 grid(bpf, n);
    if (hist) {
        VQA_CUDA(c, cudaMemsetAsync(hist, 0, sizeof(uint32_t) * 1024 * (size_t)n, c->stream));
        VQA_BYTES(c, 4.0 * P * n);
        VQA_LAUNCH(c, k_gray_hist<true>, grid, GH_THREADS, 0, bgr, frame_stride, P, gray, hist);
    } else {
        VQA_BYTES(c, 4.0 * P * n);
        VQA_LAUNCH(c, k_gray_hist<false>, grid, GH_THREADS, 0, bgr, frame_stride, P, gray, hist);
    }
    return VQA_OK;
}

int run_resize_bgr_gray_hist(vqa_ctx *c, const uint8_t *bgr, int n, int h, int w, size_t frame_stride, int rw, int rh,
                             uint8_t *gray_small, uint32_t *hist)
{
    VQA_CUDA(c, cudaMemsetAsync(hist, 0, sizeof(uint32_t) * 1024 * (size_t)n, c->stream));
    int bpf = cdiv((long)rw * rh, 256 * 4);
    if (bpf < 1) bpf = 1;
    VQA_BYTES(c, 13.0 * rw * rh * n);
    VQA_LAUNCH(c, k_resize_bgr_gray_hist, dim3(bpf, n), 256, 0, bgr, frame_stride, h, w, rw, rh, gray_small, hist);
    return VQA_OK;
}

int run_resize_u8(vqa_ctx *c, const uint8_t *src, int n, int h, int w, int cn, size_t frame_stride, int rw, int rh,
                  uint8_t *dst)
{
    int bpf = cdiv((long)rw * rh * cn, 256 * 4);
    if (bpf < 1) bpf = 1;
    VQA_BYTES(c, 5.0 * rw * rh * cn * n);
    VQA_LAUNCH(c, k_resize_u8, dim3(bpf, n), 256, 0, src, frame_stride, h, w, cn, rw, rh, dst);
    return VQA_OK;
}

int run_entropy(vqa_ctx *c, const uint32_t *hist, int n, float *hist_entropy, float *color_entropy)
{
    VQA_BYTES(c, 4096.0 * n);
    VQA_LAUNCH(c, k_entropy, cdiv(n, 4), 128, 0, hist, n, hist_entropy, color_entropy);
    return VQA_OK;
}

int run_sq_sum(vqa_ctx *c, const uint8_t *x, int n, long per_frame, unsigned long long *out)
{
    VQA_CUDA(c, cudaMemsetAsync(out, 0, sizeof(unsigned long long) * (size_t)n, c->stream));
    int bpf = cdiv(per_frame, 256 * 16 * 4);
    if (bpf < 1) bpf = 1;
    VQA_BYTES(c, (double)per_frame * n);
    VQA_LAUNCH(c, k_sq_sum, dim3(bpf, n), 256, 0, x, per_frame, out);
    return VQA_OK;
}

}  // namespace vqa
