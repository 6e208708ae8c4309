// Full-reference PSNR + SSIM on 8-bit planes in ONE pass over the pair (replaces the
// `psnr=` / `ssim=` FFmpeg filter graphs of run_ffmpeg_metrics, video_processing.py:270-297).
//
// FFmpeg semantics (vf_psnr.c / vf_ssim.c): per-plane SSE over every pixel;
// SSIM from integer 4x4 block sums (s1, s2, ss, s12), 8x8 windows = 2x2 groups of blocks on a
// 4-px grid, ssim_c1 = 416, ssim_c2 = 235963, float per-window value, mean over
// (w/4-1)(h/4-1) windows.  Integer sums are exact; the per-window floats are accumulated in
// double.  Roofline: HBM, algorithmic bytes = 2 * plane bytes (each pixel of both planes read once).
#include "vqa_common.cuh"

namespace vqa {

constexpr int PB_X = 32, PB_Y = 8;

__device__ __forceinline__ void block_sums(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, int stride,
                                           bool vec, int &s1, int &s2, int &ss, int &s12, unsigned &sse)
{
    s1 = s2 = ss = s12 = 0;
    sse = 0;
#pragma unroll
    for (int y = 0; y < 4; y++) {
        unsigned wa, wb;
        if (vec) {
            wa = *reinterpret_cast<const unsigned *>(a + (size_t)y * stride);
            wb = *reinterpret_cast<const unsigned *>(b + (size_t)y * stride);
        } else {
            const uint8_t *pa = a + (size_t)y * stride, *pb = b + (size_t)y * stride;
            wa = pa[0] | (pa[1] << 8) | (pa[2] << 16) | ((unsigned)pa[3] << 24);
            wb = pb[0] | (pb[1] << 8) | (pb[2] << 16) | ((unsigned)pb[3] << 24);
        }
#pragma unroll
        for (int x = 0; x < 4; x++) {
            int p = (wa >> (8 * x)) & 255, q = (wb >> (8 * x)) & 255, d = p - q;
            s1 += p; s2 += q; ss += p * p + q * q; s12 += p * q;
            sse += (unsigned)(d * d);
        }
    }
}

// a = main (distorted), b = reference.  grid (ceil(bw/32), ceil(bh/8), n)
__global__ void __launch_bounds__(PB_X * PB_Y)
k_psnr_ssim(const uint8_t *__restrict__ A, const uint8_t *__restrict__ B, int h, int w, int stride,
            unsigned long long *__restrict__ sse_out, double *__restrict__ ssim_out)
{
    __shared__ int4 sums[PB_Y + 1][PB_X + 1];
    __shared__ double red_d[8];
    __shared__ unsigned long long red_u[8];
    const int frame = blockIdx.z;
    const uint8_t *a = A + (size_t)frame * h * stride, *b = B + (size_t)frame * h * stride;
    const int bw = w >> 2, bh = h >> 2;
    const int bx0 = blockIdx.x * PB_X, by0 = blockIdx.y * PB_Y;
    const bool vec = (stride % 4 == 0) && ((((uintptr_t)a | (uintptr_t)b) & 3) == 0);
    unsigned long long sse = 0;
    for (int i = threadIdx.x; i < (PB_Y + 1) * (PB_X + 1); i += PB_X * PB_Y) {
        const int ly = i / (PB_X + 1), lx = i - ly * (PB_X + 1);
        const int bx = bx0 + lx, by = by0 + ly;
        int4 s = make_int4(0, 0, 0, 0);
        if (bx < bw && by < bh) {
            unsigned e;
            const size_t o = (size_t)(by * 4) * stride + bx * 4;
            block_sums(a + o, b + o, stride, vec, s.x, s.y, s.z, s.w, e);
            if (lx < PB_X && ly < PB_Y) sse += e;
        }
        sums[ly][lx] = s;
    }
    // pixels outside the 4x4 grid (w % 4 columns, h % 4 rows) only enter the SSE
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        const int wr = w - bw * 4, hr = h - bh * 4;
        for (int i = threadIdx.x; i < wr * h; i += PB_X * PB_Y) {
            int y = i / wr, x = bw * 4 + (i - y * wr);
            int d = (int)a[(size_t)y * stride + x] - (int)b[(size_t)y * stride + x];
            sse += (unsigned)(d * d);
        }
        for (int i = threadIdx.x; i < hr * bw * 4; i += PB_X * PB_Y) {
            int y = bh * 4 + i / (bw * 4), x = i % (bw * 4);
            int d = (int)a[(size_t)y * stride + x] - (int)b[(size_t)y * stride + x];
            sse += (unsigned)(d * d);
        }
    }
    __syncthreads();
    double v = 0;
    {
        const int lx = threadIdx.x % PB_X, ly = threadIdx.x / PB_X;
        const int bx = bx0 + lx, by = by0 + ly;
        if (bx < bw - 1 && by < bh - 1) {
            int4 p = sums[ly][lx], q = sums[ly][lx + 1], r = sums[ly + 1][lx], t = sums[ly + 1][lx + 1];
            int s1 = p.x + q.x + r.x + t.x, s2 = p.y + q.y + r.y + t.y;
            int ss = p.z + q.z + r.z + t.z, s12 = p.w + q.w + r.w + t.w;
            int vars = ss * 64 - s1 * s1 - s2 * s2, covar = s12 * 64 - s1 * s2;
            float num = __fmul_rn((float)(2 * s1 * s2 + 416), (float)(2 * covar + 235963));
            float den = __fmul_rn((float)(s1 * s1 + s2 * s2 + 416), (float)(vars + 235963));
            v = (double)__fdiv_rn(num, den);
        }
    }
    v = warp_sum(v);
    sse = warp_sum(sse);
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red_d[warp] = v; red_u[warp] = sse; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sv = 0;
        unsigned long long su = 0;
        for (int i = 0; i < 8; i++) { sv += red_d[i]; su += red_u[i]; }
        if (sv != 0) atomicAdd(&ssim_out[frame], sv);
        if (su) atomicAdd(&sse_out[frame], su);
    }
}

int run_psnr_ssim_plane(vqa_ctx *c, const uint8_t *a, const uint8_t *b, int n, int h, int w, int stride,
                        unsigned long long *sse, double *ssim_sum)
{
    VQA_CUDA(c, cudaMemsetAsync(sse, 0, sizeof(unsigned long long) * (size_t)n, c->stream));
    VQA_CUDA(c, cudaMemsetAsync(ssim_sum, 0, sizeof(double) * (size_t)n, c->stream));
    const int bw = w >> 2, bh = h >> 2;
    dim3 grid(cdiv(bw > 0 ? bw : 1, PB_X), cdiv(bh > 0 ? bh : 1, PB_Y), n);
    VQA_BYTES(c, 2.0 * h * w * n);
    VQA_LAUNCH(c, k_psnr_ssim, grid, PB_X * PB_Y, 0, a, b, h, w, stride, sse, ssim_sum);
    return VQA_OK;
}

}  // namespace vqa
