// ORB keypoint count at the reference's hard-wired 64x64 (complexity_metrics.py:385-387):
//   len(cv2.ORB_create().detectAndCompute(gray(resize(frame,(64,64)))))
// With edgeThreshold 31 only level 0 survives the border filter and only x,y in {31,32} can hold a
// keypoint, so the count is FAST-9/16 (threshold 20) + 3x3 non-maximum suppression evaluated at
// those four pixels.  The kernel therefore resizes only the live 10x10 window
// (rows/cols 27..36) of the 64x64 image: 100 bilinear gathers per frame.
#include "vqa_common.cuh"

namespace vqa {

__device__ __forceinline__ void linear_tap64(int d, int sn, bool vertical, int &i0, int &i1, int &w0, int &w1)
{
    const double scale = (double)sn / 64.0;
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

__device__ __forceinline__ unsigned bilin64(int p00, int p01, int p10, int p11, int a0, int a1, int b0, int b1)
{
    int t0 = p00 * a0 + p01 * a1, t1 = p10 * a0 + p11 * a1;
    return (unsigned)((((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2);
}

// source pixel (x, y) of a frame handed over as yuv420p planes, converted as yuv.cu converts it
struct YuvSrc {
    const uint8_t *y, *u, *v;
    int sy, su, sv;
    size_t fy, fu, fv;
};
__device__ __forceinline__ void yuv_bgr_at(const YuvSrc &s, int frame, int x, int y, int bgr[3])
{
    int cb, cg, cr;
    yuv_chroma(s.u[frame * s.fu + (size_t)(y >> 1) * s.su + (x >> 1)], s.v[frame * s.fv + (size_t)(y >> 1) * s.sv + (x >> 1)], cb, cg, cr);
    uint8_t B, G, R;
    yuv_px(s.y[frame * s.fy + (size_t)y * s.sy + x], cb, cg, cr, B, G, R);
    bgr[0] = B; bgr[1] = G; bgr[2] = R;
}

template <bool YUV>
__global__ void __launch_bounds__(128)
k_orb64(const uint8_t *__restrict__ bgr, size_t frame_stride, YuvSrc ys, int h, int w, int *__restrict__ counts,
        int *__restrict__ dbg)
{
    __shared__ int win[10][10];
    __shared__ int sc[4][4];
    const int frame = blockIdx.x, t = threadIdx.x;
    if (t < 100) {
        const int wy = t / 10, wx = t - wy * 10;
        int x0, x1, a0, a1, y0, y1, b0, b1;
        linear_tap64(27 + wx, w, false, x0, x1, a0, a1);
        linear_tap64(27 + wy, h, true, y0, y1, b0, b1);
        int p00[3], p01[3], p10[3], p11[3];
        if (YUV) {
            yuv_bgr_at(ys, frame, x0, y0, p00);
            yuv_bgr_at(ys, frame, x1, y0, p01);
            yuv_bgr_at(ys, frame, x0, y1, p10);
            yuv_bgr_at(ys, frame, x1, y1, p11);
        } else {
            const uint8_t *src = bgr + (size_t)frame * frame_stride;
            const uint8_t *r0 = src + (size_t)y0 * w * 3, *r1 = src + (size_t)y1 * w * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                p00[ch] = r0[x0 * 3 + ch]; p01[ch] = r0[x1 * 3 + ch];
                p10[ch] = r1[x0 * 3 + ch]; p11[ch] = r1[x1 * 3 + ch];
            }
        }
        unsigned B = bilin64(p00[0], p01[0], p10[0], p11[0], a0, a1, b0, b1);
        unsigned G = bilin64(p00[1], p01[1], p10[1], p11[1], a0, a1, b0, b1);
        unsigned R = bilin64(p00[2], p01[2], p10[2], p11[2], a0, a1, b0, b1);
        win[wy][wx] = (int)gray_of(B, G, R);
    }
    __syncthreads();
    if (t < 16) {
        // FAST-9 on the 16-pixel Bresenham ring of radius 3 (cv2 cornerScore<16>)
        const int RX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
        const int RY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
        const int py = 3 + (t >> 2), pxx = 3 + (t & 3);
        const int v = win[py][pxx];
        // strength = max over the 16 arcs of 9 contiguous ring pixels of
        //   min(v - p)  (centre brighter than the whole arc)  and  min(p - v)  (centre darker).
        // Both differences are kept explicitly and only `min` is used inside an arc: the equivalent
        // max(mn, -mx) form was mis-folded by ptxas 12.9 for sm_100a into max(mn, mx) (negated
        // VIMNMX3 operand dropped; seen on hardware, see profiles/r01_notes.md).  The empty asm pins
        // the per-arc minima as plain register values before the outer max.
        int dn[16], up[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int p = win[py + RY[k]][pxx + RX[k]];
            dn[k] = v - p;
            up[k] = p - v;
        }
        int best = -512;
#pragma unroll
        for (int s = 0; s < 16; s++) {
            int a = dn[s], b = up[s];
#pragma unroll
            for (int k = 1; k < 9; k++) {
                a = min(a, dn[(s + k) & 15]);
                b = min(b, up[(s + k) & 15]);
            }
            asm volatile("" : "+r"(a), "+r"(b));
            best = max(best, max(a, b));
        }
        sc[t >> 2][t & 3] = best > 20 ? best - 1 : 0;
    }
    __syncthreads();
    if (t < 4) {
        const int y = 1 + (t >> 1), x = 1 + (t & 1), s = sc[y][x];
        bool ok = s > 0;
        for (int j = -1; j <= 1; j++)
            for (int i = -1; i <= 1; i++)
                if ((i || j) && sc[y + j][x + i] >= s) ok = false;
        unsigned m = __ballot_sync(0xfu, ok);
        if (t == 0) counts[frame] = __popc(m);
    }
    if (dbg && frame == 0) {
        if (t < 100) dbg[t] = win[t / 10][t % 10];
        if (t < 16) dbg[100 + t] = sc[t >> 2][t & 3];
    }
}

int run_orb64(vqa_ctx *c, const uint8_t *bgr, int n, int h, int w, size_t frame_stride, int *counts, int *dbg)
{
    VQA_BYTES(c, (double)n * 100 * 12);
    VQA_LAUNCH(c, k_orb64<false>, n, 128, 0, bgr, frame_stride, YuvSrc{}, h, w, counts, dbg);
    return VQA_OK;
}

int run_orb64_yuv(vqa_ctx *c, const uint8_t *const planes[3], const int stride[3], const size_t frame_stride[3], int n, int h,
                  int w, int *counts)
{
    const YuvSrc ys{planes[0], planes[1], planes[2], stride[0], stride[1], stride[2], frame_stride[0], frame_stride[1], frame_stride[2]};
    VQA_BYTES(c, (double)n * 100 * 12);
    VQA_LAUNCH(c, k_orb64<true>, n, 128, 0, (const uint8_t *)nullptr, (size_t)0, ys, h, w, counts, (int *)nullptr);
    return VQA_OK;
}

}  // namespace vqa
