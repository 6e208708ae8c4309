// yuv420p -> BGR exactly as the reference's frame reader produces it.
//
// calculate_average_scene_complexity decodes the ENCODED file with cv2.VideoCapture
// (complexity_metrics.py:38-111; called on the encode at video_processing.py:242), i.e. libavcodec
// yuv420p frames run through libswscale's unscaled yuv420p -> bgr24 converter.  That converter (the
// x86 SIMD path every build of cv2's bundled FFmpeg takes for even frame sizes) is a per-pixel integer
// function with nearest chroma (one U,V sample per 2x2 block):
//     y = ((Y << 3) - 128) * 9539 >> 16            (arithmetic shifts = pmulhw)
//     u = (U << 3) - 1024,  v = (V << 3) - 1024
//     B = sat8(y + (u * 16525 >> 16))
//     G = sat8(y + (u * -3209 >> 16) + (v * -6660 >> 16))
//     R = sat8(y + (v * 13075 >> 16))
// Pinned exhaustively (all 2^24 (Y,U,V) triples, and even sizes from 2x2 to 1080p) on
// cv2.VideoCapture reading yuv4mpeg files: oracle/make_golden.py --yuv2bgr, tests/golden/yuv2bgr_cv2.*.
// With it one upload of the yuv420p planes feeds both halves of the path (PSNR/SSIM on the planes, the
// seven complexity metrics on the derived BGR frames) instead of two uploads in two formats.
// Roofline: HBM, algorithmic bytes 1.5 HW read + 3 HW written.
#include "vqa_common.cuh"

namespace vqa {

// one thread: 8 pixels x 2 rows (4 chroma samples).  grid (ceil(w/8/64), h/2, n), block 64.
__global__ void __launch_bounds__(64)
k_yuv420_to_bgr(const uint8_t *__restrict__ Yp, const uint8_t *__restrict__ Up, const uint8_t *__restrict__ Vp,
                int h, int w, int sy, int su, int sv, size_t fy, size_t fu, size_t fv, uint8_t *__restrict__ bgr,
                int vec)
{
    const int frame = blockIdx.z, y0 = blockIdx.y * 2;
    const int x0 = (blockIdx.x * 64 + threadIdx.x) * 8;
    if (x0 >= w) return;
    const uint8_t *yr0 = Yp + frame * fy + (size_t)y0 * sy + x0, *yr1 = yr0 + sy;
    const uint8_t *ur = Up + frame * fu + (size_t)(y0 >> 1) * su + (x0 >> 1);
    const uint8_t *vr = Vp + frame * fv + (size_t)(y0 >> 1) * sv + (x0 >> 1);
    uint8_t *o0 = bgr + ((size_t)frame * h + y0) * (size_t)w * 3 + (size_t)x0 * 3, *o1 = o0 + (size_t)w * 3;
    uint8_t ya[8], yb[8], uu[4], vv[4];
    const int npx = min(8, w - x0);
    if (vec) {
        const uint2 a = __ldg(reinterpret_cast<const uint2 *>(yr0)), b = __ldg(reinterpret_cast<const uint2 *>(yr1));
        const unsigned u4 = __ldg(reinterpret_cast<const unsigned *>(ur)), v4 = __ldg(reinterpret_cast<const unsigned *>(vr));
#pragma unroll
        for (int j = 0; j < 4; j++) {
            ya[j] = (a.x >> (8 * j)) & 255; ya[4 + j] = (a.y >> (8 * j)) & 255;
            yb[j] = (b.x >> (8 * j)) & 255; yb[4 + j] = (b.y >> (8 * j)) & 255;
            uu[j] = (u4 >> (8 * j)) & 255; vv[j] = (v4 >> (8 * j)) & 255;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            ya[j] = j < npx ? yr0[j] : 0;
            yb[j] = j < npx ? yr1[j] : 0;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uu[j] = 2 * j < npx ? ur[j] : 0;
            vv[j] = 2 * j < npx ? vr[j] : 0;
        }
    }
    uint8_t r0[24], r1[24];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int cb, cg, cr;
        yuv_chroma(uu[j], vv[j], cb, cg, cr);
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int p = 2 * j + k;
            yuv_px(ya[p], cb, cg, cr, r0[3 * p], r0[3 * p + 1], r0[3 * p + 2]);
            yuv_px(yb[p], cb, cg, cr, r1[3 * p], r1[3 * p + 1], r1[3 * p + 2]);
        }
    }
    if (vec) {
        uint2 w0[3], w1[3];
#pragma unroll
        for (int q = 0; q < 3; q++) {
            w0[q].x = r0[8 * q] | (r0[8 * q + 1] << 8) | (r0[8 * q + 2] << 16) | ((unsigned)r0[8 * q + 3] << 24);
            w0[q].y = r0[8 * q + 4] | (r0[8 * q + 5] << 8) | (r0[8 * q + 6] << 16) | ((unsigned)r0[8 * q + 7] << 24);
            w1[q].x = r1[8 * q] | (r1[8 * q + 1] << 8) | (r1[8 * q + 2] << 16) | ((unsigned)r1[8 * q + 3] << 24);
            w1[q].y = r1[8 * q + 4] | (r1[8 * q + 5] << 8) | (r1[8 * q + 6] << 16) | ((unsigned)r1[8 * q + 7] << 24);
            reinterpret_cast<uint2 *>(o0)[q] = w0[q];
            reinterpret_cast<uint2 *>(o1)[q] = w1[q];
        }
    } else {
        for (int j = 0; j < 3 * npx; j++) { o0[j] = r0[j]; o1[j] = r1[j]; }
    }
}

// planes: dense stacks, frame f of plane p at plane[p] + f * frame_stride[p]; rows stride[p] bytes apart.
// h and w must be even (swscale takes another converter for odd sizes; yuv420p video never has them).
int run_yuv420_to_bgr(vqa_ctx *c, const uint8_t *const planes[3], const int stride[3], const size_t frame_stride[3],
                      int n, int h, int w, uint8_t *bgr)
{
    if (n <= 0) return VQA_OK;
    if ((h | w) & 1) return set_err(c, VQA_E_UNSUPPORTED, "yuv420p -> BGR needs even frame sizes (got %dx%d)", w, h);
    const bool vec = (w % 8 == 0) && (stride[0] % 8 == 0) && (stride[1] % 4 == 0) && (stride[2] % 4 == 0) &&
                     (frame_stride[0] % 8 == 0) && (frame_stride[1] % 4 == 0) && (frame_stride[2] % 4 == 0) &&
                     (((uintptr_t)planes[0] | (uintptr_t)bgr) % 8 == 0) && (((uintptr_t)planes[1] | (uintptr_t)planes[2]) % 4 == 0);
    VQA_BYTES(c, 4.5 * h * w * n);
    VQA_LAUNCH(c, k_yuv420_to_bgr, dim3(cdiv(cdiv(w, 8), 64), h / 2, n), 64, 0, planes[0], planes[1], planes[2], h, w,
               stride[0], stride[1], stride[2], frame_stride[0], frame_stride[1], frame_stride[2], bgr, vec ? 1 : 0);
    return VQA_OK;
}

// yuv420p -> gray + the four histograms in ONE pass (native-resolution analysis of an encode handed over as planes): the
// BGR triple of a pixel exists only in registers -- converted exactly as above, fed to OpenCV's BGR -> gray and to the
// B / G / R / gray histograms -- so the 3 B/px BGR frame is neither written nor read back (k_yuv420_to_bgr + k_gray_hist
// moved 8.5 B/px, this kernel 2.5).  One thread: 8 pixels x 2 rows per work item (64-bit luma loads, 32-bit chroma loads,
// 64-bit gray stores); histograms are warp-private in shared memory like k_gray_hist's and flushed with one global atomic
// per non-empty bin and block.  Requires the vector layout (w % 8 == 0, 8-byte aligned rows): run_yuv420_gray_hist says
// when it does not apply.
constexpr int YG_THREADS = 256, YG_WARPS = YG_THREADS / 32;

template <bool HIST>
__global__ void __launch_bounds__(YG_THREADS)
k_yuv420_gray_hist(const uint8_t *__restrict__ Yp, const uint8_t *__restrict__ Up, const uint8_t *__restrict__ Vp,
                   int h, int w, int sy, int su, int sv, size_t fy, size_t fu, size_t fv, uint8_t *__restrict__ gray,
                   uint32_t *__restrict__ hist)
{
    __shared__ unsigned sh[HIST ? YG_WARPS * 1024 : 1];
    const int frame = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned *wh = sh + (HIST ? warp * 1024 : 0);
    if (HIST) {
        for (int i = threadIdx.x; i < YG_WARPS * 1024; i += YG_THREADS) sh[i] = 0;
        __syncthreads();
    }
    const uint8_t *yb = Yp + frame * fy, *ub = Up + frame * fu, *vb = Vp + frame * fv;
    uint8_t *gb = gray + (size_t)frame * h * w;
    const int wq = w >> 3, items = (h >> 1) * wq, stride = gridDim.x * YG_THREADS;
    for (int base = blockIdx.x * YG_THREADS; base < items; base += stride) {
        const int it = base + threadIdx.x;
        const bool valid = it < items;
        unsigned ya[2] = {0, 0}, yc[2] = {0, 0}, u4 = 0, v4 = 0;
        int r2 = 0, xq = 0;
        if (valid) {
            r2 = it / wq;
            xq = it - r2 * wq;
            const uint2 a = __ldg(reinterpret_cast<const uint2 *>(yb + (size_t)(2 * r2) * sy + 8 * xq));
            const uint2 b = __ldg(reinterpret_cast<const uint2 *>(yb + (size_t)(2 * r2 + 1) * sy + 8 * xq));
            ya[0] = a.x; ya[1] = a.y; yc[0] = b.x; yc[1] = b.y;
            u4 = __ldg(reinterpret_cast<const unsigned *>(ub + (size_t)r2 * su + 4 * xq));
            v4 = __ldg(reinterpret_cast<const unsigned *>(vb + (size_t)r2 * sv + 4 * xq));
        }
        unsigned g0[2] = {0, 0}, g1[2] = {0, 0};
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int cb, cg, cr;
            yuv_chroma((int)((u4 >> (8 * j)) & 255u), (int)((v4 >> (8 * j)) & 255u), cb, cg, cr);
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int p = 2 * j + k;
#pragma unroll
                for (int row = 0; row < 2; row++) {
                    const unsigned Yv = ((row ? yc[p >> 2] : ya[p >> 2]) >> (8 * (p & 3))) & 255u;
                    uint8_t B, G, R;
                    yuv_px((int)Yv, cb, cg, cr, B, G, R);
                    const unsigned gv = gray_of(B, G, R);
                    if (row) g1[p >> 2] |= gv << (8 * (p & 3));
                    else g0[p >> 2] |= gv << (8 * (p & 3));
                    if (HIST) hist_add4(wh, B, G, R, gv, valid, vm, lane);
                }
            }
        }
        if (valid) {
            *reinterpret_cast<uint2 *>(gb + (size_t)(2 * r2) * w + 8 * xq) = make_uint2(g0[0], g0[1]);
            *reinterpret_cast<uint2 *>(gb + (size_t)(2 * r2 + 1) * w + 8 * xq) = make_uint2(g1[0], g1[1]);
        }
    }
    if (HIST) {
        __syncthreads();
        uint32_t *gh = hist + (size_t)frame * 1024;
        for (int i = threadIdx.x; i < 1024; i += YG_THREADS) {
            unsigned s = 0;
#pragma unroll
            for (int k = 0; k < YG_WARPS; k++) s += sh[k * 1024 + i];
            if (s) atomicAdd(&gh[i], s);
        }
    }
}

bool yuv420_gray_hist_ok(const uint8_t *const planes[3], const int stride[3], const size_t frame_stride[3], int h, int w,
                         const uint8_t *gray)
{
    return ((h | w) & 1) == 0 && (w % 8 == 0) && (stride[0] % 8 == 0) && (stride[1] % 4 == 0) && (stride[2] % 4 == 0) &&
           (frame_stride[0] % 8 == 0) && (frame_stride[1] % 4 == 0) && (frame_stride[2] % 4 == 0) &&
           (((uintptr_t)planes[0] | (uintptr_t)gray) % 8 == 0) && (((uintptr_t)planes[1] | (uintptr_t)planes[2]) % 4 == 0);
}

int run_yuv420_gray_hist(vqa_ctx *c, const uint8_t *const planes[3], const int stride[3], const size_t frame_stride[3],
                         int n, int h, int w, uint8_t *gray, uint32_t *hist)
{
    if (n <= 0) return VQA_OK;
    if (!yuv420_gray_hist_ok(planes, stride, frame_stride, h, w, gray))
        return set_err(c, VQA_E_UNSUPPORTED, "fused yuv420p -> gray path needs the vector layout (%dx%d)", w, h);
    int bpf = cdiv((h / 2) * (w / 8), YG_THREADS * 8);
    if (bpf < 1) bpf = 1;
    VQA_BYTES(c, 2.5 * h * w * n);
    if (hist) {
        VQA_CUDA(c, cudaMemsetAsync(hist, 0, sizeof(uint32_t) * 1024 * (size_t)n, c->stream));
        VQA_LAUNCH(c, k_yuv420_gray_hist<true>, dim3(bpf, n), YG_THREADS, 0, planes[0], planes[1], planes[2], h, w, stride[0],
                   stride[1], stride[2], frame_stride[0], frame_stride[1], frame_stride[2], gray, hist);
    } else {
        VQA_LAUNCH(c, k_yuv420_gray_hist<false>, dim3(bpf, n), YG_THREADS, 0, planes[0], planes[1], planes[2], h, w, stride[0],
                   stride[1], stride[2], frame_stride[0], frame_stride[1], frame_stride[2], gray, hist);
    }
    return VQA_OK;
}

}  // namespace vqa
