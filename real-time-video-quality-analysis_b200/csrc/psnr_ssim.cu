// Full-reference PSNR + SSIM on 8-bit planes in ONE pass over the pair (replaces the
// `psnr=` / `ssim=` FFmpeg filter graphs of run_ffmpeg_metrics, video_processing.py:270-297).
//
// FFmpeg semantics (vf_psnr.c / vf_ssim.c): per-plane SSE over every pixel;
// SSIM from integer 4x4 block sums (s1, s2, ss, s12), 8x8 windows = 2x2 groups of blocks on a
// 4-px grid, ssim_c1 = 416, ssim_c2 = 235963, float per-window value, mean over
// (w/4-1)(h/4-1) windows.  Integer sums are exact; the per-window floats are accumulated in
// double.  Roofline: HBM, algorithmic bytes = 2 * plane bytes (each pixel of both planes read once).
//
// Two kernels:
//   k_psnr_ssim_v   the production path for 16-byte aligned planes whose width divides by 16 (every
//                   yuv420p plane of 1080p / 4K video): all three planes and all frames of a chunk in ONE
//                   launch.  A warp owns a strip of 512 columns and marches down block rows: each lane
//                   loads 16 px x 4 rows of both planes with 128-bit loads (next block row prefetched
//                   while the current one is reduced), builds its four 4x4 block sums with dp4a, gets
//                   the block to its right by shuffle and keeps the block row above in registers, so no
//                   pixel is read twice inside a strip segment (one overlap block row per 32).  SSE of a
//                   block is ss - 2 s12 (exact), so PSNR costs no extra arithmetic.  Per-warp partials
//                   go to a table and k_fr_finish adds them in a fixed order: results are deterministic.
//   k_psnr_ssim     generic shapes (odd widths, unaligned strides): 32x8 tiles of blocks in shared memory.
#include "vqa_common.cuh"

namespace vqa {

constexpr int PB_X = 32, PB_Y = 8;

__device__ __forceinline__ float ssim_end1(int s1, int s2, int ss, int s12)
{
    const int vars = ss * 64 - s1 * s1 - s2 * s2, covar = s12 * 64 - s1 * s2;
    const float num = __fmul_rn((float)(2 * s1 * s2 + 416), (float)(2 * covar + 235963));
    const float den = __fmul_rn((float)(s1 * s1 + s2 * s2 + 416), (float)(vars + 235963));
    return __fdiv_rn(num, den);
}

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
k_psnr_ssim(const uint8_t *__restrict__ A, const uint8_t *__restrict__ B, int h, int w, int stride, size_t frame_stride,
            unsigned long long *__restrict__ sse_out, double *__restrict__ ssim_out)
{
    __shared__ int4 sums[PB_Y + 1][PB_X + 1];
    __shared__ double red_d[8];
    __shared__ unsigned long long red_u[8];
    const int frame = blockIdx.z;
    const uint8_t *a = A + (size_t)frame * frame_stride, *b = B + (size_t)frame * frame_stride;
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
            v = (double)ssim_end1(p.x + q.x + r.x + t.x, p.y + q.y + r.y + t.y, p.z + q.z + r.z + t.z, p.w + q.w + r.w + t.w);
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

// ------------------------------------------------------------------------------ vector path
constexpr int FR_SEG = 32;                       // block rows (of 4 pixel rows) per strip segment

struct FrPlane {
    int w, h, stride, bw, bh, strips, segs, unit0;
    unsigned long long frame_stride;
};
struct FrGeom {
    FrPlane p[3];
    int units;                                   // work units (warps) per frame over the three planes
};
struct FrPartial {
    double ssim;
    unsigned long long sse;
};

// sums of the four 4x4 blocks of a 16-pixel x 4-row patch: one 128-bit word per row and plane
__device__ __forceinline__ void fr_patch_sums(const uint4 (&ra)[4], const uint4 (&rb)[4], int (&s1)[5], int (&s2)[5],
                                              int (&ss)[5], int (&s12)[5])
{
#pragma unroll
    for (int j = 0; j < 4; j++) {
        unsigned a1 = 0, a2 = 0, as = 0, ax = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const unsigned wa = j == 0 ? ra[r].x : j == 1 ? ra[r].y : j == 2 ? ra[r].z : ra[r].w;
            const unsigned wb = j == 0 ? rb[r].x : j == 1 ? rb[r].y : j == 2 ? rb[r].z : rb[r].w;
            a1 = __dp4a(wa, 0x01010101u, a1);
            a2 = __dp4a(wb, 0x01010101u, a2);
            as = __dp4a(wa, wa, as);
            as = __dp4a(wb, wb, as);
            ax = __dp4a(wa, wb, ax);
        }
        s1[j] = (int)a1; s2[j] = (int)a2; ss[j] = (int)as; s12[j] = (int)ax;
    }
}

__global__ void __launch_bounds__(128)
k_psnr_ssim_v(const uint8_t *__restrict__ A0, const uint8_t *__restrict__ A1, const uint8_t *__restrict__ A2,
              const uint8_t *__restrict__ B0, const uint8_t *__restrict__ B1, const uint8_t *__restrict__ B2,
              FrGeom g, int n, FrPartial *__restrict__ part)
{
    const int lane = threadIdx.x & 31;
    const int wg = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int frame = wg / g.units;
    if (frame >= n) return;                                           // warp-uniform
    int u = wg - frame * g.units;
    const int pl = u >= g.p[2].unit0 ? 2 : (u >= g.p[1].unit0 ? 1 : 0);
    const FrPlane P = pl == 0 ? g.p[0] : (pl == 1 ? g.p[1] : g.p[2]);
    u -= P.unit0;
    const int seg = u / P.strips, strip = u - seg * P.strips;
    const uint8_t *a = (pl == 0 ? A0 : pl == 1 ? A1 : A2) + (size_t)frame * P.frame_stride;
    const uint8_t *b = (pl == 0 ? B0 : pl == 1 ? B1 : B2) + (size_t)frame * P.frame_stride;
    const int x0 = (strip * 32 + lane) * 16;
    const bool active = x0 < P.w;                                      // w % 16 == 0: a lane is all in or all out
    const bool edge = active && (lane == 31) && (x0 + 16 < P.w);       // the block to the right lives in the next strip
    const int bx0 = x0 >> 2;
    const int br0 = seg * FR_SEG, br_own = min(br0 + FR_SEG, P.bh), br_end = min(br_own + 1, P.bh);
    const uint8_t *pa = a + (size_t)(br0 * 4) * P.stride + x0, *pb = b + (size_t)(br0 * 4) * P.stride + x0;
    const size_t rstep = (size_t)P.stride;
    uint4 na[4], nb[4];
    unsigned ea[4], eb[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        na[r] = nb[r] = make_uint4(0, 0, 0, 0);
        ea[r] = eb[r] = 0;
    }
    auto load_row = [&](const uint8_t *qa, const uint8_t *qb) {
        if (active) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                na[r] = ld_stream(reinterpret_cast<const uint4 *>(qa + r * rstep));
                nb[r] = ld_stream(reinterpret_cast<const uint4 *>(qb + r * rstep));
            }
        }
        if (edge) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                ea[r] = __ldg(reinterpret_cast<const unsigned *>(qa + r * rstep + 16));
                eb[r] = __ldg(reinterpret_cast<const unsigned *>(qb + r * rstep + 16));
            }
        }
    };
    if (br0 < br_end) load_row(pa, pb);
    int p1[5], p2[5], ps[5], p12[5];                                   // block row above (index 4 = right neighbour)
#pragma unroll
    for (int j = 0; j < 5; j++) p1[j] = p2[j] = ps[j] = p12[j] = 0;
    double v = 0;
    unsigned long long sse = 0;
    for (int br = br0; br < br_end; br++) {
        uint4 ca[4], cb[4];
        unsigned fa[4], fb[4];
#pragma unroll
        for (int r = 0; r < 4; r++) { ca[r] = na[r]; cb[r] = nb[r]; fa[r] = ea[r]; fb[r] = eb[r]; }
        pa += 4 * rstep;
        pb += 4 * rstep;
        if (br + 1 < br_end) load_row(pa, pb);                         // prefetch the next block row
        int c1[5], c2[5], cs[5], c12[5];
        fr_patch_sums(ca, cb, c1, c2, cs, c12);
        // block to the right: first block of the next lane, or the extra words of the strip's last lane
        c1[4] = __shfl_down_sync(0xffffffffu, c1[0], 1);
        c2[4] = __shfl_down_sync(0xffffffffu, c2[0], 1);
        cs[4] = __shfl_down_sync(0xffffffffu, cs[0], 1);
        c12[4] = __shfl_down_sync(0xffffffffu, c12[0], 1);
        if (lane == 31) {
            unsigned a1 = 0, a2 = 0, as = 0, ax = 0;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                a1 = __dp4a(fa[r], 0x01010101u, a1);
                a2 = __dp4a(fb[r], 0x01010101u, a2);
                as = __dp4a(fa[r], fa[r], as);
                as = __dp4a(fb[r], fb[r], as);
                ax = __dp4a(fa[r], fb[r], ax);
            }
            c1[4] = (int)a1; c2[4] = (int)a2; cs[4] = (int)as; c12[4] = (int)ax;
        }
        if (br < br_own && active) {
#pragma unroll
            for (int j = 0; j < 4; j++) sse += (unsigned)(cs[j] - 2 * c12[j]);
        }
        if (br > br0 && active) {                                      // windows whose top block row is br - 1
            float rowv = 0.f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (bx0 + j < P.bw - 1)
                    rowv += ssim_end1(p1[j] + p1[j + 1] + c1[j] + c1[j + 1], p2[j] + p2[j + 1] + c2[j] + c2[j + 1],
                                      ps[j] + ps[j + 1] + cs[j] + cs[j + 1], p12[j] + p12[j + 1] + c12[j] + c12[j + 1]);
            }
            v += (double)rowv;
        }
#pragma unroll
        for (int j = 0; j < 5; j++) { p1[j] = c1[j]; p2[j] = c2[j]; ps[j] = cs[j]; p12[j] = c12[j]; }
    }
    // rows below the 4-row grid (h % 4) only enter the SSE; the plane's last segment takes them
    if (seg == P.segs - 1 && active) {
        for (int y = P.bh * 4; y < P.h; y++) {
            const uint4 wa = ld_stream(reinterpret_cast<const uint4 *>(a + (size_t)y * P.stride + x0));
            const uint4 wb = ld_stream(reinterpret_cast<const uint4 *>(b + (size_t)y * P.stride + x0));
            const unsigned d0 = __vabsdiffu4(wa.x, wb.x), d1 = __vabsdiffu4(wa.y, wb.y), d2 = __vabsdiffu4(wa.z, wb.z),
                           d3 = __vabsdiffu4(wa.w, wb.w);
            sse += __dp4a(d0, d0, __dp4a(d1, d1, __dp4a(d2, d2, __dp4a(d3, d3, 0u))));
        }
    }
    v = warp_sum(v);
    sse = warp_sum(sse);
    if (lane == 0) {
        FrPartial o;
        o.ssim = v;
        o.sse = sse;
        part[(size_t)frame * g.units + (wg - frame * g.units)] = o;
    }
}

// fixed-order sum of the per-warp partials of one (frame, plane)
__global__ void k_fr_finish(const FrPartial *__restrict__ part, FrGeom g, int n, unsigned long long *__restrict__ sse_out,
                            double *__restrict__ ssim_out, int out_pitch)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * n) return;
    const int pl = i / n, frame = i - pl * n;
    const int u0 = g.p[pl].unit0, u1 = pl == 2 ? g.units : g.p[pl + 1].unit0;
    double s = 0;
    unsigned long long e = 0;
    for (int u = u0; u < u1; u++) {
        const FrPartial q = part[(size_t)frame * g.units + u];
        s += q.ssim;
        e += q.sse;
    }
    ssim_out[(size_t)pl * out_pitch + frame] = s;
    sse_out[(size_t)pl * out_pitch + frame] = e;
}

int run_psnr_ssim_plane(vqa_ctx *c, const uint8_t *a, const uint8_t *b, int n, int h, int w, int stride, size_t frame_stride,
                        unsigned long long *sse, double *ssim_sum)
{
    VQA_CUDA(c, cudaMemsetAsync(sse, 0, sizeof(unsigned long long) * (size_t)n, c->stream));
    VQA_CUDA(c, cudaMemsetAsync(ssim_sum, 0, sizeof(double) * (size_t)n, c->stream));
    const int bw = w >> 2, bh = h >> 2;
    dim3 grid(cdiv(bw > 0 ? bw : 1, PB_X), cdiv(bh > 0 ? bh : 1, PB_Y), n);
    VQA_BYTES(c, 2.0 * h * w * n);
    VQA_LAUNCH(c, k_psnr_ssim, grid, PB_X * PB_Y, 0, a, b, h, w, stride, frame_stride, sse, ssim_sum);
    return VQA_OK;
}

// PSNR + SSIM sums of n frames of three planes: sse / ssim_sum are [3][out_pitch] (plane-major), this call
// fills columns 0..n-1.  Frame f of plane p starts at plane[p] + f * frame_stride[p].
int run_psnr_ssim_planes(vqa_ctx *c, const uint8_t *const a[3], const uint8_t *const b[3], int n, const int plane_h[3],
                         const int plane_w[3], const int stride[3], const size_t frame_stride[3],
                         unsigned long long *sse, double *ssim_sum, int out_pitch)
{
    if (n <= 0) return VQA_OK;
    bool fast = true;
    for (int p = 0; p < 3; p++)
        fast = fast && plane_w[p] % 16 == 0 && stride[p] % 16 == 0 && frame_stride[p] % 16 == 0 && plane_h[p] >= 4 &&
               (((uintptr_t)a[p] | (uintptr_t)b[p]) & 15) == 0;
    if (!fast) {
        for (int p = 0; p < 3; p++) {
            int rc = run_psnr_ssim_plane(c, a[p], b[p], n, plane_h[p], plane_w[p], stride[p], frame_stride[p],
                                         sse + (size_t)p * out_pitch, ssim_sum + (size_t)p * out_pitch);
            if (rc) return rc;
        }
        return VQA_OK;
    }
    FrGeom g;
    int units = 0;
    double bytes = 0;
    for (int p = 0; p < 3; p++) {
        FrPlane &P = g.p[p];
        P.w = plane_w[p]; P.h = plane_h[p]; P.stride = stride[p];
        P.bw = P.w >> 2; P.bh = P.h >> 2;
        P.strips = cdiv(P.w, 512);
        P.segs = cdiv(P.bh, FR_SEG);
        P.unit0 = units;
        P.frame_stride = frame_stride[p];
        units += P.strips * P.segs;
        bytes += 2.0 * P.w * P.h * n;
    }
    g.units = units;
    VQA_BUF(c, part, FrPartial, "fr.part", (size_t)n * units);
    VQA_BYTES(c, bytes);
    VQA_LAUNCH(c, k_psnr_ssim_v, cdiv((long)n * units, 4), 128, 0, a[0], a[1], a[2], b[0], b[1], b[2], g, n, part);
    VQA_LAUNCH(c, k_fr_finish, cdiv(3L * n, 128), 128, 0, part, g, n, sse, ssim_sum, out_pitch);
    return VQA_OK;
}

}  // namespace vqa
