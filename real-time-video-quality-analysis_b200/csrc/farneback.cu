// "Advanced motion complexity" (process_frame_complexity, complexity_metrics.py:313-343):
//   mean |calcOpticalFlowFarneback(prev_gray, curr_gray, None, 0.5, 3, 15, 3, 5, 1.2, 0)|
// restated for the GPU from the published algorithm (OpenCV optflowgf.cpp):
//
//   per level k = levels..0 (coarse -> fine), batched over all pairs of the chunk:
//     k_fb_pyramid      I_k  = resize_f32(GaussianBlur(float(gray)), level size)   per FRAME
//     k_fb_polyexp      R_k  = 11x11 separable polynomial expansion (5 x f32 / px)  per FRAME
//     k_fb_upsample     flow = 2 * resize_f32(flow_{k+1})  (zeros at the coarsest level)
//     k_fb_matrices     M    = UpdateMatrices(R_prev, R_cur, flow)                   per PAIR
//     3 x k_fb_blur_solve  flow = solve2x2(boxmean15x15(M));  M rebuilt after iterations 1, 2
//   (last blur pass)    sum sqrt(fx^2 + fy^2) at level 0, fused into k_fb_blur_solve
//
// R is computed once per frame and serves the frame both as "next" of one pair and as "prev"
// of the following pair.  R and M are stored per pixel as one float4 (components 0..3) + one float (component 4) in
// two arrays: the four bilinear taps of UpdateMatrices are 4 x (128-bit + 32-bit) loads instead of 20 scalar gathers,
// the marching blur reads a row as 2 loads instead of 5 (round 2; planar storage made both kernels L1-bound).
// R, M and flow stay fp32 (reduced precision does not survive flat
// content).  Roofline: HBM.
#include <math.h>
#include <stdlib.h>

#include "vqa_common.cuh"

namespace vqa {

struct GaussTaps {
    int ksz;
    float k[32];
};

struct PolyConst {
    float g[11], xg[11], xxg[11];
    double ig11, ig03, ig33, ig55;
};

__device__ __forceinline__ int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

__device__ __forceinline__ void lin_tap_f32(int d, int sn, int dn, bool vertical, int &i0, int &i1, float &a)
{
    // sn/dn is an exact power of two for every pyramid level of an image whose sides divide by 8;
    // the reciprocal-multiply form below is then bit-identical to the division and avoids the
    // ~30-instruction double division per call
    const double scale = (sn == 2 * dn) ? 2.0 : (2 * sn == dn) ? 0.5 : (sn == 4 * dn) ? 4.0 : (sn == 8 * dn) ? 8.0
                                                                                                  : (double)sn / (double)dn;
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int i = (int)floorf(f);
    a = f - (float)i;
    if (!vertical) {
        if (i < 0) { i = 0; a = 0.f; }
        if (i >= sn - 1) { i = sn - 1; a = 0.f; }
    }
    i0 = clampi(i, 0, sn - 1);
    i1 = clampi(i + 1, 0, sn - 1);
}

// Source taps of one destination index.  mode 0: same size; 1: exact 2x decimation (cv2.resize
// switches INTER_LINEAR to the INTER_AREA fast path: mean of the 2x2 block); 2: bilinear.
__device__ __forceinline__ void py_tap(int d, int sn, int dn, int mode, bool vertical, int &p0, int &p1, float &a)
{
    if (mode == 0) { p0 = p1 = d; a = 0.f; }
    else if (mode == 1) { p0 = 2 * d; p1 = 2 * d + 1; a = 0.5f; }
    else lin_tap_f32(d, sn, dn, vertical, p0, p1, a);
}

constexpr int PY_TW = 32, PY_TH = 8;

// I_k = resize_f32(GaussianBlur(float(gray), ksz, sigma, REFLECT_101), level size), one output per
// thread, shared-memory tiled: the uint8 source region of the tile (reflect-101 applied while
// loading) -> horizontal Gaussian at the <= 64 source columns the tile samples -> vertical Gaussian
// at the <= 2 rows each output samples -> area / bilinear combine.  Evaluation order matches
// OpenCV's separable float32 filter (rows then columns, centre tap + symmetric pairs).
// RT > 0: Gaussian radius known at compile time (taps in registers, loops unrolled); RT = 0: generic
template <int RT>
__global__ void __launch_bounds__(PY_TW * PY_TH)
k_fb_pyramid(const uint8_t *__restrict__ gray, int H, int W, int lh, int lw, int mode, GaussTaps taps,
             float *__restrict__ I, int rw_pitch)
{
    extern __shared__ __align__(16) uint8_t py_smem[];
    __shared__ int xtab[2 * PY_TW];
    const int frame = blockIdx.z, tid = threadIdx.x;
    const uint8_t *img = gray + (size_t)frame * H * W;
    const int tx0 = blockIdx.x * PY_TW, ty0 = blockIdx.y * PY_TH;
    const int r = RT > 0 ? RT : (taps.ksz >> 1), np = mode == 0 ? 1 : 2;
    float tk[RT > 0 ? RT + 1 : 1];                                     // tk[k] = tap at distance k from the centre
    if (RT > 0) {
#pragma unroll
        for (int k = 0; k <= RT; k++) tk[k] = taps.k[RT + k];
    }
    int p0, p1, q0, q1;
    float fa;
    py_tap(tx0, W, lw, mode, false, p0, q0, fa);
    py_tap(min(tx0 + PY_TW - 1, lw - 1), W, lw, mode, false, q1, p1, fa);
    const int x_lo = (p0 - r) & ~3, RW = p1 + r - x_lo + 1;          // region starts on a 4-byte boundary of the source row
    py_tap(ty0, H, lh, mode, true, p0, q0, fa);
    py_tap(min(ty0 + PY_TH - 1, lh - 1), H, lh, mode, true, q1, p1, fa);
    const int y_lo = p0 - r, RH = p1 + r - y_lo + 1;
    uint8_t *src = py_smem;                                            // [RH][rw_pitch]
    float *hb = reinterpret_cast<float *>(py_smem + (((size_t)RH * rw_pitch + 15) & ~(size_t)15));   // [RH][np*PY_TW]
    if (tid < np * PY_TW) {
        int a0, a1;
        py_tap(min(tx0 + tid / np, lw - 1), W, lw, mode, false, a0, a1, fa);
        xtab[tid] = ((np == 2 && (tid & 1)) ? a1 : a0) - x_lo;
    }
    const int lane = tid & 31, wrp = tid >> 5;
    const int RW4 = (RW + 3) >> 2;                                     // 32-bit words per region row (<= rw_pitch / 4)
    if ((W & 3) == 0 && x_lo >= 0 && x_lo + 4 * RW4 <= W) {
        // interior columns: whole words, all loads of a thread independent (the byte loop below spent
        // ~750 instructions per thread on LDG.U8 / STS.U8 pairs)
        for (int i = tid; i < RH * RW4; i += PY_TW * PY_TH) {
            const int ry = i / RW4, rq = i - ry * RW4;
            const uint8_t *g = img + (size_t)reflect101(y_lo + ry, H) * W + x_lo;
            reinterpret_cast<uint32_t *>(src + ry * rw_pitch)[rq] = __ldg(reinterpret_cast<const uint32_t *>(g) + rq);
        }
    } else {
        for (int ry = wrp; ry < RH; ry += PY_TH) {                     // one warp per source row
            const uint8_t *g = img + (size_t)reflect101(y_lo + ry, H) * W;
            uint8_t *d = src + ry * rw_pitch;
            for (int rx = lane; rx < RW; rx += 32) {
                const int gxx = x_lo + rx;
                d[rx] = __ldg(g + ((gxx >= 0 && gxx < W) ? gxx : reflect101(gxx, W)));
            }
        }
    }
    __syncthreads();
    const int hbw = np * PY_TW;                                        // 32 or 64
    {
        const int p = tid & (hbw - 1), rstep = (PY_TW * PY_TH) / hbw;
        const int xo = xtab[p];
        for (int ry = tid / hbw; ry < RH; ry += rstep) {
            const uint8_t *c = src + ry * rw_pitch + xo;
            // (float)(a + b) == (float)a + (float)b exactly for 8-bit a, b: one int->float conversion
            // (quarter-rate XU pipe) per symmetric tap pair instead of two
            float s;
            if (RT > 0) {
                s = tk[0] * (float)c[0];
#pragma unroll
                for (int k = 1; k <= RT; k++) s += tk[k] * (float)((int)c[-k] + (int)c[k]);
            } else {
                s = taps.k[r] * (float)c[0];
                for (int k = 1; k <= r; k++) s += taps.k[r + k] * (float)((int)c[-k] + (int)c[k]);
            }
            hb[ry * hbw + p] = s;
        }
    }
    __syncthreads();
    const int x = tx0 + (tid & (PY_TW - 1)), y = ty0 + tid / PY_TW;
    if (x >= lw || y >= lh) return;
    int ya, yb, xa, xb;
    float ax, ay;
    py_tap(y, H, lh, mode, true, ya, yb, ay);
    py_tap(x, W, lw, mode, false, xa, xb, ax);
    float v[2][2];
    for (int wy = 0; wy < np; wy++) {
        const int yy = (wy ? yb : ya) - y_lo;
        for (int wx = 0; wx < np; wx++) {
            const float *c = hb + (size_t)yy * hbw + (tid & (PY_TW - 1)) * np + wx;
            float s;
            if (RT > 0) {
                s = tk[0] * c[0];
#pragma unroll
                for (int k = 1; k <= RT; k++) s += tk[k] * (c[-k * hbw] + c[k * hbw]);
            } else {
                s = taps.k[r] * c[0];
                for (int k = 1; k <= r; k++) s += taps.k[r + k] * (c[-k * hbw] + c[k * hbw]);
            }
            v[wy][wx] = s;
        }
    }
    float out;
    if (mode == 0) out = v[0][0];
    else if (mode == 1) out = (v[0][0] + v[0][1] + v[1][0] + v[1][1]) * 0.25f;
    else {
        const float a0 = 1.f - ax, b0 = 1.f - ay;
        const float t0 = __fadd_rn(__fmul_rn(v[0][0], a0), __fmul_rn(v[0][1], ax));
        const float t1 = __fadd_rn(__fmul_rn(v[1][0], a0), __fmul_rn(v[1][1], ax));
        out = __fadd_rn(__fmul_rn(t0, b0), __fmul_rn(t1, ay));
    }
    I[(size_t)frame * lh * lw + (size_t)y * lw + x] = out;
}

// Fast path of the two largest pyramid levels (3-tap Gaussian): MODE 0 = level 0 (same size),
// MODE 1 = level 1 (exact 2x decimation = mean of the 2x2 block of blurred pixels).  A thread owns
// 4 adjacent SOURCE columns: 32-bit shared loads of the uint8 tile, 4 (MODE 0) or 2 (MODE 1)
// outputs written with one vector store.  Same evaluation order as the generic kernel.
template <int MODE>
__global__ void __launch_bounds__(256)
k_fb_pyramid3(const uint8_t *__restrict__ gray, int H, int W, int lh, int lw, float ke, float kc, float *__restrict__ I)
{
    constexpr int SR = MODE == 0 ? 10 : 18, NR = MODE == 0 ? 3 : 4, TROWS = MODE == 0 ? 8 : 16;
    __shared__ __align__(16) uint8_t tile[SR][144];                  // source columns sx0-4 .. sx0+131
    const int frame = blockIdx.z, tid = threadIdx.x;
    const uint8_t *img = gray + (size_t)frame * H * W;
    const int sx0 = blockIdx.x * 128, sy0 = blockIdx.y * TROWS;
    const bool w4 = (W & 3) == 0;
    for (int i = tid; i < SR * 34; i += 256) {
        const int ry = i / 34, wx = i - ry * 34;
        const uint8_t *row = img + (size_t)reflect101(sy0 - 1 + ry, H) * W;
        const int gx = sx0 - 4 + wx * 4;
        uint32_t v;
        if (w4 && gx >= 0 && gx + 4 <= W) v = __ldg(reinterpret_cast<const uint32_t *>(row + gx));
        else {
            v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int x = gx + b;
                const int xr = (x >= -1 && x <= W) ? reflect101(x, W) : clampi(x, 0, W - 1);   // only x = -1 / W are ever used
                v |= (uint32_t)__ldg(row + xr) << (8 * b);
            }
        }
        *reinterpret_cast<uint32_t *>(&tile[ry][wx * 4]) = v;
    }
    __syncthreads();
    const int xq = tid & 31, ty = tid >> 5;
    float hr[NR][4];
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(&tile[(MODE == 0 ? ty : 2 * ty) + r][4 * xq]);
        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
        const float b[6] = {(float)(w0 >> 24), (float)(w1 & 255u), (float)((w1 >> 8) & 255u), (float)((w1 >> 16) & 255u),
                            (float)(w1 >> 24), (float)(w2 & 255u)};
#pragma unroll
        for (int j = 0; j < 4; j++) hr[r][j] = kc * b[j + 1] + ke * (b[j] + b[j + 2]);
    }
    if (MODE == 0) {
        const int y = sy0 + ty, x = sx0 + 4 * xq;
        if (y >= lh || x >= lw) return;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) o[j] = kc * hr[1][j] + ke * (hr[0][j] + hr[2][j]);
        float *dst = I + (size_t)frame * lh * lw + (size_t)y * lw + x;
        if (w4 && x + 4 <= lw) *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        else
            for (int j = 0; j < 4; j++)
                if (x + j < lw) dst[j] = o[j];
    } else {
        const int y = sy0 / 2 + ty, x = sx0 / 2 + 2 * xq;
        if (y >= lh || x >= lw) return;
        float b0[4], b1[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            b0[j] = kc * hr[1][j] + ke * (hr[0][j] + hr[2][j]);
            b1[j] = kc * hr[2][j] + ke * (hr[1][j] + hr[3][j]);
        }
        const float o0 = (b0[0] + b0[1] + b1[0] + b1[1]) * 0.25f, o1 = (b0[2] + b0[3] + b1[2] + b1[3]) * 0.25f;
        float *dst = I + (size_t)frame * lh * lw + (size_t)y * lw + x;
        if ((lw & 1) == 0 && x + 2 <= lw) *reinterpret_cast<float2 *>(dst) = make_float2(o0, o1);
        else {
            dst[0] = o0;
            if (x + 1 < lw) dst[1] = o1;
        }
    }
}

// Pyramid levels 2 and 3 of a frame whose sides divide by 8 (exact 4x / 8x decimation; Gaussian radius 4 / 9): the same
// arithmetic as k_fb_pyramid in mode 2 -- horizontal Gaussian at the two source columns S*x + S/2 - 1, + 0 / + 1 each
// output samples (bilinear weight exactly 0.5), vertical Gaussian at the two source rows, combine -- restructured for
// instruction issue (the generic kernel spent 1.4 + 1.5 ms per 300 frames on two levels that hold 8 % of the pixels:
// byte loads from shared memory, a runtime division per loaded word, 20 % issue utilisation).  Here a horizontal work
// item produces BOTH columns of an output from one aligned 24- / 12-byte read of the uint8 tile (the two 19- / 9-tap
// windows overlap in all but one byte), rows are loaded with one 2-D thread mapping, and a vertical work item reads
// float2 pairs.  Tile: 32 x TO outputs per block.
template <int S, int R, int TO>
__global__ void __launch_bounds__(256)
k_fb_pyramid_dec(const uint8_t *__restrict__ gray, int H, int W, int lh, int lw, GaussTaps taps, float *__restrict__ I)
{
    constexpr int A = S == 8 ? 8 : 4;                   // tile column 0 = source column S*tx0 - A (keeps the work-item reads aligned)
    constexpr int NB = S == 8 ? 24 : 12;                // bytes a horizontal work item reads: tile columns S*x .. S*x + NB - 1
    constexpr int C0 = A + S / 2 - 1;                   // centre of the first window inside those bytes (11 / 5); second: C0 + 1
    constexpr int WB = 32 * S + NB - S, WP = (WB + 15) & ~15;       // tile width in bytes (272 / 136) and its pitch
    constexpr int RH = S * (TO - 1) + 2 * R + 2;        // source rows S*ty0 + S/2 - 1 - R .. (140 / 134)
    extern __shared__ __align__(16) uint8_t pd_smem[];
    uint8_t *src = pd_smem;                                             // [RH][WP]
    float2 *hb = reinterpret_cast<float2 *>(pd_smem + RH * WP);         // [RH][32]: both blurred columns of output x
    const int frame = blockIdx.z, tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const uint8_t *img = gray + (size_t)frame * H * W;
    const int tx0 = blockIdx.x * 32, ty0 = blockIdx.y * TO;
    const int xs0 = S * tx0 - A, ys0 = S * ty0 + S / 2 - 1 - R;
    float tk[R + 1];
#pragma unroll
    for (int k = 0; k <= R; k++) tk[k] = taps.k[R + k];
    // source columns at and beyond W + R + 1 feed no output that exists (the ragged last tile of a row)
    const int q_end = min(WP / 4, (W + R + 1 - xs0 + 3) >> 2);
    for (int ry = wrp; ry < RH; ry += 8) {
        const uint8_t *g = img + (size_t)reflect101(ys0 + ry, H) * W;
        uint32_t *d = reinterpret_cast<uint32_t *>(src + ry * WP);
        for (int q = lane; q < q_end; q += 32) {
            const int gx = xs0 + 4 * q;                                  // W % 4 == 0 (S divides W): words never straddle the border
            if (gx >= 0 && gx + 4 <= W) {
                // asynchronous copy: the ~40 words a thread fetches are all in flight at once (with plain loads the
                // row loop ran one dependent load at a time: 120 us per 24 frames, 4x its instruction time)
                const unsigned dst = (unsigned)__cvta_generic_to_shared(d + q);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(g + gx) : "memory");
            } else {
                uint32_t v = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) v |= (uint32_t)__ldg(g + reflect101(gx + b, W)) << (8 * b);
                d[q] = v;
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int ry = wrp; ry < RH; ry += 8) {
        uint32_t wv[NB / 4];
        if (S == 8) {
            const uint2 *p = reinterpret_cast<const uint2 *>(src + ry * WP + 8 * lane);
#pragma unroll
            for (int j = 0; j < 3; j++) { const uint2 v = p[j]; wv[2 * j] = v.x; wv[2 * j + 1] = v.y; }
        } else {
            const uint32_t *p = reinterpret_cast<const uint32_t *>(src + ry * WP + 4 * lane);
#pragma unroll
            for (int j = 0; j < 3; j++) wv[j] = p[j];
        }
#define PD_B(i) ((int)((wv[(i) >> 2] >> (8 * ((i) & 3))) & 255u))
        float s0 = tk[0] * (float)PD_B(C0), s1 = tk[0] * (float)PD_B(C0 + 1);
#pragma unroll
        for (int k = 1; k <= R; k++) {
            s0 += tk[k] * (float)(PD_B(C0 - k) + PD_B(C0 + k));
            s1 += tk[k] * (float)(PD_B(C0 + 1 - k) + PD_B(C0 + 1 + k));
        }
#undef PD_B
        hb[ry * 32 + lane] = make_float2(s0, s1);
    }
    __syncthreads();
    const int x = tx0 + lane;
    for (int oy = wrp; oy < TO; oy += 8) {
        const int y = ty0 + oy;
        if (y >= lh || x >= lw) continue;
        const float2 *c = hb + (S * oy + R) * 32 + lane;               // centre row of the first vertical window; second: + 1 row
        const float2 c0 = c[0], c1 = c[32];
        float v00 = tk[0] * c0.x, v01 = tk[0] * c0.y, v10 = tk[0] * c1.x, v11 = tk[0] * c1.y;
#pragma unroll
        for (int k = 1; k <= R; k++) {
            const float2 a0 = c[-k * 32], b0 = c[k * 32], a1 = c[(1 - k) * 32], b1 = c[(1 + k) * 32];
            v00 += tk[k] * (a0.x + b0.x);
            v01 += tk[k] * (a0.y + b0.y);
            v10 += tk[k] * (a1.x + b1.x);
            v11 += tk[k] * (a1.y + b1.y);
        }
        const float t0 = __fadd_rn(__fmul_rn(v00, 0.5f), __fmul_rn(v01, 0.5f));
        const float t1 = __fadd_rn(__fmul_rn(v10, 0.5f), __fmul_rn(v11, 0.5f));
        I[(size_t)frame * lh * lw + (size_t)y * lw + x] = __fadd_rn(__fmul_rn(t0, 0.5f), __fmul_rn(t1, 0.5f));
    }
}

template <int S, int R, int TO>
static int launch_pyramid_dec(vqa_ctx *c, const uint8_t *gray, int h, int w, int lh, int lw, int nf, const GaussTaps &taps, float *I)
{
    constexpr int NB = S == 8 ? 24 : 12, WP = ((32 * S + NB - S) + 15) & ~15, RH = S * (TO - 1) + 2 * R + 2;
    constexpr int smem = RH * WP + RH * 32 * (int)sizeof(float2);
    VQA_CUDA(c, cudaFuncSetAttribute(k_fb_pyramid_dec<S, R, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    VQA_LAUNCH(c, (k_fb_pyramid_dec<S, R, TO>), dim3(cdiv(lw, 32), cdiv(lh, TO), nf), 256, smem, gray, h, w, lh, lw, taps, I);
    return VQA_OK;
}

constexpr int PE_TW = 64, PE_TH = 32, PE_R = 5;
constexpr int PE_P = 76;                         // tile pitch (64 + 2*5 = 74, padded so rows stay 16-byte aligned)

// FarnebackPolyExp.  Register-blocked: every work item produces 4 adjacent columns from 128-bit
// shared-memory reads (vertical pass: 11 LDS.128 per 4 outputs, horizontal pass: 12 LDS.128 per 4
// outputs).  Both passes accumulate in float (OpenCV's horizontal accumulators are double; against the
// unmodified reference the clip mean moves by < 1e-7 rel, profiles/r01_notes.md).
// A block walks PE_STRIP vertically adjacent tiles; the pixel tile is double-buffered and filled with
// 4-byte cp.async (any alignment, border clamp in the address), so the loads of tile k+1 fly while
// tile k is computed (the synchronous version sat 63 % of its stall samples on the tile load).
constexpr int PE_STRIP = 4, PE_VR = 4;
constexpr int PE_SMEM = (2 * (PE_TH + 2 * PE_R) + 3 * PE_TH) * PE_P * (int)sizeof(float);

// One warp per tile row, lanes over the columns (three column groups of 32: 76 = 32 + 32 + 12): the clamped source column of a
// lane is computed once per tile and the row pointer once per row, so an element costs an add and the copy (the flat
// `i / PE_P` loop spent ~25 instructions per element on division, clamps and 64-bit address arithmetic: 15 % of the kernel).
__device__ __forceinline__ void pe_load_tile(float (*tile)[PE_P], const float *__restrict__ src, int h, int w, int tx0, int ty0, int tid)
{
    const int lane = tid & 31, wrp = tid >> 5;
    int cx[3];
#pragma unroll
    for (int k = 0; k < 3; k++) cx[k] = clampi(tx0 - PE_R + lane + 32 * k, 0, w - 1);
    for (int y = wrp; y < PE_TH + 2 * PE_R; y += 8) {
        const float *g = src + (size_t)clampi(ty0 - PE_R + y, 0, h - 1) * w;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&tile[y][lane]);
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (lane + 32 * k < PE_P)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 128u * k), "l"(g + cx[k]) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(256)
k_fb_polyexp(const float *__restrict__ I, int h, int w, PolyConst pc, float4 *__restrict__ R4, float *__restrict__ Rs)
{
    extern __shared__ __align__(16) float pe_smem[];             // PE_SMEM bytes (> 48 KB: opt-in dynamic)
    float (*tile)[PE_TH + 2 * PE_R][PE_P] = reinterpret_cast<float (*)[PE_TH + 2 * PE_R][PE_P]>(pe_smem);
    float (*v0)[PE_P] = reinterpret_cast<float (*)[PE_P]>(pe_smem + 2 * (PE_TH + 2 * PE_R) * PE_P);
    float (*v1)[PE_P] = v0 + PE_TH, (*v2)[PE_P] = v1 + PE_TH;
    const int frame = blockIdx.z, tid = threadIdx.x;
    const float *src = I + (size_t)frame * h * w;
    const int tx0 = blockIdx.x * PE_TW;
    const int tiles_y = (h + PE_TH - 1) / PE_TH, k0 = blockIdx.y * PE_STRIP, nk = min(PE_STRIP, tiles_y - k0);
    const size_t plane = (size_t)h * w;
    const int x4 = (tid & 15) * 4, gx0 = tx0 + x4;
    pe_load_tile(tile[0], src, h, w, tx0, k0 * PE_TH, tid);
    for (int kt = 0; kt < nk; kt++) {
        const int ty0 = (k0 + kt) * PE_TH;
        float (*T)[PE_P] = tile[kt & 1];
        if (kt + 1 < nk) {
            pe_load_tile(tile[(kt + 1) & 1], src, h, w, tx0, ty0 + PE_TH, tid);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        // vertical pass: a work item owns ONE column and PE_VR consecutive rows, so the 2 * PE_R + PE_VR tile values it needs are
        // read once (14 loads for 4 outputs; one row per item read 11 values per output: the kernel is bound by the L1 data
        // pipe, 86 %, and this pass was 55 % of its shared-memory wavefronts).  Same expressions per output as before.
        for (int i = tid; i < (PE_TH / PE_VR) * PE_P; i += 256) {
            const int yg = i / PE_P, x = i - yg * PE_P, y0 = yg * PE_VR;
            float r[PE_VR + 2 * PE_R];
#pragma unroll
            for (int q = 0; q < PE_VR + 2 * PE_R; q++) r[q] = T[y0 + q][x];
#pragma unroll
            for (int o = 0; o < PE_VR; o++) {
                float t0 = r[o + PE_R] * pc.g[PE_R], t1 = 0.f, t2 = 0.f;
#pragma unroll
                for (int k = 1; k <= PE_R; k++) {
                    const float a = r[o + PE_R - k], b = r[o + PE_R + k];
                    const float p = a + b;
                    t0 = t0 + pc.g[PE_R + k] * p;
                    t1 = t1 + pc.xg[PE_R + k] * (b - a);
                    t2 = t2 + pc.xxg[PE_R + k] * p;
                }
                v0[y0 + o][x] = t0;
                v1[y0 + o][x] = t1;
                v2[y0 + o][x] = t2;
            }
        }
        __syncthreads();
        if (gx0 < w) {
            for (int y = tid >> 4; y < PE_TH; y += 16) {
                const int gy = ty0 + y;
                if (gy >= h) break;
                float a0[16], a1[16], a2[16];                 // columns x4 .. x4+15 of the three vertical results
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 p0 = *reinterpret_cast<const float4 *>(&v0[y][x4 + 4 * q]);
                    const float4 p1 = *reinterpret_cast<const float4 *>(&v1[y][x4 + 4 * q]);
                    const float4 p2 = *reinterpret_cast<const float4 *>(&v2[y][x4 + 4 * q]);
                    a0[4 * q] = p0.x; a0[4 * q + 1] = p0.y; a0[4 * q + 2] = p0.z; a0[4 * q + 3] = p0.w;
                    a1[4 * q] = p1.x; a1[4 * q + 1] = p1.y; a1[4 * q + 2] = p1.z; a1[4 * q + 3] = p1.w;
                    a2[4 * q] = p2.x; a2[4 * q + 1] = p2.y; a2[4 * q + 2] = p2.z; a2[4 * q + 3] = p2.w;
                }
                float o[5][4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c = j + PE_R;
                    const float g0 = pc.g[PE_R];
                    float b1 = a0[c] * g0, b2 = 0, b3 = a1[c] * g0, b4 = 0, b5 = a2[c] * g0, b6 = 0;
#pragma unroll
                    for (int k = 1; k <= PE_R; k++) {
                        const float gk = pc.g[PE_R + k], xgk = pc.xg[PE_R + k], xxgk = pc.xxg[PE_R + k];
                        const float tg = (a0[c + k] + a0[c - k]);
                        b1 += tg * gk;
                        b4 += tg * xxgk;
                        b2 += (a0[c + k] - a0[c - k]) * xgk;
                        b3 += (a1[c + k] + a1[c - k]) * gk;
                        b6 += (a1[c + k] - a1[c - k]) * xgk;
                        b5 += (a2[c + k] + a2[c - k]) * gk;
                    }
                    o[0][j] = (float)(b3 * (float)pc.ig11);
                    o[1][j] = (float)(b2 * (float)pc.ig11);
                    o[2][j] = (float)(b1 * (float)pc.ig03 + b5 * (float)pc.ig33);
                    o[3][j] = (float)(b1 * (float)pc.ig03 + b4 * (float)pc.ig33);
                    o[4][j] = (float)(b6 * (float)pc.ig55);
                }
                const size_t o0 = (size_t)frame * plane + (size_t)gy * w + gx0;
                float4 *d4 = R4 + o0;
                float *d1 = Rs + o0;
                const bool vec = (w & 3) == 0 && gx0 + 4 <= w;          // arrays are 256-byte aligned, w % 4 == 0 keeps rows aligned
                if (vec) {
                    // the four pixels of this thread are 64 contiguous bytes: two 256-bit stores (sm_100), each a full
                    // 32-byte sector (four 128-bit stores with the lanes 64 bytes apart touched every sector twice)
#pragma unroll
                    for (int j = 0; j < 4; j += 2)
                        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(d4 + j), "f"(o[0][j]), "f"(o[1][j]),
                                     "f"(o[2][j]), "f"(o[3][j]), "f"(o[0][j + 1]), "f"(o[1][j + 1]), "f"(o[2][j + 1]), "f"(o[3][j + 1])
                                     : "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (gx0 + j < w) d4[j] = make_float4(o[0][j], o[1][j], o[2][j], o[3][j]);
                }
                if (vec) *reinterpret_cast<float4 *>(d1) = make_float4(o[4][0], o[4][1], o[4][2], o[4][3]);
                else
                    for (int j = 0; j < 4; j++)
                        if (gx0 + j < w) d1[j] = o[4][j];
            }
        }
        __syncthreads();                                     // v0..v2 and this tile buffer are rewritten next round
    }
}

// flow_k(x, y) = 2 * resize_f32(flow_{k+1}) (INTER_LINEAR upscale, OpenCV float tap rules).  The
// up-sampled flow is consumed only by the first UpdateMatrices of a level (the box-blur solve then
// rewrites the whole field), so it is evaluated on the fly there and never stored.
// lin_tap_f32 for an exact 2x up-scale (dn == 2 * sn), in integer arithmetic: (d + 0.5) * 0.5 - 0.5 =
// d / 2 - 0.25, so floor = (d - 1) >> 1 and the fraction is 0.75 for even d, 0.25 for odd d -- the same
// numbers the double-precision form produces (every intermediate is exact), without its FP64 work.
__device__ __forceinline__ void up2_tap_f32(int d, int sn, bool vertical, int &i0, int &i1, float &a)
{
    int i = (d - 1) >> 1;
    a = (d & 1) ? 0.25f : 0.75f;
    if (!vertical) {
        if (i < 0) { i = 0; a = 0.f; }
        if (i >= sn - 1) { i = sn - 1; a = 0.f; }
    }
    i0 = clampi(i, 0, sn - 1);
    i1 = clampi(i + 1, 0, sn - 1);
}

__device__ __forceinline__ float2 fb_upsampled_flow(const float2 *__restrict__ p, int ph, int pw, int x, int y, int lh, int lw)
{
    int x0, x1, y0, y1;
    float ax, ay;
    if (lw == 2 * pw && lh == 2 * ph) {                              // every level of a frame whose sides divide by 8
        up2_tap_f32(x, pw, false, x0, x1, ax);
        up2_tap_f32(y, ph, true, y0, y1, ay);
    } else {
        lin_tap_f32(x, pw, lw, false, x0, x1, ax);
        lin_tap_f32(y, ph, lh, true, y0, y1, ay);
    }
    const float a0 = 1.f - ax, b0 = 1.f - ay;
    const float2 p00 = __ldg(p + (size_t)y0 * pw + x0), p01 = __ldg(p + (size_t)y0 * pw + x1);
    const float2 p10 = __ldg(p + (size_t)y1 * pw + x0), p11 = __ldg(p + (size_t)y1 * pw + x1);
    float2 o;
    o.x = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p00.x, a0), __fmul_rn(p01.x, ax)), b0),
                    __fmul_rn(__fadd_rn(__fmul_rn(p10.x, a0), __fmul_rn(p11.x, ax)), ay)) * 2.f;
    o.y = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p00.y, a0), __fmul_rn(p01.y, ax)), b0),
                    __fmul_rn(__fadd_rn(__fmul_rn(p10.y, a0), __fmul_rn(p11.y, ax)), ay)) * 2.f;
    return o;
}

// FarnebackUpdateMatrices for one pixel: returns the 5 entries of M.  q = R of the previous frame at the pixel,
// R1_4 / R1s = R of the next frame (float4 + float arrays), f = the flow at the pixel.
__device__ __forceinline__ void fb_matrix_core(float4 q, float q4, const float4 *__restrict__ R1_4, const float *__restrict__ R1s,
                                               int x, int y, int h, int w, float2 f, float m[5])
{
    const float dx = f.x, dy = f.y;
    float fx = (float)x + dx, fy = (float)y + dy;
    const int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    fx -= (float)x1;
    fy -= (float)y1;
    float r2, r3, r4, r5, r6;
    if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
        const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        const int p = y1 * w + x1;
        const float4 t00 = __ldg(R1_4 + p), t01 = __ldg(R1_4 + p + 1), t10 = __ldg(R1_4 + p + w), t11 = __ldg(R1_4 + p + w + 1);
        const float s00 = __ldg(R1s + p), s01 = __ldg(R1s + p + 1), s10 = __ldg(R1s + p + w), s11 = __ldg(R1s + p + w + 1);
#define FB_TAP(v00, v01, v10, v11) (a00 * (v00) + a01 * (v01) + a10 * (v10) + a11 * (v11))
        r2 = FB_TAP(t00.x, t01.x, t10.x, t11.x);
        r3 = FB_TAP(t00.y, t01.y, t10.y, t11.y);
        r4 = FB_TAP(t00.z, t01.z, t10.z, t11.z);
        r5 = FB_TAP(t00.w, t01.w, t10.w, t11.w);
        r6 = FB_TAP(s00, s01, s10, s11);
#undef FB_TAP
        r4 = (q.z + r4) * 0.5f;
        r5 = (q.w + r5) * 0.5f;
        r6 = (q4 + r6) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = q.z;
        r5 = q.w;
        r6 = q4 * 0.5f;
    }
    r2 = (q.x - r2) * 0.5f;
    r3 = (q.y - r3) * 0.5f;
    r2 += r4 * dy + r6 * dx;
    r3 += r6 * dy + r5 * dx;
    if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
        // OpenCV's border[] = {0.14, 0.14, 0.4472, 0.4472, 0.4472} by distance to the edge
#define FB_BORDER(d) ((d) < 2 ? 0.14f : 0.4472f)
        const float sc = (x < 5 ? FB_BORDER(x) : 1.f) * (x >= w - 5 ? FB_BORDER(w - x - 1) : 1.f) *
                         (y < 5 ? FB_BORDER(y) : 1.f) * (y >= h - 5 ? FB_BORDER(h - y - 1) : 1.f);
#undef FB_BORDER
        r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
    }
    m[0] = r4 * r4 + r6 * r6;
    m[1] = (r4 + r5) * r6;
    m[2] = r5 * r5 + r6 * r6;
    m[3] = r4 * r2 + r6 * r3;
    m[4] = r6 * r2 + r5 * r3;
}

// UpdateMatrices, one pixel per lane (32 x 8 pixels per block).  INIT 0: flow read from memory; 1: flow up-sampled on
// the fly from the previous (coarser) level (first UpdateMatrices of a level; never stored); 2: zero flow (coarsest).
// Per pixel: R0 = 128-bit + 32-bit load, flow = 64-bit load, the four taps of R1 = 4 x (128-bit + 32-bit) loads, M =
// 128-bit + 32-bit store, every one coalesced over the 32 adjacent pixels of a warp.  (Round 1 kept R and M planar: 5 + 20
// scalar loads per pixel, a transposed 4-pixel variant to make them 128-bit, and both at 79-89 % of the L1 data pipe.)
// The pair is the FASTEST block index: the blocks of one image tile for consecutive pairs are scheduled together, and R
// of frame p+1 is both R1 of pair p and R0 of pair p+1 -- the second read hits L2 instead of DRAM (with the pair as the
// slowest index the two reads were one pair's 140 MB working set apart: -22 % DRAM bytes, profiles/r02_notes.md 5).
template <int INIT>
__global__ void __launch_bounds__(256)
k_fb_matrices(const float4 *__restrict__ R4, const float *__restrict__ Rs, const float2 *__restrict__ flow, int h, int w,
              float4 *__restrict__ M4, float *__restrict__ Ms, const float2 *__restrict__ prev, int ph, int pw)
{
    const int pair = blockIdx.x;
    const int x = blockIdx.y * 32 + (threadIdx.x & 31), y = blockIdx.z * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const size_t plane = (size_t)h * w, o = (size_t)pair * plane + (size_t)y * w + x;
    const float4 q = __ldg(R4 + o);
    const float q4 = __ldg(Rs + o);
    float2 f;
    if (INIT == 0) f = __ldg(flow + o);
    else if (INIT == 1) f = fb_upsampled_flow(prev + (size_t)pair * ph * pw, ph, pw, x, y, h, w);
    else f = make_float2(0.f, 0.f);
    float m[5];
    fb_matrix_core(q, q4, R4 + (size_t)(pair + 1) * plane, Rs + (size_t)(pair + 1) * plane, x, y, h, w, f, m);
    M4[o] = make_float4(m[0], m[1], m[2], m[3]);
    Ms[o] = m[4];
}

constexpr int MS_W = 128, MS_OUT = 112, MS_R = 7, MS_H = 96, MS_SEG = 8, MS_PF = 6;

// FarnebackUpdateFlow_Blur: 15x15 replicate-border box mean of the 5 planes of M, then the 2x2 solve.
// "Column marching": a block owns a strip of 112 output columns (+8 halo columns each side = 128
// threads, one column each) and walks down 64 rows.  Every thread keeps the vertical 15-row running
// sums of its column for the 5 planes in DOUBLE registers (OpenCV's vsum is double too: a float
// running sum would keep eps*|edge value| of error in flat areas next to strong edges), so each row
// costs one incoming + one outgoing load per plane.  The horizontal 15-tap sums are built from the
// shared row of vertical sums by 70 work items (5 planes x 14 segments of 8 outputs), sliding in double.
// Shared-memory layout of the row of vertical sums and of the horizontal sums: logical column i of a
// plane lives at word i + 4*(i >> 5) (four pad words after every 32).  With it all four access
// patterns are bank-conflict free: (A) thread t stores column t (an aligned group of 32 per warp),
// (B) work item (plane, segment s) loads columns 8s .. 8s+23 as six 128-bit words (a quarter-warp of
// eight segments covers all 32 banks exactly once), (C) it stores its eight sums as two 128-bit
// words, (D) thread t loads the sums of output column t - 8, stored at index t.  The earlier
// "one pad word per 8" layout had 2-way conflicts on (A), (B) and (D) (ncu: 45 % excessive wavefronts).
__device__ __forceinline__ int ms_sw(int i) { return i + ((i >> 5) << 2); }
constexpr int MS_VP = 144;

// Shared-memory accessors on 32-bit shared-window addresses held in registers.  The swizzled addresses of a thread are loop
// invariants, but written as C++ pointers the compiler re-derived them from threadIdx inside the row loop (shift / mask / add
// chains, ~25 of the kernel's 211 instructions per pixel, to stay at 56 registers); `opaque` hides their provenance so they are
// computed once and kept.
__device__ __forceinline__ uint32_t ms_opaque(uint32_t a) { asm volatile("" : "+r"(a)); return a; }
template <int OFF>
__device__ __forceinline__ void ms_sts(uint32_t a, float v) { asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(a), "n"(OFF), "f"(v) : "memory"); }
template <int OFF>
__device__ __forceinline__ float ms_lds(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF) : "memory");
    return v;
}
__device__ __forceinline__ float4 ms_lds128(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void ms_sts128(uint32_t a, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// FarnebackUpdateFlow_Blur, column-marching (see above).  On B200 this kernel is bound by warp
// instruction ISSUE (ncu: issue slots 78 % busy, 3 eligible warps per cycle; DRAM 46 %, L1 70 %), so
// the design minimises instructions per output, in this order of discovery (profiles/r01_notes.md):
//   * 32-bit shared rows in a conflict-free swizzle, horizontal phase on 128-bit accesses;
//   * the horizontal 15-tap sums are plain float sums (core + suffix + prefix, no sliding window, so
//     no error persists) with a dependent chain of 8;
//   * the 2x2 solve runs in float with FMA-recovered product errors (no conversions);
//   * M stored per pixel as float4 + float: a row entering / leaving the window is 2 loads instead of 5 (round 1 kept M
//     planar in 64 Ki-pixel super-chunks so that the ten loads came from two pointers);
//   * the vertical 15-row running sums are DOUBLE registers: 10 DADD + 15 conversions per row.  A plain
//     float running sum would keep eps*|edge value| of error in flat areas below strong edges (OpenCV
//     uses double here too); float-float pairs updated with TwoSum measured 8 % slower, evaluating the
//     next UpdateMatrices in the epilogue 7 % slower (profiles/r01_notes.md; both variants removed).
// The new flow is stored; on the last iteration of level 0 the caller may ask for sum |flow| instead.
// NEXT 1 (iterations 0 and 1 of a level): the new flow never leaves the registers -- UpdateMatrices of the NEXT iteration is
// pointwise in the pixel, so it is evaluated right here from R and the fresh flow and written to the other M buffer: the
// separate UpdateMatrices launch, the flow store and the flow load of that iteration disappear (round 1 measured this 7 %
// slower with planar R: 20 scalar gathers exposed once per row; with the float4 + float layout the taps are 8 loads).
template <int NEXT>
__global__ void __launch_bounds__(MS_W)
k_fb_blur_solve(const float4 *__restrict__ M4, const float *__restrict__ Ms, int h, int w, float2 *__restrict__ flow,
                int rows_per_block, double *__restrict__ mag_sum, int write_flow, const float4 *__restrict__ R4,
                const float *__restrict__ Rs, float4 *__restrict__ Mn4, float *__restrict__ Mns)
{
    __shared__ __align__(16) float row[5][MS_VP];
    __shared__ __align__(16) float hs[5][MS_VP];
    const int pair = blockIdx.z, t = threadIdx.x;
    const size_t plane = (size_t)h * w;
    const float4 *src4 = M4 + (size_t)pair * plane;
    const float *src1 = Ms + (size_t)pair * plane;
    const int sx0 = blockIdx.x * MS_OUT, y0 = blockIdx.y * rows_per_block;
    const int gx = clampi(sx0 - 8 + t, 0, w - 1);
    const int y_end = min(y0 + rows_per_block, h);
    double vd[5];
#pragma unroll
    for (int c = 0; c < 5; c++) vd[c] = 0.0;
    for (int k = -MS_R; k <= MS_R; k++) {
        const int p = clampi(y0 + k, 0, h - 1) * w + gx;
        const float4 v = __ldg(src4 + p);
        vd[0] += (double)v.x; vd[1] += (double)v.y; vd[2] += (double)v.z; vd[3] += (double)v.w;
        vd[4] += (double)__ldg(src1 + p);
    }
    // horizontal work item: 16 lanes per plane (14 segments of 8 outputs + 2 idle lanes), so that a
    // quarter-warp of a 128-bit access never mixes planes
    const int hc = t >> 4, hseg = t & 15;
    const bool hwork = t < 80 && hseg < 14;
    const int ox = t - 8, gxo = sx0 + ox;                           // output column of this thread
    const bool has_out = ox >= 0 && ox < MS_OUT && gxo < w;
    const uint32_t row_a = (uint32_t)__cvta_generic_to_shared(&row[0][0]), hs_a = (uint32_t)__cvta_generic_to_shared(&hs[0][0]);
    const uint32_t my_row = ms_opaque(row_a + 4u * ms_sw(t)), my_hs = ms_opaque(hs_a + 4u * ms_sw(t));
    const int hplane = (hwork ? hc : 0) * MS_VP, hcol = (hwork ? hseg : 0) * MS_SEG;
    uint32_t hin[6];
#pragma unroll
    for (int j = 0; j < 6; j++) hin[j] = ms_opaque(row_a + 4u * (hplane + ms_sw(hcol + 4 * j)));
    const uint32_t hout0 = ms_opaque(hs_a + 4u * (hplane + ms_sw(hcol + 8))), hout1 = ms_opaque(hs_a + 4u * (hplane + ms_sw(hcol + 12)));
    float2 *fout = NEXT ? nullptr : flow + (size_t)pair * plane + (size_t)y0 * w + (has_out ? gxo : 0);
    size_t onext = (size_t)pair * plane + (size_t)y0 * w + (has_out ? gxo : 0);     // NEXT: this thread's pixel in R0 / the next M
    // rows entering (yi) / leaving (yo) the 15-row window when the output row advances to y+1
    int yi = y0 + 1 + MS_R, yo = y0 - MS_R;
    unsigned p_in = (unsigned)(clampi(yi, 0, h - 1) * w + gx), p_out = (unsigned)(clampi(yo, 0, h - 1) * w + gx);   // linear pixel indices
    float nin[5], nout[5];
    double mag_acc = 0;
    for (int y = y0; y < y_end; y++) {
        const bool more = y + 1 < y_end;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        float q4 = 0.f;
        if (NEXT && has_out) {                                      // R of the previous frame at the output pixel: no dependence on the flow
            q = __ldg(R4 + onext);
            q4 = __ldg(Rs + onext);
        }
        if (MS_PF > 0 && y + MS_PF < y_end) {
            // the row that enters the window MS_PF rows from now -> L2 (its first touch: a DRAM access whose latency otherwise
            // sets the duration of a row iteration; the row that leaves was read 15 rows earlier and mostly still is in L2)
            const unsigned pf = (unsigned)(min(y + MS_PF + MS_R, h - 1) * w + gx);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(src4 + pf));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(src1 + pf));
        }
        if (more) {
            const float4 vi = __ldg(src4 + p_in), vo = __ldg(src4 + p_out);
            nin[0] = vi.x; nin[1] = vi.y; nin[2] = vi.z; nin[3] = vi.w; nin[4] = __ldg(src1 + p_in);
            nout[0] = vo.x; nout[1] = vo.y; nout[2] = vo.z; nout[3] = vo.w; nout[4] = __ldg(src1 + p_out);
            p_in += ((unsigned)yi < (unsigned)(h - 1)) ? (unsigned)w : 0u;     // replicate border: the row stops at 0 / h-1
            p_out += ((unsigned)yo < (unsigned)(h - 1)) ? (unsigned)w : 0u;
            yi++;
            yo++;
        }
        ms_sts<0>(my_row, (float)vd[0]);
        ms_sts<4 * MS_VP>(my_row, (float)vd[1]);
        ms_sts<8 * MS_VP>(my_row, (float)vd[2]);
        ms_sts<12 * MS_VP>(my_row, (float)vd[3]);
        ms_sts<16 * MS_VP>(my_row, (float)vd[4]);
        __syncthreads();
        if (hwork) {
            // p[k] = column 8s + k; output o of the segment sums columns 8s+o+1 .. 8s+o+15
            float p[24];
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const float4 v = ms_lds128(hin[j]);
                p[4 * j] = v.x; p[4 * j + 1] = v.y; p[4 * j + 2] = v.z; p[4 * j + 3] = v.w;
            }
            const float core = ((p[8] + p[9]) + (p[10] + p[11])) + ((p[12] + p[13]) + (p[14] + p[15]));
            float L[8], R[8];
            L[7] = 0.f;
#pragma unroll
            for (int j = 6; j >= 0; j--) L[j] = L[j + 1] + p[j + 1];
            R[0] = 0.f;
#pragma unroll
            for (int j = 1; j < 8; j++) R[j] = R[j - 1] + p[15 + j];
            float o8[MS_SEG];
#pragma unroll
            for (int j = 0; j < MS_SEG; j++) o8[j] = (core + L[j]) + R[j];
            ms_sts128(hout0, make_float4(o8[0], o8[1], o8[2], o8[3]));
            ms_sts128(hout1, make_float4(o8[4], o8[5], o8[6], o8[7]));
        }
        __syncthreads();
        if (has_out) {
            // 2x2 solve on the raw window sums s.. (g.. = s.. / 225): OpenCV evaluates
            // det = g11*g22 - g12^2 + 1e-3 and the two numerators in double; here every a*b - c*d is
            // formed in float with the product error recovered by FMA (w = c*d, e = fma(-c, d, w),
            // a*b - c*d = fma(a, b, -w) + e: below 2 ulp even under cancellation), so the XU pipe sees
            // one reciprocal instead of ten float<->double conversions per output
            const float s11 = ms_lds<0>(my_hs), s12 = ms_lds<4 * MS_VP>(my_hs), s22 = ms_lds<8 * MS_VP>(my_hs),
                        t1 = ms_lds<12 * MS_VP>(my_hs), t2 = ms_lds<16 * MS_VP>(my_hs);
            const float k2 = 1.f / (225.f * 225.f);
            const float w0 = __fmul_rn(s12, s12), w1 = __fmul_rn(s12, t1), w2 = __fmul_rn(s12, t2);
            const float det = __fmaf_rn(__fadd_rn(__fmaf_rn(s11, s22, -w0), __fmaf_rn(-s12, s12, w0)), k2, 1e-3f);
            const float nx = __fadd_rn(__fmaf_rn(s11, t2, -w1), __fmaf_rn(-s12, t1, w1));
            const float ny = __fadd_rn(__fmaf_rn(s22, t1, -w2), __fmaf_rn(-s12, t2, w2));
            float idet = __frcp_rn(det);
            float2 o;
            o.x = __fmul_rn(__fmul_rn(nx, k2), idet);
            o.y = __fmul_rn(__fmul_rn(ny, k2), idet);
            if (NEXT) {
                float m[5];
                fb_matrix_core(q, q4, R4 + (size_t)(pair + 1) * plane, Rs + (size_t)(pair + 1) * plane, gxo, y, h, w, o, m);
                Mn4[onext] = make_float4(m[0], m[1], m[2], m[3]);
                Mns[onext] = m[4];
            } else {
                if (write_flow) *fout = o;
                // last iteration of level 0: the flow field itself is not needed any more, only
                // sum |flow| (cartToPolar magnitude, complexity_metrics.py:342-343)
                if (mag_sum) mag_acc += (double)sqrtf(__fadd_rn(__fmul_rn(o.x, o.x), __fmul_rn(o.y, o.y)));
            }
        }
        if (NEXT) onext += w;
        else fout += w;
        if (more) {
#pragma unroll
            for (int c = 0; c < 5; c++) vd[c] = (vd[c] + (double)nin[c]) - (double)nout[c];
        }
    }
    if (!NEXT && mag_sum) {
        __shared__ double red[MS_W / 32];
        mag_acc = warp_sum(mag_acc);
        __syncthreads();
        if ((t & 31) == 0) red[t >> 5] = mag_acc;
        __syncthreads();
        if (t == 0) {
            double sum = 0;
            for (int i = 0; i < MS_W / 32; i++) sum += red[i];
            atomicAdd(&mag_sum[pair], sum);
        }
    }
}

// ---------------------------------------------------------------------------------- host side
static void make_gauss(int ksz, double sigma, GaussTaps &t)
{
    t.ksz = ksz;
    memset(t.k, 0, sizeof(t.k));
    if (sigma <= 0 && ksz == 3) { t.k[0] = 0.25f; t.k[1] = 0.5f; t.k[2] = 0.25f; return; }
    const double s = sigma > 0 ? sigma : ((ksz - 1) * 0.5 - 1) * 0.3 + 0.8, sc = -0.5 / (s * s);
    double v[32], sum = 0;
    for (int i = 0; i < ksz; i++) { double x = i - (ksz - 1) * 0.5; v[i] = exp(sc * x * x); sum += v[i]; }
    for (int i = 0; i < ksz; i++) t.k[i] = (float)(v[i] / sum);
}

static void make_poly(PolyConst &pc)
{
    const int n = 5;
    double sigma = 1.2, s = 0;
    float *g = pc.g + n, *xg = pc.xg + n, *xxg = pc.xxg + n;
    for (int x = -n; x <= n; x++) { g[x] = (float)exp(-x * x / (2 * sigma * sigma)); s += g[x]; }
    s = 1. / s;
    for (int x = -n; x <= n; x++) {
        g[x] = (float)(g[x] * s);
        xg[x] = (float)(x * g[x]);
        xxg[x] = (float)(x * x * g[x]);
    }
    double G[6][6] = {{0}};
    for (int y = -n; y <= n; y++)
        for (int x = -n; x <= n; x++) {
            const float gg = g[y] * g[x];
            G[0][0] += gg; G[1][1] += gg * x * x; G[3][3] += gg * x * x * x * x; G[5][5] += gg * x * x * y * y;
        }
    G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
    G[4][4] = G[3][3];
    G[3][4] = G[4][3] = G[5][5];
    double m[6][12];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 12; j++) m[i][j] = j < 6 ? G[i][j] : (j - 6 == i ? 1.0 : 0.0);
    for (int c = 0; c < 6; c++) {
        int p = c;
        for (int r = c + 1; r < 6; r++) if (fabs(m[r][c]) > fabs(m[p][c])) p = r;
        if (p != c) for (int j = 0; j < 12; j++) { double t = m[c][j]; m[c][j] = m[p][j]; m[p][j] = t; }
        const double d = 1.0 / m[c][c];
        for (int j = 0; j < 12; j++) m[c][j] *= d;
        for (int r = 0; r < 6; r++)
            if (r != c) {
                const double f = m[r][c];
                if (f != 0) for (int j = 0; j < 12; j++) m[r][j] -= f * m[c][j];
            }
    }
    pc.ig11 = m[1][7]; pc.ig03 = m[0][9]; pc.ig33 = m[3][9]; pc.ig55 = m[5][11];
}

constexpr bool FB_DUAL_STREAMS = false;            // two-stream staggered schedule of a level (see run_farneback)
constexpr bool FB_FUSE_NEXT = false;               // iterations 0 and 1 of a level: the blur evaluates the next UpdateMatrices itself

int run_farneback(vqa_ctx *c, const uint8_t *gray, int npairs, int h, int w, double *mag_sum, float *flow_out,
                  const std::function<int()> *level0_hook)
{
    if (npairs <= 0) return VQA_OK;
    const int nf = npairs + 1;
    int levels = 0;
    {
        double scale = 1;
        for (levels = 0; levels < 3; levels++) {
            scale *= 0.5;
            if (w * scale < 32 || h * scale < 32) break;
        }
    }
    const size_t full = (size_t)h * w;
    VQA_BUF(c, I, float, "fb.I", full * nf);
    // R (per frame) and M (per pair): per pixel one float4 (components 0..3) in the first 4/5 of the buffer and one float
    // (component 4) behind it; frames / pairs are `level pixels` apart inside each part
    VQA_BUF(c, Rbuf, float, "fb.R", full * 5 * nf);
    VQA_BUF(c, Mbuf, float, "fb.M", full * 5 * npairs);
    float4 *R4 = reinterpret_cast<float4 *>(Rbuf), *M4 = reinterpret_cast<float4 *>(Mbuf), *N4 = nullptr;
    float *Rs = Rbuf + full * 4 * nf, *Ms = Mbuf + full * 4 * npairs, *Ns = nullptr;
    VQA_BUF(c, flowA, float2, "fb.flowA", full * npairs);
    VQA_BUF(c, flowB, float2, "fb.flowB", full * npairs);
#ifdef VQA_AB
    // development build only (the product build reads no environment): strip-height, pair-group and two-stream knobs
    const int ms_h_cap = (getenv("VQA_MS_H") && atoi(getenv("VQA_MS_H")) >= 16) ? atoi(getenv("VQA_MS_H")) : MS_H;
    const int group = (getenv("VQA_FB_GROUP") && atoi(getenv("VQA_FB_GROUP")) >= 1) ? std::min(atoi(getenv("VQA_FB_GROUP")), npairs) : npairs;
    const bool dual = getenv("VQA_FB_DUAL") ? atoi(getenv("VQA_FB_DUAL")) != 0 : FB_DUAL_STREAMS;
    const bool fuse_next = getenv("VQA_FB_NEXT") ? atoi(getenv("VQA_FB_NEXT")) != 0 : FB_FUSE_NEXT;
#else
    constexpr bool fuse_next = FB_FUSE_NEXT;
    constexpr bool dual = FB_DUAL_STREAMS;
    constexpr int ms_h_cap = MS_H;
    const int group = npairs;
#endif
    if (fuse_next) {                                                 // the blur that also evaluates the next UpdateMatrices writes the other M buffer
        VQA_BUF(c, Mbuf2, float, "fb.M2", full * 5 * npairs);
        N4 = reinterpret_cast<float4 *>(Mbuf2);
        Ns = Mbuf2 + full * 4 * npairs;
    }
    PolyConst pc;
    make_poly(pc);
    VQA_CUDA(c, cudaMemsetAsync(mag_sum, 0, sizeof(double) * (size_t)npairs, c->stream));
    float2 *flow = flowA, *prev = flowB;
    int ph = 0, pw = 0;
    for (int k = levels; k >= 0; k--) {
        double scale = 1;
        for (int i = 0; i < k; i++) scale *= 0.5;
        const double sigma = (1. / scale - 1) * 0.5;
        int ksz = (int)nearbyint(sigma * 5) | 1;
        if (ksz < 3) ksz = 3;
        const int lw = (int)nearbyint(w * scale), lh = (int)nearbyint(h * scale);
        GaussTaps taps;
        make_gauss(ksz, sigma, taps);
        const int mode = (lw == w && lh == h) ? 0 : ((w == 2 * lw && h == 2 * lh) ? 1 : 2);
        dim3 gF(cdiv(lw, PY_TW), cdiv(lh, PY_TH), nf);
        {
            // worst-case source region of a 32x8 output tile (+ Gaussian radius), for the dynamic smem size
            const double sx = (double)w / lw, sy = (double)h / lh;
            const int rwp = (((int)ceil(PY_TW * sx) + 2 * (ksz / 2) + 4 + 3) + 3) & ~3;   // + 3: region start aligned down to 4
            const int rhm = (int)ceil(PY_TH * sy) + 2 * (ksz / 2) + 4;
            const size_t smem = (((size_t)rhm * rwp + 15) & ~(size_t)15) + (size_t)rhm * 2 * PY_TW * sizeof(float);
            if (smem > 200 * 1024) return set_err(c, VQA_E_UNSUPPORTED, "farneback pyramid tile needs %zu B of shared memory", smem);
            if (smem > 48 * 1024) {
                VQA_CUDA(c, cudaFuncSetAttribute(k_fb_pyramid<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                VQA_CUDA(c, cudaFuncSetAttribute(k_fb_pyramid<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                VQA_CUDA(c, cudaFuncSetAttribute(k_fb_pyramid<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            }
            VQA_BYTES(c, ((double)full + 4.0 * lw * lh) * nf);
            if (ksz == 3 && mode == 0) {
                VQA_LAUNCH(c, k_fb_pyramid3<0>, dim3(cdiv(w, 128), cdiv(h, 8), nf), 256, 0, gray, h, w, lh, lw, taps.k[0], taps.k[1], I);
            } else if (ksz == 3 && mode == 1) {
                VQA_LAUNCH(c, k_fb_pyramid3<1>, dim3(cdiv(w, 128), cdiv(h, 16), nf), 256, 0, gray, h, w, lh, lw, taps.k[0], taps.k[1], I);
            } else if (ksz == 9 && w == 4 * lw && h == 4 * lh) {
                if (int rc = launch_pyramid_dec<4, 4, 32>(c, gray, h, w, lh, lw, nf, taps, I)) return rc;
            } else if (ksz == 19 && w == 8 * lw && h == 8 * lh) {
                if (int rc = launch_pyramid_dec<8, 9, 16>(c, gray, h, w, lh, lw, nf, taps, I)) return rc;
            } else {
                if (ksz == 19) VQA_LAUNCH(c, k_fb_pyramid<9>, gF, PY_TW * PY_TH, smem, gray, h, w, lh, lw, mode, taps, I, rwp);
                else if (ksz == 9) VQA_LAUNCH(c, k_fb_pyramid<4>, gF, PY_TW * PY_TH, smem, gray, h, w, lh, lw, mode, taps, I, rwp);
                else VQA_LAUNCH(c, k_fb_pyramid<0>, gF, PY_TW * PY_TH, smem, gray, h, w, lh, lw, mode, taps, I, rwp);
            }
        }
        VQA_BYTES(c, 24.0 * lw * lh * nf);
        const dim3 gE(cdiv(lw, PE_TW), cdiv(cdiv(lh, PE_TH), PE_STRIP), nf);
        VQA_CUDA(c, cudaFuncSetAttribute(k_fb_polyexp, cudaFuncAttributeMaxDynamicSharedMemorySize, PE_SMEM));
        VQA_LAUNCH(c, k_fb_polyexp, gE, 256, PE_SMEM, I, lh, lw, pc, R4, Rs);
        // One group of pairs through the three iterations of the level, on the context's current stream.  `after_first`
        // is recorded behind the group's first kernel (the stagger of the two-stream schedule below).
        if (k == 0 && level0_hook)
            if (int rc = (*level0_hook)()) return rc;
        const size_t lpx = (size_t)lw * lh, ppx = (size_t)pw * ph;
        auto run_group = [&](int g0, int gn, size_t m0, cudaEvent_t after_first) -> int {
            // frame g0 of R, pair m0 of M (M is scratch: groups of the sequential schedule reuse pair 0 onwards), pair g0 of the flows
            const float4 *Rg4 = R4 + (size_t)g0 * lpx;
            const float *Rgs = Rs + (size_t)g0 * lpx;
            float4 *Mg4 = M4 + m0 * lpx;
            float *Mgs = Ms + m0 * lpx;
            float2 *fg = flow + (size_t)g0 * lpx;
            const float2 *pg = prev + (size_t)g0 * ppx;
            const dim3 gPg(gn, cdiv(lw, 32), cdiv(lh, 8));
            // first UpdateMatrices of the level: zero flow at the coarsest level, else the coarser flow up-sampled on the fly
            if (k == levels) {
                VQA_BYTES(c, 60.0 * lpx * gn);
                VQA_LAUNCH(c, k_fb_matrices<2>, gPg, 256, 0, Rg4, Rgs, fg, lh, lw, Mg4, Mgs, pg, ph, pw);
            } else {
                VQA_BYTES(c, (60.0 * lpx + 8.0 * ppx) * gn);
                VQA_LAUNCH(c, k_fb_matrices<1>, gPg, 256, 0, Rg4, Rgs, fg, lh, lw, Mg4, Mgs, pg, ph, pw);
            }
            if (after_first) VQA_CUDA(c, cudaEventRecord(after_first, c->stream));
            // rows per block of the marching blur: long strips amortise the 15-row warm-up, but the small
            // pyramid levels need shorter strips to put >= ~3 waves of blocks on 148 SMs x 8 blocks
            int rows_pb = ms_h_cap;
            {
                const long want = 3L * c->sm_count * 8, per_row_strip = (long)cdiv(lw, MS_OUT) * gn;
                const long strips = (want + per_row_strip - 1) / per_row_strip;
                if (strips > 0) rows_pb = (int)((lh + strips - 1) / strips);
                if (rows_pb < 16) rows_pb = 16;
                if (rows_pb > ms_h_cap) rows_pb = ms_h_cap;
            }
            const dim3 gB(cdiv(lw, MS_OUT), cdiv(lh, rows_pb), gn);
#ifdef VQA_AB
            float4 *Ng4 = N4 + m0 * lpx;
            float *Ngs = Ns + m0 * lpx;
#endif
            for (int it = 0; it < 3; it++) {
                const bool last = (k == 0 && it == 2);
                double *ms = last ? mag_sum + g0 : (double *)nullptr;
                const int wf = (!last || flow_out) ? 1 : 0;
#ifdef VQA_AB                                                        // the losing variant is compiled into the development build only
                if (it < 2 && fuse_next) {
                    // blur + solve + the next iteration's UpdateMatrices: M 20 + R0 20 + R1 20 read, M' 20 written
                    VQA_BYTES(c, 80.0 * lpx * gn);
                    VQA_LAUNCH(c, k_fb_blur_solve<1>, gB, MS_W, 0, Mg4, Mgs, lh, lw, fg, rows_pb, (double *)nullptr, 0, Rg4, Rgs, Ng4, Ngs);
                    std::swap(Mg4, Ng4);
                    std::swap(Mgs, Ngs);
                    continue;
                }
#endif
                VQA_BYTES(c, 28.0 * lpx * gn);
                VQA_LAUNCH(c, k_fb_blur_solve<0>, gB, MS_W, 0, Mg4, Mgs, lh, lw, fg, rows_pb, ms, wf, (const float4 *)nullptr,
                           (const float *)nullptr, (float4 *)nullptr, (float *)nullptr);
                if (it < 2) {
                    VQA_BYTES(c, 68.0 * lpx * gn);
                    VQA_LAUNCH(c, k_fb_matrices<0>, gPg, 256, 0, Rg4, Rgs, fg, lh, lw, Mg4, Mgs, pg, ph, pw);
                }
            }
            return VQA_OK;
        };
        if (dual && npairs >= 8 && !c->ktiming) {
            // Two-stream schedule: the chain alternates a DRAM-bound kernel (UpdateMatrices, 0.75-0.84 of the copy peak, issue
            // slots half idle) with an issue-bound one (the marching blur, DRAM half idle).  The pairs are cut in two halves on
            // two streams, the second ONE KERNEL BEHIND the first, so that a blur of one half always runs next to an
            // UpdateMatrices of the other and the two share the SMs instead of taking turns.
            const int ga = (npairs + 1) / 2, gb = npairs - ga;
            cudaStream_t main_stream = c->stream;
            VQA_CUDA(c, cudaEventRecord(c->ev_fb_fork, main_stream));
            VQA_CUDA(c, cudaStreamWaitEvent(c->fb_stream, c->ev_fb_fork, 0));
            if (int rc = run_group(0, ga, 0, c->ev_fb_stagger)) return rc;
            {
                c->stream = c->fb_stream;
                cudaError_t e = cudaStreamWaitEvent(c->fb_stream, c->ev_fb_stagger, 0);
                int rc = e == cudaSuccess ? run_group(ga, gb, (size_t)ga, nullptr) : VQA_E_CUDA;
                if (rc == VQA_OK && cudaEventRecord(c->ev_fb_join, c->fb_stream) != cudaSuccess) rc = VQA_E_CUDA;
                c->stream = main_stream;
                if (rc) return rc == VQA_E_CUDA ? set_err(c, VQA_E_CUDA, "two-stream Farneback schedule: %s", cudaGetErrorString(cudaGetLastError())) : rc;
            }
            VQA_CUDA(c, cudaStreamWaitEvent(main_stream, c->ev_fb_join, 0));
        } else {
            // the pairs of the chunk go through the level group by group (product build: ONE group; the development build can
            // cut the chunk into groups whose M field stays in L2 between its writer and its reader: measured negative,
            // profiles/r02_notes.md 3)
            for (int g0 = 0; g0 < npairs; g0 += group)
                if (int rc = run_group(g0, std::min(group, npairs - g0), 0, nullptr)) return rc;
        }
        float2 *t = prev; prev = flow; flow = t;
        ph = lh; pw = lw;
    }
    // `prev` now holds the level-0 flow (only written when the caller asked for it); sum |flow| was
    // accumulated by the last blur pass
    if (flow_out)
        VQA_CUDA(c, cudaMemcpyAsync(flow_out, prev, sizeof(float2) * full * npairs, cudaMemcpyDeviceToDevice, c->stream));
    return VQA_OK;
}

}  // namespace vqa
