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
//   k_fb_mag_sum        sum sqrt(fx^2 + fy^2) at level 0
//
// R is computed once per frame and serves the frame both as "next" of one pair and as "prev"
// of the following pair.  R, M and flow stay fp32 (reduced precision does not survive flat
// content).  Roofline: HBM.
#include <math.h>

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

// horizontal Gaussian at (row y, column x) of the uint8 image, symmetric evaluation order.
// Interior pixels (the whole footprint inside the row) skip the reflect-101 index arithmetic.
__device__ __forceinline__ float hblur_u8(const uint8_t *__restrict__ row, int x, int w, const GaussTaps &t)
{
    const int r = t.ksz >> 1;
    float s = t.k[r] * (float)__ldg(row + x);
    if (x - r >= 0 && x + r < w) {
        for (int i = 1; i <= r; i++) s += t.k[r + i] * ((float)__ldg(row + x - i) + (float)__ldg(row + x + i));
    } else {
        for (int i = 1; i <= r; i++)
            s += t.k[r + i] * ((float)__ldg(row + reflect101(x - i, w)) + (float)__ldg(row + reflect101(x + i, w)));
    }
    return s;
}

__device__ __forceinline__ float blur_at(const uint8_t *__restrict__ img, int x, int y, int h, int w, const GaussTaps &t)
{
    const int r = t.ksz >> 1;
    float s = t.k[r] * hblur_u8(img + (size_t)y * w, x, w, t);
    if (y - r >= 0 && y + r < h) {
        for (int j = 1; j <= r; j++)
            s += t.k[r + j] * (hblur_u8(img + (size_t)(y - j) * w, x, w, t) + hblur_u8(img + (size_t)(y + j) * w, x, w, t));
    } else {
        for (int j = 1; j <= r; j++)
            s += t.k[r + j] * (hblur_u8(img + (size_t)reflect101(y - j, h) * w, x, w, t) +
                               hblur_u8(img + (size_t)reflect101(y + j, h) * w, x, w, t));
    }
    return s;
}

__device__ __forceinline__ void lin_tap_f32(int d, int sn, int dn, bool vertical, int &i0, int &i1, float &a)
{
    const double scale = (double)sn / (double)dn;
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

// mode 0: same size; 1: exact 2x decimation (INTER_AREA fast path); 2: bilinear
__global__ void __launch_bounds__(256)
k_fb_pyramid(const uint8_t *__restrict__ gray, int H, int W, int lh, int lw, int mode, GaussTaps taps,
             float *__restrict__ I)
{
    const int frame = blockIdx.z;
    const uint8_t *img = gray + (size_t)frame * H * W;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= lw || y >= lh) return;
    float v;
    if (mode == 0) {
        v = blur_at(img, x, y, H, W, taps);
    } else if (mode == 1) {
        v = (blur_at(img, 2 * x, 2 * y, H, W, taps) + blur_at(img, 2 * x + 1, 2 * y, H, W, taps) +
             blur_at(img, 2 * x, 2 * y + 1, H, W, taps) + blur_at(img, 2 * x + 1, 2 * y + 1, H, W, taps)) * 0.25f;
    } else {
        int x0, x1, y0, y1;
        float ax, ay;
        lin_tap_f32(x, W, lw, false, x0, x1, ax);
        lin_tap_f32(y, H, lh, true, y0, y1, ay);
        const float a0 = 1.f - ax, b0 = 1.f - ay;
        float t0 = __fadd_rn(__fmul_rn(blur_at(img, x0, y0, H, W, taps), a0), __fmul_rn(blur_at(img, x1, y0, H, W, taps), ax));
        float t1 = __fadd_rn(__fmul_rn(blur_at(img, x0, y1, H, W, taps), a0), __fmul_rn(blur_at(img, x1, y1, H, W, taps), ax));
        v = __fadd_rn(__fmul_rn(t0, b0), __fmul_rn(t1, ay));
    }
    I[(size_t)frame * lh * lw + (size_t)y * lw + x] = v;
}

constexpr int PE_TW = 64, PE_TH = 16, PE_R = 5;

// FarnebackPolyExp: vertical pass in float, horizontal pass with double accumulators.
__global__ void __launch_bounds__(256)
k_fb_polyexp(const float *__restrict__ I, int h, int w, PolyConst pc, float *__restrict__ R)
{
    __shared__ float tile[PE_TH + 2 * PE_R][PE_TW + 2 * PE_R];
    __shared__ float v0[PE_TH][PE_TW + 2 * PE_R], v1[PE_TH][PE_TW + 2 * PE_R], v2[PE_TH][PE_TW + 2 * PE_R];
    const int frame = blockIdx.z;
    const float *src = I + (size_t)frame * h * w;
    const int tx0 = blockIdx.x * PE_TW, ty0 = blockIdx.y * PE_TH;
    for (int i = threadIdx.x; i < (PE_TH + 2 * PE_R) * (PE_TW + 2 * PE_R); i += 256) {
        const int y = i / (PE_TW + 2 * PE_R), x = i - y * (PE_TW + 2 * PE_R);
        const int gy = clampi(ty0 - PE_R + y, 0, h - 1), gx = clampi(tx0 - PE_R + x, 0, w - 1);
        tile[y][x] = src[(size_t)gy * w + gx];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PE_TH * (PE_TW + 2 * PE_R); i += 256) {
        const int y = i / (PE_TW + 2 * PE_R), x = i - y * (PE_TW + 2 * PE_R);
        // rows are clamped to the IMAGE (replicate), which the clamped tile load reproduces only if the
        // tile row index maps to the clamped image row: true because the load clamps gy itself.
        float t0 = tile[y + PE_R][x] * pc.g[PE_R], t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int k = 1; k <= PE_R; k++) {
            const float a = tile[y + PE_R - k][x], b = tile[y + PE_R + k][x];
            const float p = a + b;
            t0 = t0 + pc.g[PE_R + k] * p;
            t1 = t1 + pc.xg[PE_R + k] * (b - a);
            t2 = t2 + pc.xxg[PE_R + k] * p;
        }
        v0[y][x] = t0; v1[y][x] = t1; v2[y][x] = t2;
    }
    __syncthreads();
    const size_t plane = (size_t)h * w;
    float *dst = R + (size_t)frame * 5 * plane;
    for (int i = threadIdx.x; i < PE_TH * PE_TW; i += 256) {
        const int y = i / PE_TW, x = i - y * PE_TW;
        const int gy = ty0 + y, gx = tx0 + x;
        if (gy >= h || gx >= w) continue;
        const int c = x + PE_R;
        const double g0 = pc.g[PE_R];
        double b1 = v0[y][c] * g0, b2 = 0, b3 = v1[y][c] * g0, b4 = 0, b5 = v2[y][c] * g0, b6 = 0;
#pragma unroll
        for (int k = 1; k <= PE_R; k++) {
            const double gk = pc.g[PE_R + k], xgk = pc.xg[PE_R + k], xxgk = pc.xxg[PE_R + k];
            const double tg = (double)(v0[y][c + k] + v0[y][c - k]);
            b1 += tg * gk;
            b4 += tg * xxgk;
            b2 += (double)(v0[y][c + k] - v0[y][c - k]) * xgk;
            b3 += (double)(v1[y][c + k] + v1[y][c - k]) * gk;
            b6 += (double)(v1[y][c + k] - v1[y][c - k]) * xgk;
            b5 += (double)(v2[y][c + k] + v2[y][c - k]) * gk;
        }
        const size_t o = (size_t)gy * w + gx;
        dst[o] = (float)(b3 * pc.ig11);
        dst[plane + o] = (float)(b2 * pc.ig11);
        dst[2 * plane + o] = (float)(b1 * pc.ig03 + b5 * pc.ig33);
        dst[3 * plane + o] = (float)(b1 * pc.ig03 + b4 * pc.ig33);
        dst[4 * plane + o] = (float)(b6 * pc.ig55);
    }
}

// flow_k = 2 * resize_f32(flow_{k+1}) (INTER_LINEAR upscale, OpenCV float tap rules)
__global__ void __launch_bounds__(256)
k_fb_upsample(const float2 *__restrict__ prev, int ph, int pw, float2 *__restrict__ flow, int lh, int lw)
{
    const int pair = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= lw || y >= lh) return;
    const float2 *p = prev + (size_t)pair * ph * pw;
    int x0, x1, y0, y1;
    float ax, ay;
    lin_tap_f32(x, pw, lw, false, x0, x1, ax);
    lin_tap_f32(y, ph, lh, true, y0, y1, ay);
    const float a0 = 1.f - ax, b0 = 1.f - ay;
    const float2 p00 = p[(size_t)y0 * pw + x0], p01 = p[(size_t)y0 * pw + x1];
    const float2 p10 = p[(size_t)y1 * pw + x0], p11 = p[(size_t)y1 * pw + x1];
    float2 o;
    o.x = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p00.x, a0), __fmul_rn(p01.x, ax)), b0),
                    __fmul_rn(__fadd_rn(__fmul_rn(p10.x, a0), __fmul_rn(p11.x, ax)), ay)) * 2.f;
    o.y = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p00.y, a0), __fmul_rn(p01.y, ax)), b0),
                    __fmul_rn(__fadd_rn(__fmul_rn(p10.y, a0), __fmul_rn(p11.y, ax)), ay)) * 2.f;
    flow[(size_t)pair * lh * lw + (size_t)y * lw + x] = o;
}

// FarnebackUpdateMatrices for one pixel: returns the 5 entries of M
__device__ __forceinline__ void fb_matrix_at(const float *__restrict__ R0, const float *__restrict__ R1, size_t plane,
                                             int x, int y, int h, int w, float2 f, float m[5])
{
    const size_t o = (size_t)y * w + x;
    const float dx = f.x, dy = f.y;
    float fx = (float)x + dx, fy = (float)y + dy;
    const int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    fx -= (float)x1;
    fy -= (float)y1;
    float r2, r3, r4, r5, r6;
    const float q0 = R0[o], q1 = R0[plane + o], q2 = R0[2 * plane + o], q3 = R0[3 * plane + o], q4 = R0[4 * plane + o];
    if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
        const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        const size_t p = (size_t)y1 * w + x1;
#define FB_TAP(c) (a00 * R1[(c) * plane + p] + a01 * R1[(c) * plane + p + 1] + a10 * R1[(c) * plane + p + w] + a11 * R1[(c) * plane + p + w + 1])
        r2 = FB_TAP(0);
        r3 = FB_TAP(1);
        r4 = FB_TAP(2);
        r5 = FB_TAP(3);
        r6 = FB_TAP(4);
#undef FB_TAP
        r4 = (q2 + r4) * 0.5f;
        r5 = (q3 + r5) * 0.5f;
        r6 = (q4 + r6) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = q2;
        r5 = q3;
        r6 = q4 * 0.5f;
    }
    r2 = (q0 - r2) * 0.5f;
    r3 = (q1 - r3) * 0.5f;
    r2 += r4 * dy + r6 * dx;
    r3 += r6 * dy + r5 * dx;
    if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
        const float border[5] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
        const float sc = (x < 5 ? border[x] : 1.f) * (x >= w - 5 ? border[w - x - 1] : 1.f) *
                         (y < 5 ? border[y] : 1.f) * (y >= h - 5 ? border[h - y - 1] : 1.f);
        r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
    }
    m[0] = r4 * r4 + r6 * r6;
    m[1] = (r4 + r5) * r6;
    m[2] = r5 * r5 + r6 * r6;
    m[3] = r4 * r2 + r6 * r3;
    m[4] = r6 * r2 + r5 * r3;
}

__global__ void __launch_bounds__(256)
k_fb_matrices(const float *__restrict__ R, const float2 *__restrict__ flow, int h, int w, float *__restrict__ M)
{
    const int pair = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const size_t plane = (size_t)h * w;
    const float *R0 = R + (size_t)pair * 5 * plane, *R1 = R0 + 5 * plane;
    float m[5];
    fb_matrix_at(R0, R1, plane, x, y, h, w, flow[(size_t)pair * plane + (size_t)y * w + x], m);
    float *dst = M + (size_t)pair * 5 * plane + (size_t)y * w + x;
#pragma unroll
    for (int c = 0; c < 5; c++) dst[c * plane] = m[c];
}

constexpr int BS_TW = 32, BS_TH = 32, BS_R = 7;
constexpr int BS_PW = BS_TW + 2 * BS_R;          // padded tile width (46)

// FarnebackUpdateFlow_Blur: 15x15 replicate-border box mean of the 5 planes of M, then the 2x2 solve.
// Both box passes use register sliding windows: a work item sums 15 taps once and slides (7 steps
// vertically, 3 horizontally), ~5 shared-memory reads per output instead of 30.
__global__ void __launch_bounds__(256)
k_fb_blur_solve(const float *__restrict__ M, int h, int w, float2 *__restrict__ flow)
{
    __shared__ float tile[BS_TH + 2 * BS_R][BS_PW + 1];
    __shared__ float vs[BS_TH][BS_PW + 1];
    const int pair = blockIdx.z;
    const size_t plane = (size_t)h * w;
    const float *src = M + (size_t)pair * 5 * plane;
    const int tx0 = blockIdx.x * BS_TW, ty0 = blockIdx.y * BS_TH;
    const int oy = threadIdx.x >> 3, ox = (threadIdx.x & 7) * 4;   // each thread: row oy, 4 columns from ox
    float g[5][4];
#pragma unroll
    for (int c = 0; c < 5; c++) {
        const float *pl = src + c * plane;
        for (int i = threadIdx.x; i < (BS_TH + 2 * BS_R) * BS_PW; i += 256) {
            const int y = i / BS_PW, x = i - y * BS_PW;
            const int gy = clampi(ty0 - BS_R + y, 0, h - 1), gx = clampi(tx0 - BS_R + x, 0, w - 1);
            tile[y][x] = __ldg(pl + (size_t)gy * w + gx);
        }
        __syncthreads();
        if (threadIdx.x < 4 * BS_PW) {                              // vertical: column x, 8 rows from y0
            const int x = threadIdx.x % BS_PW, y0 = (threadIdx.x / BS_PW) * 8;
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 2 * BS_R + 1; k++) s += tile[y0 + k][x];
            vs[y0][x] = s;
#pragma unroll
            for (int r = 1; r < 8; r++) {
                s += tile[y0 + r + 2 * BS_R][x] - tile[y0 + r - 1][x];
                vs[y0 + r][x] = s;
            }
        }
        __syncthreads();
        {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 2 * BS_R + 1; k++) s += vs[oy][ox + k];
            g[c][0] = s;
#pragma unroll
            for (int j = 1; j < 4; j++) {
                s += vs[oy][ox + j + 2 * BS_R] - vs[oy][ox + j - 1];
                g[c][j] = s;
            }
        }
        __syncthreads();
    }
    const double scale = 1.0 / 225.0;
    const int gy = ty0 + oy;
    if (gy >= h) return;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int gx = tx0 + ox + j;
        if (gx >= w) continue;
        const double g11 = g[0][j] * scale, g12 = g[1][j] * scale, g22 = g[2][j] * scale;
        const double h1 = g[3][j] * scale, h2 = g[4][j] * scale;
        const double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
        float2 o;
        o.x = (float)((g11 * h2 - g12 * h1) * idet);
        o.y = (float)((g22 * h1 - g12 * h2) * idet);
        flow[(size_t)pair * plane + (size_t)gy * w + gx] = o;
    }
}

__global__ void __launch_bounds__(256)
k_fb_mag_sum(const float2 *__restrict__ flow, long plane, double *__restrict__ out)
{
    __shared__ double red[8];
    const int pair = blockIdx.y;
    const float2 *f = flow + (size_t)pair * plane;
    double acc = 0;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < plane; i += (long)gridDim.x * 256) {
        const float2 v = f[i];
        acc += (double)sqrtf(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)));
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int i = 0; i < 8; i++) s += red[i];
        atomicAdd(&out[pair], s);
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

int run_farneback(vqa_ctx *c, const uint8_t *gray, int npairs, int h, int w, double *mag_sum, float *flow_out)
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
    VQA_BUF(c, R, float, "fb.R", full * 5 * nf);
    VQA_BUF(c, M, float, "fb.M", full * 5 * npairs);
    VQA_BUF(c, flowA, float2, "fb.flowA", full * npairs);
    VQA_BUF(c, flowB, float2, "fb.flowB", full * npairs);
    PolyConst pc;
    make_poly(pc);
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
        dim3 gF(cdiv(lw, 32), cdiv(lh, 8), nf), gP(cdiv(lw, 32), cdiv(lh, 8), npairs);
        VQA_BYTES(c, ((double)full + 4.0 * lw * lh) * nf);
        VQA_LAUNCH(c, k_fb_pyramid, gF, 256, 0, gray, h, w, lh, lw, mode, taps, I);
        VQA_BYTES(c, 24.0 * lw * lh * nf);
        VQA_LAUNCH(c, k_fb_polyexp, dim3(cdiv(lw, PE_TW), cdiv(lh, PE_TH), nf), 256, 0, I, lh, lw, pc, R);
        if (k == levels) {
            VQA_CUDA(c, cudaMemsetAsync(flow, 0, sizeof(float2) * (size_t)lw * lh * npairs, c->stream));
        } else {
            VQA_BYTES(c, (8.0 * lw * lh + 8.0 * pw * ph) * npairs);
            VQA_LAUNCH(c, k_fb_upsample, gP, 256, 0, prev, ph, pw, flow, lh, lw);
        }
        VQA_BYTES(c, 68.0 * lw * lh * npairs);
        VQA_LAUNCH(c, k_fb_matrices, gP, 256, 0, R, flow, lh, lw, M);
        for (int it = 0; it < 3; it++) {
            VQA_BYTES(c, 28.0 * lw * lh * npairs);
            VQA_LAUNCH(c, k_fb_blur_solve, dim3(cdiv(lw, BS_TW), cdiv(lh, BS_TH), npairs), 256, 0, M, lh, lw, flow);
            if (it < 2) {
                VQA_BYTES(c, 68.0 * lw * lh * npairs);
                VQA_LAUNCH(c, k_fb_matrices, gP, 256, 0, R, flow, lh, lw, M);
            }
        }
        float2 *t = prev; prev = flow; flow = t;
        ph = lh; pw = lw;
    }
    // `prev` now holds the level-0 flow
    VQA_CUDA(c, cudaMemsetAsync(mag_sum, 0, sizeof(double) * (size_t)npairs, c->stream));
    int bpf = cdiv((long)full, 256 * 8);
    if (bpf < 1) bpf = 1;
    VQA_BYTES(c, 8.0 * full * npairs);
    VQA_LAUNCH(c, k_fb_mag_sum, dim3(bpf, npairs), 256, 0, prev, (long)full, mag_sum);
    if (flow_out)
        VQA_CUDA(c, cudaMemcpyAsync(flow_out, prev, sizeof(float2) * full * npairs, cudaMemcpyDeviceToDevice, c->stream));
    return VQA_OK;
}

}  // namespace vqa
