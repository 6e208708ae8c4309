/*
 * CPU oracle (plain C) for the per-frame complexity + PSNR/SSIM hot path.
 *
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline through oracle/c_oracle.py (ctypes).  The product library never links it.
 *
 * The reference (complexity_metrics.py / video_processing.py) delegates all arithmetic on
 * this path to opencv-python 4.10.0.84 and to FFmpeg's psnr/ssim filters, neither of which
 * is vendored under /root/reference.  Each function below restates the published algorithm
 * of the library call named in its header comment and cites the reference call site.
 * Pinned against cv2 4.13.0 / the imported reference by oracle/make_golden.py (fixtures in
 * tests/golden/); PSNR/SSIM is pinned only by known-answer cases ("parity unpinned": no
 * ffmpeg binary in the image).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VQO_API __attribute__((visibility("default")))

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int cv_round(double v) { return (int)nearbyint(v); } /* cvRound: half-to-even */

/* ------------------------------------------------------------------------------------------
 * cv2.cvtColor(BGR2GRAY), uint8 (complexity_metrics.py:327-328,358,386,405,493,530)
 * ---------------------------------------------------------------------------------------- */
VQO_API void vqo_bgr2gray(const uint8_t *bgr, int h, int w, uint8_t *gray)
{
    for (long i = 0; i < (long)h * w; i++)
        gray[i] = (uint8_t)((3735 * bgr[3 * i] + 19235 * bgr[3 * i + 1] + 9798 * bgr[3 * i + 2] + (1 << 14)) >> 15);
}

/* ------------------------------------------------------------------------------------------
 * cv2.resize(..., INTER_LINEAR) on uint8 with cn interleaved channels
 * (complexity_metrics.py:359,386,404,430,490,531).  11-bit fixed-point taps, no antialias.
 * ---------------------------------------------------------------------------------------- */
/* Horizontal taps clamp the fraction at the borders; vertical taps keep it and clip the two
 * row indices instead (OpenCV resizeGeneric_): the two differ by rounding when upscaling. */
static void linear_taps(int sn, int dn, int vertical, int *i0, int *i1, int *w0, int *w1)
{
    double scale = (double)sn / (double)dn;
    for (int d = 0; d < dn; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int i = (int)floorf(f);
        float a = f - (float)i;
        if (!vertical) {
            if (i < 0) { i = 0; a = 0.f; }
            if (i >= sn - 1) { i = sn - 1; a = 0.f; }
        }
        i0[d] = imin(imax(i, 0), sn - 1);
        i1[d] = imin(imax(i + 1, 0), sn - 1);
        w1[d] = cv_round(a * 2048.f);
        w0[d] = cv_round((1.f - a) * 2048.f);
    }
}

VQO_API void vqo_resize_u8(const uint8_t *src, int sh, int sw, int cn, uint8_t *dst, int dh, int dw)
{
    if (sh == dh && sw == dw) { memcpy(dst, src, (size_t)sh * sw * cn); return; }
    int *x0 = malloc(sizeof(int) * dw * 4), *x1 = x0 + dw, *a0 = x1 + dw, *a1 = a0 + dw;
    int *y0 = malloc(sizeof(int) * dh * 4), *y1 = y0 + dh, *b0 = y1 + dh, *b1 = b0 + dh;
    linear_taps(sw, dw, 0, x0, x1, a0, a1);
    linear_taps(sh, dh, 1, y0, y1, b0, b1);
    for (int y = 0; y < dh; y++) {
        const uint8_t *r0 = src + (size_t)y0[y] * sw * cn, *r1 = src + (size_t)y1[y] * sw * cn;
        for (int x = 0; x < dw; x++)
            for (int c = 0; c < cn; c++) {
                int t0 = r0[x0[x] * cn + c] * a0[x] + r0[x1[x] * cn + c] * a1[x];
                int t1 = r1[x0[x] * cn + c] * a0[x] + r1[x1[x] * cn + c] * a1[x];
                dst[((size_t)y * dw + x) * cn + c] =
                    (uint8_t)((((b0[y] * (t0 >> 4)) >> 16) + ((b1[y] * (t1 >> 4)) >> 16) + 2) >> 2);
            }
    }
    free(x0); free(y0);
}

/* ------------------------------------------------------------------------------------------
 * cv2.Canny(gray, 100, 200) (aperture 3, L1 gradient) -> number of edge pixels
 * (complexity_metrics.py:503-504).  If edges != NULL also writes the 0/255 map.
 * ---------------------------------------------------------------------------------------- */
VQO_API long vqo_canny_count(const uint8_t *g, int h, int w, int low, int high, uint8_t *edges)
{
    const int TG22 = 13573; /* tan(22.5 deg) * 2^15 */
    size_t n = (size_t)h * w;
    short *dx = malloc(n * sizeof(short)), *dy = malloc(n * sizeof(short));
    int mw = w + 2;
    int *mag = calloc((size_t)(h + 2) * mw, sizeof(int)); /* zero ring around the image */
    uint8_t *st = calloc(n, 1);                           /* 0 none, 1 weak, 2 strong/visited */
    int *stack = malloc(n * sizeof(int));
    long sp = 0, count = 0;
#define PX(yy, xx) ((int)g[(size_t)imin(imax((yy), 0), h - 1) * w + imin(imax((xx), 0), w - 1)])
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int gx = (PX(y - 1, x + 1) + 2 * PX(y, x + 1) + PX(y + 1, x + 1)) -
                     (PX(y - 1, x - 1) + 2 * PX(y, x - 1) + PX(y + 1, x - 1));
            int gy = (PX(y + 1, x - 1) + 2 * PX(y + 1, x) + PX(y + 1, x + 1)) -
                     (PX(y - 1, x - 1) + 2 * PX(y - 1, x) + PX(y - 1, x + 1));
            dx[(size_t)y * w + x] = (short)gx;
            dy[(size_t)y * w + x] = (short)gy;
            mag[(size_t)(y + 1) * mw + x + 1] = abs(gx) + abs(gy);
        }
#undef PX
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int *mp = mag + (size_t)(y + 1) * mw + x + 1;
            int m = mp[0];
            if (m <= low) continue;
            int xs = dx[(size_t)y * w + x], ys = dy[(size_t)y * w + x];
            int ax = abs(xs), ay = abs(ys) << 15;
            int t = ax * TG22, keep;
            if (ay < t) keep = m > mp[-1] && m >= mp[1];
            else if (ay > t + (ax << 16)) keep = m > mp[-mw] && m >= mp[mw];
            else { int s = (xs ^ ys) < 0 ? -1 : 1; keep = m > mp[-mw - s] && m > mp[mw + s]; }
            if (!keep) continue;
            if (m > high) { st[(size_t)y * w + x] = 2; stack[sp++] = y * w + x; }
            else st[(size_t)y * w + x] = 1;
        }
    while (sp > 0) {            /* 8-connected hysteresis */
        int p = stack[--sp], y = p / w, x = p % w;
        count++;
        for (int j = -1; j <= 1; j++)
            for (int i = -1; i <= 1; i++) {
                int yy = y + j, xx = x + i;
                if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                size_t q = (size_t)yy * w + xx;
                if (st[q] == 1) { st[q] = 2; stack[sp++] = (int)q; }
            }
    }
    if (edges) for (size_t i = 0; i < n; i++) edges[i] = st[i] == 2 ? 255 : 0;
    free(dx); free(dy); free(mag); free(st); free(stack);
    return count;
}

/* ------------------------------------------------------------------------------------------
 * FAST-9/16 score (cv2 cornerScore<16>) and ORB keypoint count on a 64x64 level-0 image
 * (complexity_metrics.py:385-387: ORB_create() defaults on gray(resize(frame,(64,64)))).
 * With edgeThreshold 31 the border filter leaves x,y in [31, dim-31): SURVEY.md A.7.
 * ---------------------------------------------------------------------------------------- */
static const int RING_X[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int RING_Y[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

static int fast_strength(const uint8_t *g, int w, int x, int y)
{
    int v = g[(size_t)y * w + x], d[16], best = -512;
    for (int k = 0; k < 16; k++) d[k] = v - g[(size_t)(y + RING_Y[k]) * w + x + RING_X[k]];
    for (int s = 0; s < 16; s++) {
        int mn = 1 << 20, mx = -(1 << 20);
        for (int k = 0; k < 9; k++) { int q = d[(s + k) & 15]; mn = imin(mn, q); mx = imax(mx, q); }
        best = imax(best, imax(mn, -mx));
    }
    return best;
}

/* score map: (strength-1) where strength > thr on [3,dim-3), 0 elsewhere; NMS: strict > 8 nbrs */
VQO_API int vqo_fast_count(const uint8_t *g, int h, int w, int thr, int border, uint8_t *kp_map)
{
    uint8_t *score = calloc((size_t)h * w, 1);
    int count = 0;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = fast_strength(g, w, x, y);
            if (s > thr) score[(size_t)y * w + x] = (uint8_t)(s - 1);
        }
    if (kp_map) memset(kp_map, 0, (size_t)h * w);
    for (int y = imax(border, 3); y < h - imax(border, 3); y++)
        for (int x = imax(border, 3); x < w - imax(border, 3); x++) {
            int s = score[(size_t)y * w + x], ok = s > 0;
            for (int j = -1; j <= 1 && ok; j++)
                for (int i = -1; i <= 1; i++)
                    if ((i || j) && score[(size_t)(y + j) * w + x + i] >= s) { ok = 0; break; }
            if (ok) { count++; if (kp_map) kp_map[(size_t)y * w + x] = (uint8_t)s; }
        }
    free(score);
    return count;
}

VQO_API int vqo_orb_count_64(const uint8_t *gray64)
{
    return vqo_fast_count(gray64, 64, 64, 20, 31, NULL);
}

/* ------------------------------------------------------------------------------------------
 * cv2.calcOpticalFlowFarneback(prev, next, None, 0.5, 3, 15, 3, 5, 1.2, 0)
 * (complexity_metrics.py:340) followed by mean(sqrt(fx^2+fy^2)) (:342-343).
 * Restates OpenCV's optflowgf.cpp: pyramid by Gaussian blur at full resolution + bilinear
 * decimation, 11x11 polynomial expansion, 15x15 box-blurred normal equations, 3 iterations
 * per level.  Double accumulators where OpenCV uses them.
 * ---------------------------------------------------------------------------------------- */
static void gauss_kernel(int n, double sigma, float *k)
{
    if (sigma <= 0 && n == 3) { k[0] = 0.25f; k[1] = 0.5f; k[2] = 0.25f; return; }
    double s = sigma > 0 ? sigma : ((n - 1) * 0.5 - 1) * 0.3 + 0.8, sc = -0.5 / (s * s), sum = 0, t[64];
    for (int i = 0; i < n; i++) { double x = i - (n - 1) * 0.5; t[i] = exp(sc * x * x); sum += t[i]; }
    for (int i = 0; i < n; i++) k[i] = (float)(t[i] / sum);
}

static inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

/* GaussianBlur(float32, ksz, sigma, BORDER_REFLECT_101): rows then columns, float32 */
static void gaussian_blur_f32(const float *src, int h, int w, int ksz, double sigma, float *dst)
{
    float k[64];
    int r = ksz / 2;
    gauss_kernel(ksz, sigma, k);
    float *tmp = malloc(sizeof(float) * (size_t)h * w);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float a = 0.f;
            for (int i = -r; i <= r; i++) a += k[i + r] * src[(size_t)y * w + reflect101(x + i, w)];
            tmp[(size_t)y * w + x] = a;
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float a = 0.f;
            for (int i = -r; i <= r; i++) a += k[i + r] * tmp[(size_t)reflect101(y + i, h) * w + x];
            dst[(size_t)y * w + x] = a;
        }
    free(tmp);
}

static void linear_taps_f32(int sn, int dn, int vertical, int *i0, int *i1, float *a1)
{
    double scale = (double)sn / (double)dn;
    for (int d = 0; d < dn; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int i = (int)floorf(f);
        float a = f - (float)i;
        if (!vertical) {
            if (i < 0) { i = 0; a = 0.f; }
            if (i >= sn - 1) { i = sn - 1; a = 0.f; }
        }
        i0[d] = imin(imax(i, 0), sn - 1); i1[d] = imin(imax(i + 1, 0), sn - 1); a1[d] = a;
    }
}

/* cv2.resize float32 INTER_LINEAR, cn channels (exact 2x decimation = INTER_AREA fast path) */
static void resize_f32(const float *src, int sh, int sw, int cn, float *dst, int dh, int dw)
{
    if (sh == dh && sw == dw) { memcpy(dst, src, sizeof(float) * (size_t)sh * sw * cn); return; }
    if (sw == 2 * dw && sh == 2 * dh) {
        for (int y = 0; y < dh; y++)
            for (int x = 0; x < dw; x++)
                for (int c = 0; c < cn; c++) {
                    const float *p = src + ((size_t)(2 * y) * sw + 2 * x) * cn + c;
                    dst[((size_t)y * dw + x) * cn + c] = (p[0] + p[cn] + p[(size_t)sw * cn] + p[(size_t)sw * cn + cn]) * 0.25f;
                }
        return;
    }
    int *x0 = malloc(sizeof(int) * dw * 2), *x1 = x0 + dw, *y0 = malloc(sizeof(int) * dh * 2), *y1 = y0 + dh;
    float *ax = malloc(sizeof(float) * dw), *ay = malloc(sizeof(float) * dh);
    linear_taps_f32(sw, dw, 0, x0, x1, ax);
    linear_taps_f32(sh, dh, 1, y0, y1, ay);
    for (int y = 0; y < dh; y++) {
        const float *r0 = src + (size_t)y0[y] * sw * cn, *r1 = src + (size_t)y1[y] * sw * cn;
        float b1 = ay[y], b0 = 1.f - b1;
        for (int x = 0; x < dw; x++) {
            float a1 = ax[x], a0 = 1.f - a1;
            for (int c = 0; c < cn; c++) {
                float t0 = r0[x0[x] * cn + c] * a0 + r0[x1[x] * cn + c] * a1;
                float t1 = r1[x0[x] * cn + c] * a0 + r1[x1[x] * cn + c] * a1;
                dst[((size_t)y * dw + x) * cn + c] = t0 * b0 + t1 * b1;
            }
        }
    }
    free(x0); free(y0); free(ax); free(ay);
}

static void invert6(double a[6][6], double inv[6][6])
{
    double m[6][12];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 12; j++) m[i][j] = j < 6 ? a[i][j] : (j - 6 == i ? 1.0 : 0.0);
    for (int c = 0; c < 6; c++) {
        int p = c;
        for (int r = c + 1; r < 6; r++) if (fabs(m[r][c]) > fabs(m[p][c])) p = r;
        if (p != c) for (int j = 0; j < 12; j++) { double t = m[c][j]; m[c][j] = m[p][j]; m[p][j] = t; }
        double d = 1.0 / m[c][c];
        for (int j = 0; j < 12; j++) m[c][j] *= d;
        for (int r = 0; r < 6; r++) if (r != c) {
            double f = m[r][c];
            if (f != 0) for (int j = 0; j < 12; j++) m[r][j] -= f * m[c][j];
        }
    }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) inv[i][j] = m[i][j + 6];
}

/* FarnebackPrepareGaussian: taps g, x g, x^2 g for x in [-n,n] and the 4 used entries of G^-1 */
VQO_API void vqo_polyexp_setup(int n, double sigma, float *g, float *xg, float *xxg, double *ig)
{
    double s = 0;
    if (sigma < 1.19209290e-07) sigma = n * 0.3;
    for (int x = -n; x <= n; x++) { g[x + n] = (float)exp(-x * x / (2 * sigma * sigma)); s += g[x + n]; }
    s = 1. / s;
    for (int x = -n; x <= n; x++) {
        g[x + n] = (float)(g[x + n] * s);
        xg[x + n] = (float)(x * g[x + n]);
        xxg[x + n] = (float)(x * x * g[x + n]);
    }
    double G[6][6] = {{0}}, iG[6][6];
    for (int y = -n; y <= n; y++)
        for (int x = -n; x <= n; x++) {
            float gg = g[y + n] * g[x + n];            /* float products, as OpenCV evaluates them */
            G[0][0] += gg; G[1][1] += gg * x * x; G[3][3] += gg * x * x * x * x; G[5][5] += gg * x * x * y * y;
        }
    G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
    G[4][4] = G[3][3];
    G[3][4] = G[4][3] = G[5][5];
    invert6(G, iG);
    ig[0] = iG[1][1]; ig[1] = iG[0][3]; ig[2] = iG[3][3]; ig[3] = iG[5][5];
}

/* FarnebackPolyExp: I (h x w float) -> R (h x w x 5 float, interleaved) */
static void poly_exp(const float *I, int h, int w, float *R)
{
    enum { N = 5 };
    float gb[2 * N + 1], xgb[2 * N + 1], xxgb[2 * N + 1], *g = gb + N, *xg = xgb + N, *xxg = xxgb + N;
    double ig[4];
    vqo_polyexp_setup(N, 1.2, gb, xgb, xxgb, ig);
    double ig11 = ig[0], ig03 = ig[1], ig33 = ig[2], ig55 = ig[3];
    float *buf = malloc(sizeof(float) * (size_t)(w + 2 * N) * 3), *row = buf + N * 3;
    for (int y = 0; y < h; y++) {
        const float *s0 = I + (size_t)y * w;
        for (int x = 0; x < w; x++) { row[x * 3] = s0[x] * g[0]; row[x * 3 + 1] = row[x * 3 + 2] = 0.f; }
        for (int k = 1; k <= N; k++) {
            const float *a = I + (size_t)imax(y - k, 0) * w, *b = I + (size_t)imin(y + k, h - 1) * w;
            for (int x = 0; x < w; x++) {
                float p = a[x] + b[x];
                row[x * 3] = row[x * 3] + g[k] * p;
                row[x * 3 + 1] = row[x * 3 + 1] + xg[k] * (b[x] - a[x]);
                row[x * 3 + 2] = row[x * 3 + 2] + xxg[k] * p;
            }
        }
        for (int x = 0; x < N * 3; x++) { row[-1 - x] = row[2 - x]; row[w * 3 + x] = row[w * 3 + x - 3]; }
        float *d = R + (size_t)y * w * 5;
        for (int x = 0; x < w; x++) {
            double b1 = row[x * 3] * g[0], b2 = 0, b3 = row[x * 3 + 1] * g[0], b4 = 0, b5 = row[x * 3 + 2] * g[0], b6 = 0;
            for (int k = 1; k <= N; k++) {
                double tg = row[(x + k) * 3] + row[(x - k) * 3];
                b1 += tg * g[k];
                b4 += tg * xxg[k];
                b2 += (row[(x + k) * 3] - row[(x - k) * 3]) * xg[k];
                b3 += (row[(x + k) * 3 + 1] + row[(x - k) * 3 + 1]) * g[k];
                b6 += (row[(x + k) * 3 + 1] - row[(x - k) * 3 + 1]) * xg[k];
                b5 += (row[(x + k) * 3 + 2] + row[(x - k) * 3 + 2]) * g[k];
            }
            d[x * 5 + 1] = (float)(b2 * ig11);
            d[x * 5] = (float)(b3 * ig11);
            d[x * 5 + 3] = (float)(b1 * ig03 + b4 * ig33);
            d[x * 5 + 2] = (float)(b1 * ig03 + b5 * ig33);
            d[x * 5 + 4] = (float)(b6 * ig55);
        }
    }
    free(buf);
}

/* FarnebackUpdateMatrices over all rows */
static void update_matrices(const float *R0, const float *R1, const float *flow, float *M, int h, int w)
{
    static const float border[5] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
    size_t step = (size_t)w * 5;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const float *r0 = R0 + ((size_t)y * w + x) * 5;
            float dx = flow[((size_t)y * w + x) * 2], dy = flow[((size_t)y * w + x) * 2 + 1];
            float fx = x + dx, fy = y + dy;
            int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
            float r2, r3, r4, r5, r6;
            fx -= x1; fy -= y1;
            if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
                const float *p = R1 + (size_t)y1 * step + (size_t)x1 * 5;
                float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
                r2 = a00 * p[0] + a01 * p[5] + a10 * p[step] + a11 * p[step + 5];
                r3 = a00 * p[1] + a01 * p[6] + a10 * p[step + 1] + a11 * p[step + 6];
                r4 = a00 * p[2] + a01 * p[7] + a10 * p[step + 2] + a11 * p[step + 7];
                r5 = a00 * p[3] + a01 * p[8] + a10 * p[step + 3] + a11 * p[step + 8];
                r6 = a00 * p[4] + a01 * p[9] + a10 * p[step + 4] + a11 * p[step + 9];
                r4 = (r0[2] + r4) * 0.5f;
                r5 = (r0[3] + r5) * 0.5f;
                r6 = (r0[4] + r6) * 0.25f;
            } else {
                r2 = r3 = 0.f;
                r4 = r0[2]; r5 = r0[3]; r6 = r0[4] * 0.5f;
            }
            r2 = (r0[0] - r2) * 0.5f;
            r3 = (r0[1] - r3) * 0.5f;
            r2 += r4 * dy + r6 * dx;
            r3 += r6 * dy + r5 * dx;
            if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
                float sc = (x < 5 ? border[x] : 1.f) * (x >= w - 5 ? border[w - x - 1] : 1.f) *
                           (y < 5 ? border[y] : 1.f) * (y >= h - 5 ? border[h - y - 1] : 1.f);
                r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
            }
            float *m = M + ((size_t)y * w + x) * 5;
            m[0] = r4 * r4 + r6 * r6;
            m[1] = (r4 + r5) * r6;
            m[2] = r5 * r5 + r6 * r6;
            m[3] = r4 * r2 + r6 * r3;
            m[4] = r6 * r2 + r5 * r3;
        }
}

/* FarnebackUpdateFlow_Blur: 15x15 replicate-border box mean of M (double sums), 2x2 solve */
static void update_flow_blur(const float *M, float *flow, int h, int w, int bs)
{
    int m = bs / 2;
    double scale = 1. / (bs * bs);
    double *vbuf = malloc(sizeof(double) * (size_t)(w + 2 * m + 2) * 5), *vsum = vbuf + (m + 1) * 5;
    for (int x = 0; x < w * 5; x++) vsum[x] = M[x] * (double)(m + 2);
    for (int y = 1; y < m; y++) {
        const float *s = M + (size_t)imin(y, h - 1) * w * 5;
        for (int x = 0; x < w * 5; x++) vsum[x] += s[x];
    }
    for (int y = 0; y < h; y++) {
        const float *s0 = M + (size_t)imax(y - m - 1, 0) * w * 5, *s1 = M + (size_t)imin(y + m, h - 1) * w * 5;
        for (int x = 0; x < w * 5; x++) vsum[x] += s1[x] - s0[x];
        for (int x = 0; x < (m + 1) * 5; x++) { vsum[-1 - x] = vsum[4 - x]; vsum[w * 5 + x] = vsum[w * 5 + x - 5]; }
        double g11 = vsum[0] * (m + 2), g12 = vsum[1] * (m + 2), g22 = vsum[2] * (m + 2),
               h1 = vsum[3] * (m + 2), h2 = vsum[4] * (m + 2);
        for (int x = 1; x < m; x++) {
            g11 += vsum[x * 5]; g12 += vsum[x * 5 + 1]; g22 += vsum[x * 5 + 2]; h1 += vsum[x * 5 + 3]; h2 += vsum[x * 5 + 4];
        }
        float *f = flow + (size_t)y * w * 2;
        for (int x = 0; x < w; x++) {
            g11 += vsum[(x + m) * 5] - vsum[(x - m) * 5 - 5];
            g12 += vsum[(x + m) * 5 + 1] - vsum[(x - m) * 5 - 4];
            g22 += vsum[(x + m) * 5 + 2] - vsum[(x - m) * 5 - 3];
            h1 += vsum[(x + m) * 5 + 3] - vsum[(x - m) * 5 - 2];
            h2 += vsum[(x + m) * 5 + 4] - vsum[(x - m) * 5 - 1];
            double a = g11 * scale, b = g12 * scale, c = g22 * scale, p = h1 * scale, q = h2 * scale;
            double idet = 1. / (a * c - b * b + 1e-3);
            f[x * 2] = (float)((a * q - b * p) * idet);
            f[x * 2 + 1] = (float)((c * p - b * q) * idet);
        }
    }
    free(vbuf);
}

static float pairwise_sum_f32(const float *a, size_t n)
{   /* numpy's pairwise float32 summation (block 128, 8 accumulators) as used by np.mean */
    if (n < 8) { float s = 0.f; for (size_t i = 0; i < n; i++) s += a[i]; return s; }
    if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; j++) r[j] = a[j];
        size_t i;
        for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; j++) r[j] += a[i + j];
        float s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) s += a[i];
        return s;
    }
    size_t n2 = n / 2; n2 -= n2 % 8;
    return pairwise_sum_f32(a, n2) + pairwise_sum_f32(a + n2, n - n2);
}

/* Full chain.  flow_out (h*w*2 floats, optional) receives the level-0 flow. Returns mean |flow|. */
VQO_API float vqo_farneback_mean_mag(const uint8_t *prev, const uint8_t *next, int h, int w, float *flow_out)
{
    const double pyr_scale = 0.5;
    const int levels_req = 3, winsize = 15, iters = 3, min_size = 32;
    int levels, k;
    double scale = 1;
    for (k = 0; k < levels_req; k++) {
        scale *= pyr_scale;
        if (w * scale < min_size || h * scale < min_size) break;
    }
    levels = k;
    size_t n = (size_t)h * w;
    const uint8_t *img[2] = {prev, next};
    float *fimg = malloc(sizeof(float) * n), *blur = malloc(sizeof(float) * n), *I = malloc(sizeof(float) * n);
    float *R[2] = {malloc(sizeof(float) * n * 5), malloc(sizeof(float) * n * 5)};
    float *M = malloc(sizeof(float) * n * 5);
    float *flow = NULL, *prev_flow = NULL;
    int pw = 0, ph = 0;
    for (k = levels; k >= 0; k--) {
        scale = 1;
        for (int i = 0; i < k; i++) scale *= pyr_scale;
        double sigma = (1. / scale - 1) * 0.5;
        int ksz = imax(cv_round(sigma * 5) | 1, 3);
        int lw = cv_round(w * scale), lh = cv_round(h * scale);
        flow = calloc((size_t)lw * lh * 2, sizeof(float));
        if (prev_flow) {
            resize_f32(prev_flow, ph, pw, 2, flow, lh, lw);
            for (size_t i = 0; i < (size_t)lw * lh * 2; i++) flow[i] *= (float)(1. / pyr_scale);
            free(prev_flow);
        }
        for (int i = 0; i < 2; i++) {
            for (size_t j = 0; j < n; j++) fimg[j] = (float)img[i][j];
            gaussian_blur_f32(fimg, h, w, ksz, sigma, blur);
            resize_f32(blur, h, w, 1, I, lh, lw);
            poly_exp(I, lh, lw, R[i]);
        }
        update_matrices(R[0], R[1], flow, M, lh, lw);
        for (int i = 0; i < iters; i++) {
            update_flow_blur(M, flow, lh, lw, winsize);
            if (i < iters - 1) update_matrices(R[0], R[1], flow, M, lh, lw);
        }
        prev_flow = flow; pw = lw; ph = lh;
    }
    float *mag = malloc(sizeof(float) * n);
    for (size_t i = 0; i < n; i++) mag[i] = sqrtf(flow[2 * i] * flow[2 * i] + flow[2 * i + 1] * flow[2 * i + 1]);
    float mean = pairwise_sum_f32(mag, n) / (float)n;
    if (flow_out) memcpy(flow_out, flow, sizeof(float) * n * 2);
    free(mag); free(flow); free(fimg); free(blur); free(I); free(R[0]); free(R[1]); free(M);
    return mean;
}

/* ------------------------------------------------------------------------------------------
 * FFmpeg psnr / ssim filters on 8-bit planes (video_processing.py:274-291)
 * ---------------------------------------------------------------------------------------- */
VQO_API uint64_t vqo_plane_sse(const uint8_t *a, const uint8_t *b, int h, int w, int stride_a, int stride_b)
{
    uint64_t s = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int d = (int)a[(size_t)y * stride_a + x] - (int)b[(size_t)y * stride_b + x];
            s += (uint64_t)(d * d);
        }
    return s;
}

VQO_API double vqo_ssim_plane(const uint8_t *a, const uint8_t *b, int h, int w, int stride_a, int stride_b)
{
    int bw = w >> 2, bh = h >> 2;
    if (bw < 2 || bh < 2) return 0.0;
    int (*sums)[4] = malloc(sizeof(int[4]) * (size_t)bw * bh);
    for (int by = 0; by < bh; by++)
        for (int bx = 0; bx < bw; bx++) {
            int s1 = 0, s2 = 0, ss = 0, s12 = 0;
            for (int y = 0; y < 4; y++)
                for (int x = 0; x < 4; x++) {
                    int p = a[(size_t)(by * 4 + y) * stride_a + bx * 4 + x], q = b[(size_t)(by * 4 + y) * stride_b + bx * 4 + x];
                    s1 += p; s2 += q; ss += p * p + q * q; s12 += p * q;
                }
            int *o = sums[(size_t)by * bw + bx];
            o[0] = s1; o[1] = s2; o[2] = ss; o[3] = s12;
        }
    double total = 0;
    for (int y = 0; y < bh - 1; y++) {
        float row = 0.f;
        for (int x = 0; x < bw - 1; x++) {
            int q[4];
            for (int c = 0; c < 4; c++)
                q[c] = sums[(size_t)y * bw + x][c] + sums[(size_t)y * bw + x + 1][c] +
                       sums[(size_t)(y + 1) * bw + x][c] + sums[(size_t)(y + 1) * bw + x + 1][c];
            int s1 = q[0], s2 = q[1], ss = q[2], s12 = q[3];
            int vars = ss * 64 - s1 * s1 - s2 * s2, covar = s12 * 64 - s1 * s2;
            row += (float)(2 * s1 * s2 + 416) * (float)(2 * covar + 235963) /
                   ((float)(s1 * s1 + s2 * s2 + 416) * (float)(vars + 235963));
        }
        total += row;
    }
    free(sums);
    return total / ((bw - 1) * (bh - 1));
}
