// Fused Farneback flow iteration (round 2); included by farneback.cu inside namespace vqa.
//
//   flow_out = solve2x2(boxmean15x15(UpdateMatrices(R_prev, R_cur, flow_in)))
//
// in ONE kernel: UpdateMatrices is evaluated in the load stage of the column-marching box blur, so the 5-plane M
// field (20 B/px written + 20 B/px read back per iteration, 40 of the 96 B/px of the UpdateMatrices -> blur chain)
// never exists in memory.  Algorithmic bytes per pixel and iteration: R0 20 + R1 20 + flow in 8 + flow out 8 = 56
// (first iteration of a level: 48 + the coarser flow; last iteration of level 0: 48, the flow itself is not stored).
//
// Geometry: a block of FJ_W = 256 threads owns a strip of 240 output columns (+ 8 halo columns per side, one column per
// thread; 1920 = 8 x 240 and 3840 = 16 x 240, so the 1080p / 4K pyramids have no ragged strip) and walks down
// `rows_per_block` rows.  Per incoming row every thread evaluates M of its own pixel (fb_matrix_core's arithmetic: 5 R0
// values, the flow, 20 bilinear gathers of R1) and feeds the vertical 15-row window sums.
//
// Window sums without an outgoing row (van Herk / Gil-Werman): rows are cut into segments of 15 anchored at ABSOLUTE
// rows; a thread-private shared-memory ring of 15 slots x 5 planes holds, for the current segment, the SUFFIX sums
// sum(rows o..14) and is overwritten slot by slot with the raw rows of the next segment while a register keeps their
// running PREFIX; window(o) = suffix[o] + prefix(o-1).  Every window is a plain float sum of 15 terms: nothing persists
// from row to row (the double-precision running sums of k_fb_blur_solve cost 15 float<->double conversions per row),
// no FP64, no XU, and the result does not depend on the strip partition.
//
// Latency (version 1 of this kernel, profiles/r02_notes.md 1, sat 30 % of its stall samples on the first use of the
// R0 row loaded ONE row period earlier): (a) every row stream of the strip -- 5 planes of R0, 5 of R1, the flow -- is
// prefetched into L2 FJ_PF rows ahead by one lane per 128-byte line, so the demand loads are L2 hits; (b) the matrix
// evaluation stays software-pipelined (R0 + flow of row r+2 loaded, gathers of row r+1 issued, row r finished);
// (c) ONE barrier per row instead of two: the row of vertical sums and the row of horizontal sums are double
// buffered, iteration i stores the vertical sums of row i, runs the horizontal pass of row i-1 and the 2x2 solve of
// row i-2.
constexpr int FJ_W = 256, FJ_OUT = 240, FJ_SEG = 15, FJ_RP = 288, FJ_PF = 6;
constexpr int FJ_SMEM = (FJ_SEG * 5 * FJ_W + 4 * 5 * FJ_RP) * (int)sizeof(float);

struct FiA {                      // loads of one row in flight: R0 and the flow (or the coarser level's four taps)
    float q[5];
    float2 f;                     // INIT 0
    float2 p00, p01, p10, p11;    // INIT 1
    float ay;
};
struct FiB {                      // gathers of one row in flight
    float q[5];
    float dx, dy, fx, fy;
    float t[20];
    int inb;
};

__device__ __forceinline__ void fj_prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <int INIT>
__device__ __forceinline__ void fi_load(FiA &A, const float *__restrict__ R0, const float2 *__restrict__ flow,
                                        const float2 *__restrict__ prev, int ph, int pw, int px0, int px1, size_t plane,
                                        int gx, int r, int w, int lh)
{
    const size_t o = (size_t)r * w + gx;
#pragma unroll
    for (int c = 0; c < 5; c++) A.q[c] = __ldg(R0 + c * plane + o);
    if (INIT == 0) A.f = __ldg(flow + o);
    if (INIT == 1) {
        int y0, y1;
        if (lh == 2 * ph) up2_tap_f32(r, ph, true, y0, y1, A.ay);
        else lin_tap_f32(r, ph, lh, true, y0, y1, A.ay);
        A.p00 = __ldg(prev + (size_t)y0 * pw + px0);
        A.p01 = __ldg(prev + (size_t)y0 * pw + px1);
        A.p10 = __ldg(prev + (size_t)y1 * pw + px0);
        A.p11 = __ldg(prev + (size_t)y1 * pw + px1);
    }
}

template <int INIT>
__device__ __forceinline__ void fi_issue(FiB &B, const FiA &A, const float *__restrict__ R1, size_t plane, int gx, int r,
                                         int h, int w, float ax)
{
    float2 f;
    if (INIT == 0) f = A.f;
    else if (INIT == 1) {
        const float a0 = 1.f - ax, b0 = 1.f - A.ay;
        f.x = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(A.p00.x, a0), __fmul_rn(A.p01.x, ax)), b0),
                        __fmul_rn(__fadd_rn(__fmul_rn(A.p10.x, a0), __fmul_rn(A.p11.x, ax)), A.ay)) * 2.f;
        f.y = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(A.p00.y, a0), __fmul_rn(A.p01.y, ax)), b0),
                        __fmul_rn(__fadd_rn(__fmul_rn(A.p10.y, a0), __fmul_rn(A.p11.y, ax)), A.ay)) * 2.f;
    } else f = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 5; c++) B.q[c] = A.q[c];
    B.dx = f.x;
    B.dy = f.y;
    float fx = (float)gx + f.x, fy = (float)r + f.y;
    const int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    B.fx = fx - (float)x1;
    B.fy = fy - (float)y1;
    B.inb = ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) ? 1 : 0;
    if (B.inb) {
        const float *b = R1 + (size_t)y1 * w + x1;
#pragma unroll
        for (int c = 0; c < 5; c++) {
            B.t[4 * c] = __ldg(b + c * plane);
            B.t[4 * c + 1] = __ldg(b + c * plane + 1);
            B.t[4 * c + 2] = __ldg(b + c * plane + w);
            B.t[4 * c + 3] = __ldg(b + c * plane + w + 1);
        }
    }
}

// fb_matrix_core's arithmetic on the operands gathered by fi_issue
__device__ __forceinline__ void fi_finish(const FiB &B, int x, int y, int h, int w, float m[5])
{
    float r2, r3, r4, r5, r6;
    if (B.inb) {
        const float fx = B.fx, fy = B.fy;
        const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
#define FI_TAP(c) (a00 * B.t[4 * (c)] + a01 * B.t[4 * (c) + 1] + a10 * B.t[4 * (c) + 2] + a11 * B.t[4 * (c) + 3])
        r2 = FI_TAP(0);
        r3 = FI_TAP(1);
        r4 = FI_TAP(2);
        r5 = FI_TAP(3);
        r6 = FI_TAP(4);
#undef FI_TAP
        r4 = (B.q[2] + r4) * 0.5f;
        r5 = (B.q[3] + r5) * 0.5f;
        r6 = (B.q[4] + r6) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = B.q[2];
        r5 = B.q[3];
        r6 = B.q[4] * 0.5f;
    }
    r2 = (B.q[0] - r2) * 0.5f;
    r3 = (B.q[1] - r3) * 0.5f;
    r2 += r4 * B.dy + r6 * B.dx;
    r3 += r6 * B.dy + r5 * B.dx;
    if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
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

template <int INIT>
__global__ void __launch_bounds__(FJ_W, 2)
k_fb_iter(const float *__restrict__ R, const float2 *__restrict__ flow_in, int h, int w, float2 *__restrict__ flow_out,
          int rows_per_block, double *__restrict__ mag_sum, int write_flow, const float2 *__restrict__ prev, int ph, int pw)
{
    extern __shared__ __align__(16) float fj_smem[];
    float *ring = fj_smem;                                             // [FJ_SEG][5][FJ_W], column t is private to thread t
    float *rowb = fj_smem + FJ_SEG * 5 * FJ_W;                         // [2][5][FJ_RP] vertical sums of a row
    float *hsb = rowb + 2 * 5 * FJ_RP;                                 // [2][5][FJ_RP] horizontal sums of a row
    const int pair = blockIdx.z, t = threadIdx.x, lane = t & 31;
    const size_t plane = (size_t)h * w;
    const float *R0 = R + (size_t)pair * 5 * plane, *R1 = R0 + 5 * plane;
    const float2 *fin = INIT == 0 ? flow_in + (size_t)pair * plane : nullptr;
    const float2 *pv = INIT == 1 ? prev + (size_t)pair * ph * pw : nullptr;
    const int sx0 = blockIdx.x * FJ_OUT, y0 = blockIdx.y * rows_per_block;
    const int gx = clampi(sx0 - 8 + t, 0, w - 1);
    const int y_end = min(y0 + rows_per_block, h), nrows = y_end - y0;
    const int rows_need = nrows + 2 * MS_R;                            // M rows y0-7 .. y_end+6 (clamped to the image)
    // x taps of the coarser level's flow are per column: once per thread
    int px0 = 0, px1 = 0;
    float pax = 0.f;
    if (INIT == 1) {
        if (w == 2 * pw) up2_tap_f32(gx, pw, false, px0, px1, pax);
        else lin_tap_f32(gx, pw, w, false, px0, px1, pax);
    }
#define FJ_ROW(rr) clampi(y0 - MS_R + (rr), 0, h - 1)
    // L2 prefetch of the row streams: lanes 0 and 31 of a warp touch the (at most two) 128-byte lines the warp's 32
    // columns span in each of the 10 R planes; the float2 flow row spans up to three
    const bool pf_lane = lane == 0 || lane == 31;
    auto prefetch_row = [&](int rr) {
        if (rr >= rows_need) return;
        const size_t o = (size_t)FJ_ROW(rr) * w + gx;
        if (pf_lane) {
#pragma unroll
            for (int c = 0; c < 10; c++) fj_prefetch_l2(R0 + c * plane + o);
        }
        if (INIT == 0 && (pf_lane || lane == 16)) fj_prefetch_l2(fin + o);
    };
    for (int rr = 0; rr < FJ_PF + 2; rr++) prefetch_row(rr);
    FiA A;
    FiB B;
    fi_load<INIT>(A, R0, fin, pv, ph, pw, px0, px1, plane, gx, FJ_ROW(0), w, h);
    fi_issue<INIT>(B, A, R1, plane, gx, FJ_ROW(0), h, w, pax);
    fi_load<INIT>(A, R0, fin, pv, ph, pw, px0, px1, plane, gx, FJ_ROW(1), w, h);
    // produce M of row rr (gathers issued one call earlier), then advance both pipeline stages
    auto produce = [&](int rr, float m[5]) {
        fi_finish(B, gx, FJ_ROW(rr), h, w, m);
        if (rr + 1 < rows_need) fi_issue<INIT>(B, A, R1, plane, gx, FJ_ROW(rr + 1), h, w, pax);
        if (rr + 2 < rows_need) fi_load<INIT>(A, R0, fin, pv, ph, pw, px0, px1, plane, gx, FJ_ROW(rr + 2), w, h);
        prefetch_row(rr + 2 + FJ_PF);
    };
    float *my_ring = ring + t;
    // Segments are anchored at ABSOLUTE rows (window of output row y = rows a in [y, y+14] of a = image row + 7;
    // segment k = a in [15k, 15k+14]), not at the strip start: the two float sums that make a window are then the
    // same for every strip partition, so the flow does not depend on how many pairs a launch carries.
    const int o0 = y0 % FJ_SEG;
    float pre[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    // warm-up: rows a = y0 .. y0+14.  The first 15-o0 complete the current segment (slots o0..14, then suffix sums
    // in place), the next o0 are the head of the following segment (slots 0..o0-1, running prefix)
    for (int rr = 0; rr < FJ_SEG - o0; rr++) {
        float m[5];
        produce(rr, m);
#pragma unroll
        for (int c = 0; c < 5; c++) my_ring[((o0 + rr) * 5 + c) * FJ_W] = m[c];
    }
#pragma unroll
    for (int c = 0; c < 5; c++) {
        float acc = my_ring[((FJ_SEG - 1) * 5 + c) * FJ_W];
        for (int i = FJ_SEG - 2; i >= o0; i--) {
            acc += my_ring[(i * 5 + c) * FJ_W];
            my_ring[(i * 5 + c) * FJ_W] = acc;
        }
    }
    for (int rr = FJ_SEG - o0; rr < FJ_SEG; rr++) {
        float m[5];
        produce(rr, m);
#pragma unroll
        for (int c = 0; c < 5; c++) {
            my_ring[((rr - (FJ_SEG - o0)) * 5 + c) * FJ_W] = m[c];
            pre[c] += m[c];
        }
    }
    // horizontal work item: one warp per plane, 30 segments of 8 outputs (+ 2 idle lanes); warps 5..7 skip the phase
    const int hc = t >> 5, hseg = lane;
    const bool hwork = hc < 5 && hseg < FJ_OUT / MS_SEG;
    const int ox = t - 8, gxo = sx0 + ox;
    const bool has_out = ox >= 0 && ox < FJ_OUT && gxo < w;
    const int my_sw = ms_sw(t);
    int hoff[6];
#pragma unroll
    for (int j = 0; j < 6; j++) hoff[j] = (hwork ? hc : 0) * FJ_RP + ms_sw((hwork ? hseg : 0) * MS_SEG + 4 * j);
    const int hout0 = (hwork ? hc : 0) * FJ_RP + ms_sw((hwork ? hseg : 0) * MS_SEG + 8);
    const int hout1 = (hwork ? hc : 0) * FJ_RP + ms_sw((hwork ? hseg : 0) * MS_SEG + 12);
    float2 *fout = flow_out + (size_t)pair * plane + (size_t)y0 * w + (has_out ? gxo : 0);
    int o = o0;
    double mag_acc = 0;
    // iteration i: vertical sums of row i -> rowb[i & 1]; horizontal pass of row i-1: rowb[(i-1) & 1] -> hsb[(i-1) & 1];
    // solve of row i-2 from hsb[i & 1]; the raw row entering the window of row i+1 replaces the consumed ring slot
    for (int i = 0; i < nrows + 2; i++) {
        const int buf = (i & 1) * 5 * FJ_RP;
        if (i < nrows) {
            float *slot = my_ring + o * 5 * FJ_W;
#pragma unroll
            for (int c = 0; c < 5; c++) rowb[buf + c * FJ_RP + my_sw] = slot[c * FJ_W] + pre[c];
        }
        if (hwork && i >= 1 && i <= nrows) {
            const float *hrow = rowb + (5 * FJ_RP - buf);
            float *hdst = hsb + (5 * FJ_RP - buf);
            float p[24];
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const float4 v = *reinterpret_cast<const float4 *>(hrow + hoff[j]);
                p[4 * j] = v.x; p[4 * j + 1] = v.y; p[4 * j + 2] = v.z; p[4 * j + 3] = v.w;
            }
            const float core = ((p[8] + p[9]) + (p[10] + p[11])) + ((p[12] + p[13]) + (p[14] + p[15]));
            float L[8], Rr[8];
            L[7] = 0.f;
#pragma unroll
            for (int j = 6; j >= 0; j--) L[j] = L[j + 1] + p[j + 1];
            Rr[0] = 0.f;
#pragma unroll
            for (int j = 1; j < 8; j++) Rr[j] = Rr[j - 1] + p[15 + j];
            float o8[MS_SEG];
#pragma unroll
            for (int j = 0; j < MS_SEG; j++) o8[j] = (core + L[j]) + Rr[j];
            *reinterpret_cast<float4 *>(hdst + hout0) = make_float4(o8[0], o8[1], o8[2], o8[3]);
            *reinterpret_cast<float4 *>(hdst + hout1) = make_float4(o8[4], o8[5], o8[6], o8[7]);
        }
        if (i >= 2) {
            if (has_out) {
                // 2x2 solve on the raw window sums (see k_fb_blur_solve): float with FMA-recovered product errors
                const float *my_hs = hsb + buf + my_sw;
                const float s11 = my_hs[0], s12 = my_hs[FJ_RP], s22 = my_hs[2 * FJ_RP], t1 = my_hs[3 * FJ_RP], t2 = my_hs[4 * FJ_RP];
                const float k2 = 1.f / (225.f * 225.f);
                const float w0 = __fmul_rn(s12, s12), w1 = __fmul_rn(s12, t1), w2 = __fmul_rn(s12, t2);
                const float det = __fmaf_rn(__fadd_rn(__fmaf_rn(s11, s22, -w0), __fmaf_rn(-s12, s12, w0)), k2, 1e-3f);
                const float nx = __fadd_rn(__fmaf_rn(s11, t2, -w1), __fmaf_rn(-s12, t1, w1));
                const float ny = __fadd_rn(__fmaf_rn(s22, t1, -w2), __fmaf_rn(-s12, t2, w2));
                const float idet = __frcp_rn(det);
                float2 ov;
                ov.x = __fmul_rn(__fmul_rn(nx, k2), idet);
                ov.y = __fmul_rn(__fmul_rn(ny, k2), idet);
                if (write_flow) *fout = ov;
                if (mag_sum) mag_acc += (double)sqrtf(__fadd_rn(__fmul_rn(ov.x, ov.x), __fmul_rn(ov.y, ov.y)));
            }
            fout += w;
        }
        if (i < nrows) {
            // the raw row that enters the NEXT window replaces the suffix slot this window just consumed
            float *slot = my_ring + o * 5 * FJ_W;
            const int rr = i + FJ_SEG;
            if (rr < rows_need) {
                float m[5];
                produce(rr, m);
#pragma unroll
                for (int c = 0; c < 5; c++) {
                    slot[c * FJ_W] = m[c];
                    pre[c] += m[c];
                }
            }
            if (++o == FJ_SEG) {                                       // the ring now holds a whole raw segment
                o = 0;
#pragma unroll
                for (int c = 0; c < 5; c++) {
                    pre[c] = 0.f;
                    float acc = my_ring[((FJ_SEG - 1) * 5 + c) * FJ_W];
#pragma unroll
                    for (int k = FJ_SEG - 2; k >= 0; k--) {
                        acc += my_ring[(k * 5 + c) * FJ_W];
                        my_ring[(k * 5 + c) * FJ_W] = acc;
                    }
                }
            }
        }
        __syncthreads();
    }
#undef FJ_ROW
    if (mag_sum) {
        __shared__ double red[FJ_W / 32];
        mag_acc = warp_sum(mag_acc);
        if (lane == 0) red[t >> 5] = mag_acc;
        __syncthreads();
        if (t == 0) {
            double sum = 0;
            for (int k = 0; k < FJ_W / 32; k++) sum += red[k];
            atomicAdd(&mag_sum[pair], sum);
        }
    }
}

// rows per block: tall strips amortise the 14 extra rows of UpdateMatrices (and the 2 drain iterations) a strip pays
// above and below its outputs, but the grid should fill whole waves of 2 blocks per SM: pick the strip count with the
// smallest (waves x rows walked per block)
static int fj_rows_per_block(int lh, int lw, int npairs, int sm_count, int h_cap, int h_min)
{
    const long slots = 2L * sm_count, per_row_strip = (long)cdiv(lw, FJ_OUT) * npairs;
    int best_rows = lh;
    long best_cost = -1;
    for (int sy = 1; sy <= lh; sy++) {
        const int rows = cdiv(lh, sy);
        if (rows > h_cap && sy < lh) continue;
        if (rows < h_min && sy > 1) break;
        const long blocks = per_row_strip * cdiv(lh, rows);
        const long cost = ((blocks + slots - 1) / slots) * (rows + 2 * MS_R + 2);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_rows = rows; }
    }
    return best_rows;
}

constexpr int FJ_H_CAP = 360, FJ_H_MIN = 30;

// the three iterations of one pyramid level; Q (= `flow`) receives this level's flow, P (= `prev`) holds the coarser one
static int run_fused_level(vqa_ctx *c, const float *R, float2 *Q, float2 *P, int ph, int pw, int lh, int lw, int npairs,
                           bool coarsest, bool finest, double *mag_sum, bool keep_flow)
{
#ifdef VQA_AB
    const int h_cap = (getenv("VQA_FI_H") && atoi(getenv("VQA_FI_H")) >= 16) ? atoi(getenv("VQA_FI_H")) : FJ_H_CAP;
#else
    constexpr int h_cap = FJ_H_CAP;
#endif
    VQA_CUDA(c, cudaFuncSetAttribute(k_fb_iter<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FJ_SMEM));
    VQA_CUDA(c, cudaFuncSetAttribute(k_fb_iter<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FJ_SMEM));
    VQA_CUDA(c, cudaFuncSetAttribute(k_fb_iter<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FJ_SMEM));
    const int rows_pb = fj_rows_per_block(lh, lw, npairs, c->sm_count, h_cap, FJ_H_MIN);
    const dim3 gI(cdiv(lw, FJ_OUT), cdiv(lh, rows_pb), npairs);
    const double px = (double)lw * lh * npairs;
    // iteration 0 writes Q from the coarser flow in P; iteration 1 reads Q and writes P (the coarser flow is no longer
    // needed); iteration 2 reads P and writes Q
    for (int it = 0; it < 3; it++) {
        const bool last = finest && it == 2;
        double *ms = last ? mag_sum : (double *)nullptr;
        const int wf = (!last || keep_flow) ? 1 : 0;
        if (it == 0 && coarsest) {
            VQA_BYTES(c, 48.0 * px);
            VQA_LAUNCH(c, k_fb_iter<2>, gI, FJ_W, FJ_SMEM, R, (const float2 *)nullptr, lh, lw, Q, rows_pb, ms, wf, (const float2 *)nullptr, 0, 0);
        } else if (it == 0) {
            VQA_BYTES(c, 48.0 * px + 8.0 * pw * ph * npairs);
            VQA_LAUNCH(c, k_fb_iter<1>, gI, FJ_W, FJ_SMEM, R, (const float2 *)nullptr, lh, lw, Q, rows_pb, ms, wf, P, ph, pw);
        } else {
            const float2 *in = (it == 1) ? Q : P;
            float2 *out = (it == 1) ? P : Q;
            VQA_BYTES(c, (last && !keep_flow ? 48.0 : 56.0) * px);
            VQA_LAUNCH(c, k_fb_iter<0>, gI, FJ_W, FJ_SMEM, R, in, lh, lw, out, rows_pb, ms, wf, (const float2 *)nullptr, 0, 0);
        }
    }
    return VQA_OK;
}
