// Development-only (-DVQA_AB) variant of the Farneback flow iteration; included by farneback.cu inside namespace vqa.
// Measured in round 2 (profiles/r02/): 10 % slower per step than the UpdateMatrices -> blur chain it fuses,
// kept for the A/B record, never part of the product build.
// ================================================================================================
// Fused flow iteration (round 2): UpdateMatrices evaluated in the LOAD stage of the box blur.
//
//   flow_out = solve2x2(boxmean15x15(UpdateMatrices(R_prev, R_cur, flow_in)))
//
// in one kernel, so the 5-plane M field (20 B/px written + 20 B/px read back per iteration, 60 % of the
// bytes of the round-1 chain) never exists in memory.  Geometry is the marching blur's: a block owns a strip
// of 112 output columns (+ 8 halo columns per side = 128 threads, one column each) and walks down
// `rows_per_block` rows.  Per incoming row every thread evaluates M of its own pixel (fb_matrix_core's
// arithmetic: 5 R0 values, the flow, 20 bilinear gathers of R1) and feeds the vertical 15-row window sums.
//
// Window sums without an outgoing row (van Herk / Gil-Werman): rows are cut into segments of 15; a
// thread-private shared-memory ring of 15 slots x 5 planes holds, for the current segment, the SUFFIX sums
// sum(rows o..14) and is overwritten slot by slot with the raw rows of the next segment while a register
// keeps their running PREFIX; window(o) = suffix[o] + prefix(o-1).  Every window is a plain float sum of
// 15 terms: nothing persists from row to row (the property the double-precision running sums of round 1
// bought with 15 float<->double conversions per row, 31 % of that kernel's stall samples), no FP64, no XU.
//
// The UpdateMatrices of a row is software-pipelined two rows deep: R0 + flow of row r+2 are loaded, the 20
// gathers of row r+1 are issued (their addresses need the flow) and the matrix of row r is finished from
// gathers issued one iteration earlier, so both dependent latencies hide behind a full row of blur work.
//
// Algorithmic bytes per pixel and iteration: R0 20 + R1 20 + flow in 8 + flow out 8 = 56 (first iteration of
// a level: 48 + the coarser flow; last iteration of level 0: 48, the flow itself is not stored).
constexpr int FI_W = 128, FI_OUT = 112, FI_SEG = 15;
constexpr int FI_SMEM = (FI_SEG * 5 * FI_W + 2 * 5 * MS_VP) * (int)sizeof(float);

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

template <int INIT>
__device__ __forceinline__ void fi_load(FiA &A, const float *__restrict__ R0, const float2 *__restrict__ flow,
                                        const float2 *__restrict__ prev, int ph, int pw, int px0, int px1, size_t plane,
                                        int gx, int r, int h, int w, int lh)
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
__global__ void __launch_bounds__(FI_W)
k_fb_iter(const float *__restrict__ R, const float2 *__restrict__ flow_in, int h, int w, float2 *__restrict__ flow_out,
          int rows_per_block, double *__restrict__ mag_sum, int write_flow, const float2 *__restrict__ prev, int ph, int pw)
{
    extern __shared__ __align__(16) float fi_smem[];
    float *ring = fi_smem;                                             // [FI_SEG][5][FI_W], column t is private to thread t
    float (*row)[MS_VP] = reinterpret_cast<float (*)[MS_VP]>(fi_smem + FI_SEG * 5 * FI_W);
    float (*hs)[MS_VP] = row + 5;
    const int pair = blockIdx.z, t = threadIdx.x;
    const size_t plane = (size_t)h * w;
    const float *R0 = R + (size_t)pair * 5 * plane, *R1 = R0 + 5 * plane;
    const float2 *fin = INIT == 0 ? flow_in + (size_t)pair * plane : nullptr;
    const float2 *pv = INIT == 1 ? prev + (size_t)pair * ph * pw : nullptr;
    const int sx0 = blockIdx.x * FI_OUT, y0 = blockIdx.y * rows_per_block;
    const int gx = clampi(sx0 - 8 + t, 0, w - 1);
    const int y_end = min(y0 + rows_per_block, h);
    const int rows_need = (y_end - y0) + 2 * MS_R;                     // M rows y0-7 .. y_end+6 (clamped to the image)
    // x taps of the coarser level's flow are per column: once per thread
    int px0 = 0, px1 = 0;
    float pax = 0.f;
    if (INIT == 1) {
        if (w == 2 * pw) up2_tap_f32(gx, pw, false, px0, px1, pax);
        else lin_tap_f32(gx, pw, w, false, px0, px1, pax);
    }
#define FI_ROW(rr) clampi(y0 - MS_R + (rr), 0, h - 1)
    FiA A;
    FiB B;
    fi_load<INIT>(A, R0, fin, pv, ph, pw, px0, px1, plane, gx, FI_ROW(0), h, w, h);
    fi_issue<INIT>(B, A, R1, plane, gx, FI_ROW(0), h, w, pax);
    fi_load<INIT>(A, R0, fin, pv, ph, pw, px0, px1, plane, gx, FI_ROW(1), h, w, h);
    // produce M of row rr (gathers issued one call earlier), then advance both pipeline stages
    auto produce = [&](int rr, float m[5]) {
        fi_finish(B, gx, FI_ROW(rr), h, w, m);
        if (rr + 1 < rows_need) fi_issue<INIT>(B, A, R1, plane, gx, FI_ROW(rr + 1), h, w, pax);
        if (rr + 2 < rows_need) fi_load<INIT>(A, R0, fin, pv, ph, pw, px0, px1, plane, gx, FI_ROW(rr + 2), h, w, h);
    };
    float *my_ring = ring + t;
    // Segments are anchored at ABSOLUTE rows (window of output row y = rows a in [y, y+14] of a = image row + 7;
    // segment k = a in [15k, 15k+14]), not at the strip start: the two float sums that make a window are then the
    // same for every strip partition, so the flow does not depend on how many pairs a launch carries.
    const int o0 = y0 % FI_SEG;
    float pre[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    // warm-up: rows a = y0 .. y0+14.  The first 15-o0 complete the current segment (slots o0..14, then suffix sums
    // in place), the next o0 are the head of the following segment (slots 0..o0-1, running prefix)
    for (int rr = 0; rr < FI_SEG - o0; rr++) {
        float m[5];
        produce(rr, m);
#pragma unroll
        for (int c = 0; c < 5; c++) my_ring[((o0 + rr) * 5 + c) * FI_W] = m[c];
    }
#pragma unroll
    for (int c = 0; c < 5; c++) {
        float acc = my_ring[((FI_SEG - 1) * 5 + c) * FI_W];
        for (int i = FI_SEG - 2; i >= o0; i--) {
            acc += my_ring[(i * 5 + c) * FI_W];
            my_ring[(i * 5 + c) * FI_W] = acc;
        }
    }
    for (int rr = FI_SEG - o0; rr < FI_SEG; rr++) {
        float m[5];
        produce(rr, m);
#pragma unroll
        for (int c = 0; c < 5; c++) {
            my_ring[((rr - (FI_SEG - o0)) * 5 + c) * FI_W] = m[c];
            pre[c] += m[c];
        }
    }
    // horizontal work item: 16 lanes per plane (14 segments of 8 outputs + 2 idle lanes)
    const int hc = t >> 4, hseg = t & 15;
    const bool hwork = t < 80 && hseg < 14;
    const int ox = t - 8, gxo = sx0 + ox;
    const bool has_out = ox >= 0 && ox < FI_OUT && gxo < w;
    float *my_row = &row[0][ms_sw(t)];
    const float *hrow = row[hwork ? hc : 0];
    int hoff[6];
#pragma unroll
    for (int j = 0; j < 6; j++) hoff[j] = ms_sw((hwork ? hseg : 0) * MS_SEG + 4 * j);
    float *hout0 = &hs[hwork ? hc : 0][ms_sw((hwork ? hseg : 0) * MS_SEG + 8)];
    float *hout1 = &hs[hwork ? hc : 0][ms_sw((hwork ? hseg : 0) * MS_SEG + 12)];
    const float *my_hs = &hs[0][ms_sw(t)];
    float2 *fout = flow_out + (size_t)pair * plane + (size_t)y0 * w + (has_out ? gxo : 0);
    int o = o0;
    double mag_acc = 0;
    for (int y = y0; y < y_end; y++) {
        float *slot = my_ring + o * 5 * FI_W;
#pragma unroll
        for (int c = 0; c < 5; c++) my_row[c * MS_VP] = slot[c * FI_W] + pre[c];
        __syncthreads();
        if (hwork) {
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
            *reinterpret_cast<float4 *>(hout0) = make_float4(o8[0], o8[1], o8[2], o8[3]);
            *reinterpret_cast<float4 *>(hout1) = make_float4(o8[4], o8[5], o8[6], o8[7]);
        }
        __syncthreads();
        if (has_out) {
            // 2x2 solve on the raw window sums (see k_fb_blur_solve): float with FMA-recovered product errors
            const float s11 = my_hs[0], s12 = my_hs[MS_VP], s22 = my_hs[2 * MS_VP], t1 = my_hs[3 * MS_VP], t2 = my_hs[4 * MS_VP];
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
        // the raw row that enters the NEXT window replaces the suffix slot this window just consumed
        const int rr = (y - y0) + FI_SEG;
        if (rr < rows_need) {
            float m[5];
            produce(rr, m);
#pragma unroll
            for (int c = 0; c < 5; c++) {
                slot[c * FI_W] = m[c];
                pre[c] += m[c];
            }
        }
        if (++o == FI_SEG) {                                           // the ring now holds a whole raw segment
            o = 0;
#pragma unroll
            for (int c = 0; c < 5; c++) {
                pre[c] = 0.f;
                float acc = my_ring[((FI_SEG - 1) * 5 + c) * FI_W];
#pragma unroll
                for (int i = FI_SEG - 2; i >= 0; i--) {
                    acc += my_ring[(i * 5 + c) * FI_W];
                    my_ring[(i * 5 + c) * FI_W] = acc;
                }
            }
        }
    }
#undef FI_ROW
    if (mag_sum) {
        __shared__ double red[FI_W / 32];
        mag_acc = warp_sum(mag_acc);
        __syncthreads();
        if ((t & 31) == 0) red[t >> 5] = mag_acc;
        __syncthreads();
        if (t == 0) {
            double sum = 0;
            for (int i = 0; i < FI_W / 32; i++) sum += red[i];
            atomicAdd(&mag_sum[pair], sum);
        }
    }
}


constexpr int FI_H_CAP = 270, FI_H_MIN = 45, FI_WAVES = 3;

// the three iterations of one pyramid level; Q (= `flow`) receives this level's flow, P (= `prev`) holds the coarser one
static int run_fused_level(vqa_ctx *c, const float *R, float2 *Q, float2 *P, int ph, int pw, int lh, int lw, int npairs,
                           bool coarsest, bool finest, double *mag_sum, bool keep_flow)
{
    const int fi_h_cap = (getenv("VQA_FI_H") && atoi(getenv("VQA_FI_H")) >= 16) ? atoi(getenv("VQA_FI_H")) : FI_H_CAP;
    const int fi_waves = (getenv("VQA_FI_WAVES") && atoi(getenv("VQA_FI_WAVES")) >= 1) ? atoi(getenv("VQA_FI_WAVES")) : FI_WAVES;
    VQA_CUDA(c, cudaFuncSetAttribute(k_fb_iter<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FI_SMEM));
    VQA_CUDA(c, cudaFuncSetAttribute(k_fb_iter<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FI_SMEM));
    VQA_CUDA(c, cudaFuncSetAttribute(k_fb_iter<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FI_SMEM));
    // rows per block: tall strips amortise the 14 extra rows of UpdateMatrices a strip evaluates above and below its
    // outputs; the small pyramid levels need shorter strips to fill the SMs
    int rows_pb;
    {
        const long want = (long)fi_waves * c->sm_count * 5, per_row_strip = (long)cdiv(lw, FI_OUT) * npairs;
        const long strips = std::max(1L, (want + per_row_strip - 1) / per_row_strip);
        rows_pb = (int)((lh + strips - 1) / strips);
        if (rows_pb < FI_H_MIN) rows_pb = FI_H_MIN;
        if (rows_pb > fi_h_cap) rows_pb = fi_h_cap;
        rows_pb = (lh + cdiv(lh, rows_pb) - 1) / cdiv(lh, rows_pb);      // equal strips
    }
    const dim3 gI(cdiv(lw, FI_OUT), cdiv(lh, rows_pb), npairs);
    const double px = (double)lw * lh * npairs;
    // iteration 0 writes Q from the coarser flow in P; iteration 1 reads Q and writes P (the coarser flow is no longer
    // needed); iteration 2 reads P and writes Q
    for (int it = 0; it < 3; it++) {
        const bool last = finest && it == 2;
        double *ms = last ? mag_sum : (double *)nullptr;
        const int wf = (!last || keep_flow) ? 1 : 0;
        if (it == 0 && coarsest) {
            VQA_BYTES(c, 48.0 * px);
            VQA_LAUNCH(c, k_fb_iter<2>, gI, FI_W, FI_SMEM, R, (const float2 *)nullptr, lh, lw, Q, rows_pb, ms, wf, (const float2 *)nullptr, 0, 0);
        } else if (it == 0) {
            VQA_BYTES(c, 48.0 * px + 8.0 * pw * ph * npairs);
            VQA_LAUNCH(c, k_fb_iter<1>, gI, FI_W, FI_SMEM, R, (const float2 *)nullptr, lh, lw, Q, rows_pb, ms, wf, P, ph, pw);
        } else {
            const float2 *in = (it == 1) ? Q : P;
            float2 *out = (it == 1) ? P : Q;
            VQA_BYTES(c, (last && !keep_flow ? 48.0 : 56.0) * px);
            VQA_LAUNCH(c, k_fb_iter<0>, gI, FI_W, FI_SMEM, R, in, lh, lw, out, rows_pb, ms, wf, (const float2 *)nullptr, 0, 0);
        }
    }
    return VQA_OK;
}
