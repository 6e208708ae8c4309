// General-size ORB keypoint detection and scoring: what
//   cv2.ORB_create(nfeatures, scaleFactor, nlevels, edgeThreshold).detectAndCompute(gray, None)
// does up to the keypoint list (complexity_metrics.py:385-387 takes len() of it), for any image size.
// The reference hard-wires 64x64, where the pipeline degenerates to four pixels (fast_orb.cu); this
// file is the `orb_size` knob.  Stages, all bit-exact against OpenCV (oracle/orb_oracle.py):
//   k_orb_resize_exact   pyramid level l = INTER_LINEAR_EXACT resize of level l-1 (8.8 fixed point)
//   k_orb_fast           FAST-9/16 score + 3x3 non-maximum suppression inside the edgeThreshold border;
//                        candidates appended to a per-(frame, level) list, score histogram on the side
//   k_orb_score_cut      retainBest(2 * quota) by FAST score: cut value from the 256-bin histogram
//   k_orb_harris         Harris response (7x7 block, integer gradients, float32 expression) of the kept
//   k_orb_select         retainBest(quota) by Harris response with ties: radix select of the quota-th key
// Everything is HBM / latency bound byte work; no tensor cores.
#include <math.h>

#include <algorithm>

#include "vqa_common.cuh"

namespace vqa {
namespace {

constexpr int MAXL = 16;

struct OrbLevels {                       // passed by value to the kernels
    int nlevels;
    int lh[MAXL], lw[MAXL], pitch[MAXL], quota[MAXL];
    float scale[MAXL];                   // orb.cpp getScale: (float)pow((double)scaleFactor, level)
    unsigned long long pyr_off[MAXL];    // byte offset of level l >= 1 inside one frame's pyramid block
    unsigned long long cand_off[MAXL];   // entry offset of level l inside one frame's candidate block
};

struct OrbGeom {
    int h = 0, w = 0, edge = 0, fast_thr = 0, nfeatures = 0, nlevels = 0;
    float sf = 0;
    OrbLevels L;
    size_t pyr_frame = 0, cand_frame = 0;
    size_t tap_off[MAXL] = {0};          // uint32 offset of level l's taps: lw[l] x-taps then lh[l] y-taps
    size_t tap_total = 0;
    bool taps_on_device = false;
};

// ---- host: pyramid geometry, per-level quotas, INTER_LINEAR_EXACT tap tables ---------------------
// orb.cpp: scale_l = (float)pow((double)(float)scaleFactor, l); size = cvRound(dim / scale_l) in float.
void orb_geometry(int h, int w, const vqa_orb_cfg &cfg, OrbGeom &g)
{
    g.h = h; g.w = w; g.edge = cfg.edge_threshold; g.fast_thr = cfg.fast_threshold;
    g.nfeatures = cfg.nfeatures; g.nlevels = cfg.nlevels; g.sf = cfg.scale_factor;
    OrbLevels &L = g.L;
    L.nlevels = cfg.nlevels;
    const double sfd = (double)cfg.scale_factor;
    size_t pyr = 0, cand = 0, taps = 0;
    for (int l = 0; l < cfg.nlevels; l++) {
        const float sc = (float)pow(sfd, (double)l);
        L.scale[l] = sc;
        volatile float fw = (float)w / sc, fh = (float)h / sc;
        L.lw[l] = (int)lrint((double)fw);
        L.lh[l] = (int)lrint((double)fh);
        L.pitch[l] = l == 0 ? w : (L.lw[l] + 15) & ~15;
        L.pyr_off[l] = pyr;
        if (l > 0) pyr += (size_t)L.pitch[l] * L.lh[l];
        L.cand_off[l] = cand;
        const int iw = L.lw[l] - 2 * cfg.edge_threshold, ih = L.lh[l] - 2 * cfg.edge_threshold;
        if (iw > 0 && ih > 0) cand += (size_t)((iw + 1) / 2) * ((ih + 1) / 2);   // 3x3 NMS: no two keypoints touch
        g.tap_off[l] = taps;
        if (l > 0) taps += (size_t)L.lw[l] + L.lh[l];
    }
    g.pyr_frame = (pyr + 255) & ~(size_t)255;
    g.cand_frame = std::max<size_t>(cand, 1);
    g.tap_total = std::max<size_t>(taps, 1);
    // computeKeyPoints: nfeaturesPerLevel, float recurrence
    const float factor = (float)(1.0 / sfd);
    volatile float nd = (float)cfg.nfeatures * (1.f - factor) / (1.f - (float)pow((double)factor, (double)cfg.nlevels));
    int sum = 0;
    for (int l = 0; l < cfg.nlevels - 1; l++) {
        L.quota[l] = (int)lrint((double)nd);
        sum += L.quota[l];
        nd = nd * factor;
    }
    L.quota[cfg.nlevels - 1] = std::max(cfg.nfeatures - sum, 0);
    g.taps_on_device = false;
}

// resize.cpp interpolationLinear<ufixedpoint16>::getCoeffs in IEEE double; packed (offset << 16) | c1,
// c0 = 256 - c1; outside the interpolated range the edge pixel is replicated (c1 = 0).
void exact_taps(int sn, int dn, uint32_t *out)
{
    volatile double scale = 1.0 / ((double)dn / (double)sn);
    for (int d = 0; d < dn; d++) {
        volatile double prod = scale * ((double)d + 0.5);
        volatile double f = prod - 0.5;
        const int i = (int)floor(f);
        uint32_t off = 0, c1 = 0;
        if (i >= 0 && sn > 1) {
            if (i < sn - 1) {
                volatile double fr = f - (double)i;
                volatile double sc = fr * 256.0;
                off = (uint32_t)i;
                c1 = (uint32_t)lrint(sc);
            } else {
                off = (uint32_t)(sn - 1);
            }
        }
        out[d] = (off << 16) | c1;
    }
}

// ---- kernels ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_orb_resize_exact(const uint8_t *__restrict__ src, size_t src_frame, int spitch, int sh, int sw,
                   uint8_t *__restrict__ dst, size_t dst_frame, int dpitch, int dh, int dw,
                   const uint32_t *__restrict__ xt, const uint32_t *__restrict__ yt)
{
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (y >= dh || x4 >= dw) return;
    const uint32_t ty = yt[y];
    const int yo = (int)(ty >> 16), yb = (int)(ty & 0xffff), ya = 256 - yb, y1 = min(yo + 1, sh - 1);
    const uint8_t *r0 = src + (size_t)blockIdx.z * src_frame + (size_t)yo * spitch;
    const uint8_t *r1 = src + (size_t)blockIdx.z * src_frame + (size_t)y1 * spitch;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = x4 + k;
        if (x < dw) {
            const uint32_t tx = xt[x];
            const int xo = (int)(tx >> 16), xb = (int)(tx & 0xffff), xa = 256 - xb, x1 = min(xo + 1, sw - 1);
            const int t0 = (int)r0[xo] * xa + (int)r0[x1] * xb;      // 8.8 fixed point (hlineResize)
            const int t1 = (int)r1[xo] * xa + (int)r1[x1] * xb;
            const uint32_t v = (uint32_t)(t0 * ya + t1 * yb + (1 << 15)) >> 16;   // 16.16 -> u8, round half up (vlineResize)
            out |= v << (8 * k);
        }
    }
    *(uint32_t *)(dst + (size_t)blockIdx.z * dst_frame + (size_t)y * dpitch + x4) = out;
}

// FAST-9 strength on the 16-pixel ring: max over the 16 arcs of 9 contiguous ring pixels of
// min(v - p) (centre brighter) and min(p - v) (centre darker), cv2 cornerScore<16>.  Sliding minimum
// of 9 in log steps (2, 4, 8, +1).  `up` is made opaque to the compiler: ptxas 12.9 mis-folds negated
// operands of VIMNMX3 on sm_100a (profiles/r01_notes.md), so no min/max here ever sees a negation.
__device__ __forceinline__ int fast_strength(const int v, const int (&p)[16])
{
    int dn[16], up[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        dn[k] = v - p[k];
        up[k] = p[k] - v;
        asm volatile("" : "+r"(up[k]));
    }
    int a[16], b[16], a2[16], b2[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { a[k] = min(dn[k], dn[(k + 1) & 15]); b[k] = min(up[k], up[(k + 1) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; k++) { a2[k] = min(a[k], a[(k + 2) & 15]); b2[k] = min(b[k], b[(k + 2) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; k++) { a[k] = min(a2[k], a2[(k + 4) & 15]); b[k] = min(b2[k], b2[(k + 4) & 15]); }
    int best = -512;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        int x = min(a[k], dn[(k + 8) & 15]), y = min(b[k], up[(k + 8) & 15]);
        asm volatile("" : "+r"(x), "+r"(y));
        best = max(best, max(x, y));
    }
    return best;
}

// FAST tile: FTW x FTH outputs per CTA; scores are needed one pixel further out (3x3 NMS) and pixels three
// further still (ring radius) -> 4-pixel halo.  Three phases: (1) every score position takes the cheap
// compass pre-test and the survivors are queued in shared memory, (2) the queue is processed densely (no
// divergence between corner and non-corner lanes), (3) NMS + warp-aggregated append.
constexpr int FTW = 64, FTH = 32;
constexpr int PXH = FTH + 8, PXW = FTW + 8 + 4;          // pixel tile: 72 columns + <= 3 bytes of alignment slack
constexpr int SCH = FTH + 2, SCW = FTW + 2, SCP = FTW + 4;

__global__ void __launch_bounds__(256, 5)
k_orb_fast(const uint8_t *__restrict__ img, size_t frame_stride, int pitch, int lh, int lw, int edge, int thr,
           uint32_t *__restrict__ cand, size_t cand_frame, int *__restrict__ ncand, unsigned *__restrict__ shist,
           int ctr_stride)
{
    __shared__ __align__(16) uint8_t px[PXH][PXW];
    __shared__ uint8_t sc[SCH][SCP];
    __shared__ unsigned short queue[SCH * SCW];
    __shared__ int qn;
    const int frame = blockIdx.z, tid = threadIdx.x;
    const int ox = edge + blockIdx.x * FTW, oy = edge + blockIdx.y * FTH;
    const int xa = (ox - 4) & ~3, sh = (ox - 4) - xa;      // tile columns start at a 4-byte boundary; ox - 4 >= 0
    const uint8_t *src = img + (size_t)frame * frame_stride;
    if (tid == 0) qn = 0;
    if ((((size_t)src | (size_t)pitch) & 3) == 0) {         // 32-bit loads; words past the row end are clamped
        const int maxw = (pitch >> 2) - 1;                  // (they only feed score positions that are skipped)
        for (int i = tid; i < PXH * (PXW / 4); i += 256) {
            const int r = i / (PXW / 4), q = i - r * (PXW / 4);
            const int gy = min(oy - 4 + r, lh - 1), wi = min((xa >> 2) + q, maxw);
            ((uint32_t *)px[r])[q] = ((const uint32_t *)(src + (size_t)gy * pitch))[wi];
        }
    } else {
        for (int i = tid; i < PXH * PXW; i += 256) {
            const int r = i / PXW, q = i - r * PXW;
            const int gy = min(oy - 4 + r, lh - 1), gx = min(xa + q, lw - 1);
            px[r][q] = src[(size_t)gy * pitch + gx];
        }
    }
    __syncthreads();
    // (1) compass pre-test: a 9-arc of the 16-ring always holds two adjacent compass points, and every
    // adjacent pair has one of {top, bottom} and one of {left, right}.  Score position (r, q) <-> pixel
    // tile (r + 3, q + 3 + sh).  64 columns x 4 rows per pass (no index division), the last two columns after.
    auto pretest = [&](int r, int q) {
        const int gy = oy - 1 + r, gx = ox - 1 + q;
        if (gy < lh - 3 && gx < lw - 3) {
            const int c = q + 3 + sh;
            const int v = px[r + 3][c];
            const int d0 = v - px[r + 6][c], d8 = v - px[r][c];
            const unsigned vhi = (d0 > thr) | ((d8 > thr) << 1), vlo = (d0 < -thr) | ((d8 < -thr) << 1);
            if (vhi | vlo) {
                const int d4 = v - px[r + 3][c + 3], d12 = v - px[r + 3][c - 3];
                const bool hhi = (d4 > thr) | (d12 > thr), hlo = (d4 < -thr) | (d12 < -thr);
                if ((vhi && hhi) || (vlo && hlo)) queue[atomicAdd(&qn, 1)] = (unsigned short)(r * SCW + q);
            }
        }
        sc[r][q] = 0;
    };
    for (int r = tid >> 6; r < SCH; r += 4) pretest(r, tid & 63);
    if (tid < 2 * SCH) pretest(tid >> 1, FTW + (tid & 1));
    __syncthreads();
    // (2) full strength of the survivors; ring offsets of cv2's FAST-9/16, clockwise from (0, 3)
    const int RX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int RY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    const int nq = qn;
    for (int j = tid; j < nq; j += 256) {
        const int i = queue[j], r = i / SCW, q = i - r * SCW, c = q + 3 + sh;
        const int v = px[r + 3][c];
        int p[16];
#pragma unroll
        for (int k = 0; k < 16; k++) p[k] = px[r + 3 + RY[k]][c + RX[k]];
        const int st = fast_strength(v, p);
        if (st > thr) sc[r][q] = (uint8_t)(st - 1);
    }
    __syncthreads();
    // (3) 3x3 non-maximum suppression inside the border, append
    int *nc = ncand + (size_t)frame * ctr_stride;
    unsigned *hist = shist + (size_t)frame * ctr_stride * 256;
    uint32_t *list = cand + (size_t)frame * cand_frame;
    const int lane = tid & 31, wrp = tid >> 5;
#pragma unroll
    for (int j = 0; j < (FTH / 8) * (FTW / 32); j++) {
        const int ty = wrp * (FTH / 8) + j / (FTW / 32), tx = lane + 32 * (j % (FTW / 32));
        const int gy = oy + ty, gx = ox + tx;
        const int s = sc[ty + 1][tx + 1];
        bool ok = s > 0 && gy < lh - edge && gx < lw - edge;
        if (ok) {
#pragma unroll
            for (int dy = 0; dy < 3; dy++)
#pragma unroll
                for (int dx = 0; dx < 3; dx++)
                    if ((dy != 1 || dx != 1) && sc[ty + dy][tx + dx] >= s) ok = false;
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (m) {
            int base = 0;
            const int leader = __ffs(m) - 1;
            if (lane == leader) base = atomicAdd(nc, __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (ok) {
                list[base + __popc(m & ((1u << lane) - 1))] = ((uint32_t)s << 24) | ((uint32_t)gy << 12) | (uint32_t)gx;
                atomicAdd(&hist[s], 1u);
            }
        }
    }
}

// retainBest(2 * quota) by FAST score: the cut is the score of the (2 * quota)-th best candidate; every
// candidate with score >= cut survives (KeyPointsFilter::retainBest keeps the ties).
__global__ void k_orb_score_cut(OrbLevels L, const int *__restrict__ ncand, const unsigned *__restrict__ shist,
                                int *__restrict__ cut, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * MAXL) return;
    const int l = i % MAXL;
    if (l >= L.nlevels) return;
    const int want = 2 * L.quota[l], have = ncand[i];
    int c = 0;                                             // keep everything
    if (want <= 0) c = 256;                                // retainBest(0) clears
    else if (have > want) {
        int cum = 0;
        for (int s = 255; s >= 0; s--) {
            cum += (int)shist[(size_t)i * 256 + s];
            if (cum >= want) { c = s; break; }
        }
    }
    cut[i] = c;
}

// orb.cpp HarrisResponses: blockSize 7, k = 0.04, integer gradient sums, float32 expression evaluated
// left to right with no contraction:  (a*b - c*c - k*(a+b)*(a+b)) * scale^4,  scale = 1/(4*7*255).
__global__ void __launch_bounds__(128)
k_orb_harris(OrbLevels L, const uint8_t *__restrict__ img0, size_t frame_stride0, const uint8_t *__restrict__ pyr,
             size_t pyr_frame, const uint32_t *__restrict__ cand, size_t cand_frame, const int *__restrict__ ncand,
             const int *__restrict__ cut, float *__restrict__ resp, uint32_t *__restrict__ rpos, int *__restrict__ nresp,
             float scale_sq_sq)
{
    const int l = blockIdx.y, frame = blockIdx.z, ci = frame * MAXL + l;
    const int have = ncand[ci], c = cut[ci], pitch = L.pitch[l];
    const uint8_t *img = l == 0 ? img0 + (size_t)frame * frame_stride0 : pyr + (size_t)frame * pyr_frame + L.pyr_off[l];
    const uint32_t *list = cand + (size_t)frame * cand_frame + L.cand_off[l];
    float *ro = resp + (size_t)frame * cand_frame + L.cand_off[l];
    uint32_t *po = rpos + (size_t)frame * cand_frame + L.cand_off[l];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < have; i += gridDim.x * blockDim.x) {
        const uint32_t e = list[i];
        if ((int)(e >> 24) < c) continue;
        const int x = (int)(e & 0xfff), y = (int)((e >> 12) & 0xfff);
        int a = 0, b = 0, cc = 0;
        for (int dy = -3; dy <= 3; dy++) {
            const uint8_t *rm = img + (size_t)(y + dy - 1) * pitch + x, *r0 = rm + pitch, *rp = r0 + pitch;
#pragma unroll
            for (int dx = -3; dx <= 3; dx++) {
                const int ix = ((int)r0[dx + 1] - (int)r0[dx - 1]) * 2 + ((int)rm[dx + 1] - (int)rm[dx - 1]) +
                               ((int)rp[dx + 1] - (int)rp[dx - 1]);
                const int iy = ((int)rp[dx] - (int)rm[dx]) * 2 + ((int)rp[dx - 1] - (int)rm[dx - 1]) +
                               ((int)rp[dx + 1] - (int)rm[dx + 1]);
                a += ix * ix;
                b += iy * iy;
                cc += ix * iy;
            }
        }
        const float fa = (float)a, fb = (float)b, fc = (float)cc;
        const float ab = __fadd_rn(fa, fb);
        const float det = __fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc));
        const float tr = __fmul_rn(__fmul_rn(0.04f, ab), ab);
        float r = __fmul_rn(__fsub_rn(det, tr), scale_sq_sq);
        r = __fadd_rn(r, 0.f);                             // -0 -> +0: the selection below orders bit patterns
        const int pos = atomicAdd(&nresp[ci], 1);
        ro[pos] = r;
        po[pos] = e;
    }
}

__device__ __forceinline__ uint32_t order_key(float f)     // larger float <=> larger key
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// retainBest(quota) by Harris response, one CTA per frame, levels in turn: 4-pass radix select of the
// quota-th largest key, then everything >= it is kept (ties survive).  Optionally writes the keypoints.
__global__ void __launch_bounds__(256)
k_orb_select(OrbLevels L, const float *__restrict__ resp, const uint32_t *__restrict__ rpos, size_t cand_frame,
             const int *__restrict__ nresp, int *__restrict__ counts, int *__restrict__ level_counts,
             vqa_keypoint *__restrict__ kps, int kp_cap)
{
    __shared__ int hist[256];
    __shared__ uint32_t s_prefix;
    __shared__ int s_rank, s_kept, s_out;
    const int frame = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) s_out = 0;
    int total = 0;
    for (int l = 0; l < L.nlevels; l++) {
        const int M = nresp[frame * MAXL + l], q = L.quota[l];
        const float *r = resp + (size_t)frame * cand_frame + L.cand_off[l];
        const uint32_t *p = rpos + (size_t)frame * cand_frame + L.cand_off[l];
        uint32_t kth = 0;                                  // keep everything
        if (q <= 0) {
            kth = 0xffffffffu;                             // retainBest(0) clears the level
        } else if (M > q) {
            uint32_t prefix = 0, mask = 0;
            if (tid == 0) s_rank = q;
            for (int shift = 24; shift >= 0; shift -= 8) {
                hist[tid] = 0;
                __syncthreads();
                for (int i = tid; i < M; i += 256) {
                    const uint32_t k = order_key(r[i]);
                    if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255], 1);
                }
                __syncthreads();
                if (tid == 0) {
                    int cum = 0, rank = s_rank, d = 255;
                    for (; d > 0; d--) {
                        if (cum + hist[d] >= rank) break;
                        cum += hist[d];
                    }
                    s_rank = rank - cum;
                    s_prefix = prefix | ((uint32_t)d << shift);
                }
                __syncthreads();
                prefix = s_prefix;
                mask |= 0xffu << shift;
                __syncthreads();
            }
            kth = prefix;
        }
        if (tid == 0) s_kept = 0;
        __syncthreads();
        int mine = 0;
        for (int i = tid; i < M; i += 256) {
            const float v = r[i];
            if (q > 0 && order_key(v) >= kth) {
                mine++;
                if (kps) {
                    const int o = atomicAdd(&s_out, 1);
                    if (o < kp_cap) {
                        const uint32_t e = p[i];
                        const float scale = L.scale[l];
                        vqa_keypoint k;
                        k.lx = (int)(e & 0xfff);
                        k.ly = (int)((e >> 12) & 0xfff);
                        k.octave = l;
                        k.fast_score = (int)(e >> 24);
                        k.response = v;
                        k.x = __fmul_rn((float)k.lx, scale);            // computeKeyPoints: pt *= scale
                        k.y = __fmul_rn((float)k.ly, scale);
                        k.size = __fmul_rn(31.f, scale);                // patchSize * scale
                        k.angle = -1.f;                                 // filled by k_orb_angle
                        kps[(size_t)frame * kp_cap + o] = k;
                    }
                }
            }
        }
        mine = warp_sum(mine);
        if ((tid & 31) == 0 && mine) atomicAdd(&s_kept, mine);
        __syncthreads();
        if (tid == 0 && level_counts) level_counts[frame * MAXL + l] = s_kept;
        total += s_kept;
        __syncthreads();
    }
    if (tid == 0) counts[frame] = total;
}

// orb.cpp ICAngles: orientation of a keypoint = fastAtan2(m01, m10) of the intensity moments over the
// circular patch of radius 15 (row half-widths umax[v]); pixels outside the level image come from its
// reflect-101 border (only reachable when edgeThreshold < 16).  One warp per keypoint, lane v (< 16)
// owns the row pair +-v; integer moments, float32 polynomial in OpenCV's evaluation order.
__device__ __forceinline__ int reflect101(int i, int n)
{
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

__device__ __forceinline__ float fast_atan2_deg(float y, float x, float p1, float p3, float p5, float p7)
{
    const float ax = fabsf(x), ay = fabsf(y), eps = 2.220446049250313e-16f;
    const bool flat = ax >= ay;
    const float c = flat ? __fdiv_rn(ay, __fadd_rn(ax, eps)) : __fdiv_rn(ax, __fadd_rn(ay, eps));
    const float c2 = __fmul_rn(c, c);
    float a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    if (!flat) a = __fsub_rn(90.f, a);
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

__global__ void __launch_bounds__(256)
k_orb_angle(OrbLevels L, const uint8_t *__restrict__ img0, size_t frame_stride0, const uint8_t *__restrict__ pyr,
            size_t pyr_frame, const int *__restrict__ counts, vqa_keypoint *__restrict__ kps, int kp_cap,
            float p1, float p3, float p5, float p7)
{
    const int UMAX[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    const int frame = blockIdx.y, lane = threadIdx.x & 31, ki = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (ki >= min(counts[frame], kp_cap)) return;                       // warp-uniform
    vqa_keypoint *kp = kps + (size_t)frame * kp_cap + ki;
    const int l = kp->octave, x = kp->lx, y = kp->ly, lw = L.lw[l], lh = L.lh[l], pitch = L.pitch[l];
    const uint8_t *img = l == 0 ? img0 + (size_t)frame * frame_stride0 : pyr + (size_t)frame * pyr_frame + L.pyr_off[l];
    const bool inside = x >= 15 && y >= 15 && x + 15 < lw && y + 15 < lh;
    int m01 = 0, m10 = 0;
    if (lane < 16) {
        const int v = lane, d = UMAX[v];
        if (inside) {
            const uint8_t *rp = img + (size_t)(y + v) * pitch + x, *rm = img + (size_t)(y - v) * pitch + x;
            int vs = 0;
            for (int u = -d; u <= d; u++) {
                const int vp = rp[u], vm = rm[u];
                vs += vp - vm;
                m10 += u * (v ? vp + vm : vp);
            }
            m01 = v * vs;
        } else {
            const uint8_t *rp = img + (size_t)reflect101(y + v, lh) * pitch, *rm = img + (size_t)reflect101(y - v, lh) * pitch;
            int vs = 0;
            for (int u = -d; u <= d; u++) {
                const int xx = reflect101(x + u, lw);
                const int vp = rp[xx], vm = rm[xx];
                vs += vp - vm;
                m10 += u * (v ? vp + vm : vp);
            }
            m01 = v * vs;
        }
    }
    m01 = warp_sum(m01);
    m10 = warp_sum(m10);
    if (lane == 0) kp->angle = fast_atan2_deg((float)m01, (float)m10, p1, p3, p5, p7);
}

}  // namespace

// ---- launcher --------------------------------------------------------------------------------------
void orb_release(vqa_ctx *c)
{
    delete (OrbGeom *)c->orb;
    c->orb = nullptr;
}

int orb_check_cfg(vqa_ctx *c, const vqa_orb_cfg &cfg, int h, int w)
{
    if (cfg.nfeatures < 0 || cfg.nlevels < 1 || cfg.nlevels > MAXL || !(cfg.scale_factor > 1.f) ||
        cfg.edge_threshold < 4 || cfg.fast_threshold < 1 || cfg.fast_threshold > 254)
        return set_err(c, VQA_E_INVALID, "vqa_orb_cfg out of range (nlevels 1..%d, scale_factor > 1, edge_threshold >= 4)", MAXL);
    if (h <= 0 || w <= 0 || h > 4096 || w > 4096)
        return set_err(c, VQA_E_UNSUPPORTED, "ORB: image %dx%d outside 1..4096 (12-bit packed coordinates)", w, h);
    return VQA_OK;
}

void orb_defaults(vqa_orb_cfg *cfg)
{
    cfg->nfeatures = 500;
    cfg->nlevels = 8;
    cfg->edge_threshold = 31;
    cfg->fast_threshold = 20;
    cfg->scale_factor = 1.2f;
}

// gray: [n][h] rows of `pitch0` bytes, frames `frame_stride` bytes apart, on the device.
int run_orb_general(vqa_ctx *c, const uint8_t *gray, int n, int h, int w, size_t frame_stride, int pitch0,
                    const vqa_orb_cfg *cfg_in, int *counts /* [n] dev */, int *level_counts /* [n][16] dev or null */,
                    vqa_keypoint *kps /* [n][kp_cap] dev or null */, int kp_cap)
{
    vqa_orb_cfg cfg;
    if (cfg_in) cfg = *cfg_in; else orb_defaults(&cfg);
    int rc = orb_check_cfg(c, cfg, h, w);
    if (rc) return rc;
    if (n <= 0) return VQA_OK;
    OrbGeom *g = (OrbGeom *)c->orb;
    if (!g) { g = new OrbGeom(); c->orb = g; }
    if (g->h != h || g->w != w || g->edge != cfg.edge_threshold || g->fast_thr != cfg.fast_threshold ||
        g->nfeatures != cfg.nfeatures || g->nlevels != cfg.nlevels || g->sf != cfg.scale_factor)
        orb_geometry(h, w, cfg, *g);
    g->L.pitch[0] = pitch0;
    const OrbLevels &L = g->L;
    // levels that can hold a keypoint (runByImageBorder clears a level with a side <= 2 * edgeThreshold)
    int live = 0;
    while (live < L.nlevels && L.lw[live] > 2 * cfg.edge_threshold && L.lh[live] > 2 * cfg.edge_threshold) live++;

    VQA_BUF(c, taps, uint32_t, "orb.taps", g->tap_total);
    if (!g->taps_on_device) {
        uint32_t *ht = (uint32_t *)pinned_buf(c, "orb.taps_host", sizeof(uint32_t) * g->tap_total);
        if (!ht) return VQA_E_NOMEM;
        VQA_CUDA(c, cudaStreamSynchronize(c->stream));           // a previous geometry's upload may still be in flight
        for (int l = 1; l < L.nlevels; l++) {
            exact_taps(L.lw[l - 1], L.lw[l], ht + g->tap_off[l]);
            exact_taps(L.lh[l - 1], L.lh[l], ht + g->tap_off[l] + L.lw[l]);
        }
        VQA_CUDA(c, cudaMemcpyAsync(taps, ht, sizeof(uint32_t) * g->tap_total, cudaMemcpyHostToDevice, c->stream));
        g->taps_on_device = true;
    }
    VQA_BUF(c, pyr, uint8_t, "orb.pyr", g->pyr_frame * n + 256);
    VQA_BUF(c, cand, uint32_t, "orb.cand", g->cand_frame * n);
    VQA_BUF(c, resp, float, "orb.resp", g->cand_frame * n);
    VQA_BUF(c, rpos, uint32_t, "orb.rpos", g->cand_frame * n);
    // counters of one call, zeroed together: ncand[n][16], nresp[n][16], cut[n][16], shist[n][16][256]
    const size_t nctr = (size_t)n * MAXL;
    VQA_BUF(c, ctr, int, "orb.ctr", nctr * (3 + 256));
    int *ncand = ctr, *nresp = ctr + nctr, *cut = ctr + 2 * nctr;
    unsigned *shist = (unsigned *)(ctr + 3 * nctr);
    VQA_CUDA(c, cudaMemsetAsync(ctr, 0, sizeof(int) * nctr * (3 + 256), c->stream));

    for (int l = 1; l < live; l++) {
        const uint8_t *src = l == 1 ? gray : pyr + L.pyr_off[l - 1];
        const size_t sframe = l == 1 ? frame_stride : g->pyr_frame;
        dim3 blk(32, 8), grd(cdiv(cdiv(L.lw[l], 4), 32), cdiv(L.lh[l], 8), n);
        VQA_BYTES(c, (double)n * ((double)L.lw[l - 1] * L.lh[l - 1] + (double)L.lw[l] * L.lh[l]));
        VQA_LAUNCH(c, k_orb_resize_exact, grd, blk, 0, src, sframe, L.pitch[l - 1], L.lh[l - 1], L.lw[l - 1],
                   pyr + L.pyr_off[l], g->pyr_frame, L.pitch[l], L.lh[l], L.lw[l], taps + g->tap_off[l],
                   taps + g->tap_off[l] + L.lw[l]);
    }
    for (int l = 0; l < live; l++) {
        const uint8_t *img = l == 0 ? gray : pyr + L.pyr_off[l];
        const size_t iframe = l == 0 ? frame_stride : g->pyr_frame;
        dim3 grd(cdiv(L.lw[l] - 2 * cfg.edge_threshold, FTW), cdiv(L.lh[l] - 2 * cfg.edge_threshold, FTH), n);
        VQA_BYTES(c, (double)n * L.lw[l] * L.lh[l]);
        VQA_LAUNCH(c, k_orb_fast, grd, 256, 0, img, iframe, L.pitch[l], L.lh[l], L.lw[l], cfg.edge_threshold,
                   cfg.fast_threshold, cand + L.cand_off[l], g->cand_frame, ncand + l, shist + (size_t)l * 256, MAXL);
    }
    if (live > 0) {
        VQA_LAUNCH(c, k_orb_score_cut, cdiv((long)nctr, 128), 128, 0, L, ncand, shist, cut, n);
        const float scale = 1.f / ((float)(1 << 2) * 7 * 255.f);
        const float s4 = scale * scale * scale * scale;
        VQA_LAUNCH(c, k_orb_harris, dim3(16, live, n), 128, 0, L, gray, frame_stride, pyr, g->pyr_frame, cand,
                   g->cand_frame, ncand, cut, resp, rpos, nresp, s4);
    }
    VQA_LAUNCH(c, k_orb_select, n, 256, 0, L, resp, rpos, g->cand_frame, nresp, counts, level_counts, kps, kp_cap);
    if (kps && live > 0) {
        // mathfuncs_core fastAtan2 coefficients, formed in float exactly as OpenCV's translation unit does
        const float rad = (float)(180.0 / 3.1415926535897932384626433832795);
        const float p1 = 0.9997878412794807f * rad, p3 = -0.3258083974640975f * rad;
        const float p5 = 0.1555786518463281f * rad, p7 = -0.04432655554792128f * rad;
        VQA_LAUNCH(c, k_orb_angle, dim3(cdiv(kp_cap, 8), n), 256, 0, L, gray, frame_stride, pyr, g->pyr_frame, counts, kps,
                   kp_cap, p1, p3, p5, p7);
    }
    return VQA_OK;
}

// host-only view of the pyramid geometry (tests: no GPU needed)
int orb_describe(const vqa_orb_cfg *cfg_in, int h, int w, int32_t *level_w, int32_t *level_h, int32_t *quota)
{
    vqa_orb_cfg cfg;
    if (cfg_in) cfg = *cfg_in; else orb_defaults(&cfg);
    if (cfg.nlevels < 1 || cfg.nlevels > MAXL || !(cfg.scale_factor > 1.f) || h <= 0 || w <= 0) return VQA_E_INVALID;
    OrbGeom g;
    orb_geometry(h, w, cfg, g);
    for (int l = 0; l < cfg.nlevels; l++) {
        if (level_w) level_w[l] = g.L.lw[l];
        if (level_h) level_h[l] = g.L.lh[l];
        if (quota) quota[l] = g.L.quota[l];
    }
    return cfg.nlevels;
}

int orb_level_view(vqa_ctx *c, int level, const uint8_t **ptr, int *pitch, int *lh, int *lw)
{
    const OrbGeom *g = (const OrbGeom *)c->orb;
    auto it = c->bufs.find("orb.pyr");
    if (!g || it == c->bufs.end() || level < 1 || level >= g->L.nlevels) return VQA_E_INVALID;
    *ptr = (const uint8_t *)it->second.p + g->L.pyr_off[level];
    *pitch = g->L.pitch[level];
    *lh = g->L.lh[level];
    *lw = g->L.lw[level];
    return VQA_OK;
}

int orb_exact_taps(int sn, int dn, uint32_t *out)
{
    if (sn <= 0 || dn <= 0 || sn > 65535 || !out) return VQA_E_INVALID;
    exact_taps(sn, dn, out);
    return VQA_OK;
}

}  // namespace vqa
