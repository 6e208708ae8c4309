// cv2.Canny(gray, 100, 200) edge-pixel count, bit-exact (complexity_metrics.py:503-504).
//
//   k_canny_nms     Sobel 3x3 (replicate border) + L1 magnitude (zero ring) + non-maximum
//                   suppression with OpenCV's TG22 fixed-point sectors -> 0 / 1 (weak) / 2 (strong)
//                   then connected components INSIDE the 64x16 tile (shared-memory union-find); every
//                   kept pixel gets the global index of its tile-local root, and the tile appends its
//                   local roots (index, pixel count, "contains a strong pixel") to a per-frame list
//   k_ccl_merge     global union-find (atomicMin on roots) over the pixels on tile borders only
//   k_ccl_root_flag / k_ccl_root_count   work on the list of local roots (~1 % of the pixels): flag the
//                   global root of every local component with a strong pixel, then add up the sizes
//                   of the local components whose global root is flagged
//   k_ccl_flatten / k_ccl_count   per-pixel versions, only used to paint the edge map (debug tap)
//
// Hysteresis is a connected-components problem: the fix-point is unique, so the count equals
// OpenCV's stack-based flood fill whatever the thread schedule.  Roofline: HBM; algorithmic
// bytes per frame: read HW (gray) + write/read HW (state map) + 4 B per kept pixel of labels.
#include "vqa_common.cuh"

namespace vqa {

constexpr int CT_W = 64, CT_H = 16;           // output tile
constexpr int LBL_FLAG = 0x40000000, LBL_MASK = 0x3fffffff;

// union-find on tile-local labels in shared memory (same atomicMin scheme as the global one)
__device__ __forceinline__ void sm_union(int *L, int a, int b)
{
    volatile int *V = L;                           // other threads update L with atomics
    while (true) {
        while (V[a] != a) a = V[a];
        while (V[b] != b) b = V[b];
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;
    }
}

__global__ void __launch_bounds__(256)
k_canny_nms(const uint8_t *__restrict__ gray, int h, int w, int low, int high, uint8_t *__restrict__ state,
            int *__restrict__ label, int2 *__restrict__ roots, int *__restrict__ nroots, int roots_cap)
{
    // Packed front end: the pixel tile is held as 32-bit words (image columns tx0-4 .. tx0+67), a work
    // item produces 4 adjacent gradient magnitudes from 6 word loads, and the NMS of 4 adjacent pixels
    // reads the stored (dx, dy) instead of re-deriving them (the byte-wise version was ALU-bound, SM 86 %).
    constexpr int PXW = (CT_W + 8) / 4, MGP = CT_W + 4;            // 18 words per pixel row; mag / sobel pitch 68
    __shared__ uint32_t px[CT_H + 4][PXW];
    __shared__ __align__(16) int mag[CT_H + 2][MGP];
    __shared__ __align__(16) int sob[CT_H + 2][MGP];               // (dy << 16) | (dx & 0xffff)
    __shared__ __align__(4) uint8_t st[CT_H][CT_W];
    __shared__ int lab[CT_H * CT_W];                               // tile-local union-find (indices inside the tile)
    __shared__ uint16_t sroots[CT_H * CT_W / 4];                   // local roots: at most one per 2x2 block (8-connectivity)
    __shared__ uint16_t kq[CT_H * CT_W];                           // tile indices of the kept pixels (typically 1-3 % of the tile)
    __shared__ int s_n, s_base, s_q;
    if (threadIdx.x == 0) { s_n = 0; s_q = 0; }
    const int frame = blockIdx.z;
    const uint8_t *g = gray + (size_t)frame * h * w;
    const int tx0 = blockIdx.x * CT_W, ty0 = blockIdx.y * CT_H;
    const bool w4 = (w & 3) == 0;
    const bool xin = w4 && tx0 >= 4 && tx0 + CT_W + 4 <= w;        // all 18 words inside the row
    for (int i = threadIdx.x; i < (CT_H + 4) * PXW; i += 256) {
        const int y = i / PXW, wd = i - y * PXW;
        const uint8_t *row = g + (size_t)clampi(ty0 - 2 + y, 0, h - 1) * w;
        const int gx = tx0 - 4 + wd * 4;
        uint32_t v;
        if (xin) v = __ldg(reinterpret_cast<const uint32_t *>(row + gx));
        else {
            v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) v |= (uint32_t)__ldg(row + clampi(gx + b, 0, w - 1)) << (8 * b);
        }
        px[y][wd] = v;
    }
    __syncthreads();
    // gradient magnitudes (and dx, dy) of mag columns 4q .. 4q+3 of mag row my; mag (my, mx) is image
    // pixel (ty0 - 1 + my, tx0 - 1 + mx), whose tile byte column is mx + 3
    for (int i = threadIdx.x; i < (CT_H + 2) * (MGP / 4); i += 256) {
        const int my = i / (MGP / 4), q = i - my * (MGP / 4);
        int r[3][6];                                               // byte columns 4q+2 .. 4q+7 of tile rows my .. my+2
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const uint32_t w0 = px[my + k][q], w1 = px[my + k][q + 1];
            r[k][0] = (w0 >> 16) & 255; r[k][1] = w0 >> 24;
            r[k][2] = w1 & 255; r[k][3] = (w1 >> 8) & 255; r[k][4] = (w1 >> 16) & 255; r[k][5] = w1 >> 24;
        }
        int cs[6];
#pragma unroll
        for (int j = 0; j < 6; j++) cs[j] = r[0][j] + 2 * r[1][j] + r[2][j];
        const int iy = ty0 - 1 + my;
        int m4[4], s4[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int ix = tx0 - 1 + 4 * q + k;
            const int dx = cs[k + 2] - cs[k];
            const int dy = (r[2][k] + 2 * r[2][k + 1] + r[2][k + 2]) - (r[0][k] + 2 * r[0][k + 1] + r[0][k + 2]);
            const bool in = iy >= 0 && iy < h && ix >= 0 && ix < w;
            m4[k] = in ? abs(dx) + abs(dy) : 0;
            s4[k] = (int)(((unsigned)dy << 16) | ((unsigned)dx & 0xffffu));
        }
        *reinterpret_cast<int4 *>(&mag[my][4 * q]) = make_int4(m4[0], m4[1], m4[2], m4[3]);
        *reinterpret_cast<int4 *>(&sob[my][4 * q]) = make_int4(s4[0], s4[1], s4[2], s4[3]);
    }
    __syncthreads();
    {
        const int ty = threadIdx.x >> 4, q = threadIdx.x & 15;      // pixels tx = 4q .. 4q+3 of tile row ty
        const int iy = ty0 + ty, my = ty + 1;
        int mrow[3][6];                                            // mag columns 4q .. 4q+5 of rows my-1 .. my+1
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int4 a = *reinterpret_cast<const int4 *>(&mag[my - 1 + k][4 * q]);
            const int2 b = *reinterpret_cast<const int2 *>(&mag[my - 1 + k][4 * q + 4]);
            mrow[k][0] = a.x; mrow[k][1] = a.y; mrow[k][2] = a.z; mrow[k][3] = a.w; mrow[k][4] = b.x; mrow[k][5] = b.y;
        }
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int ix = tx0 + 4 * q + k, c = k + 1;             // c: column of this pixel inside mrow
            const int m = mrow[1][c];
            uint32_t sv = 0;
            if (iy < h && ix < w && m > low) {
                const int sd = sob[my][4 * q + c];
                const int xs = (int)(short)(sd & 0xffff), ys = sd >> 16;
                const int ax = abs(xs), ay = abs(ys) << 15;
                const int t = ax * 13573;
                bool keep;
                if (ay < t) keep = m > mrow[1][c - 1] && m >= mrow[1][c + 1];
                else if (ay > t + (ax << 16)) keep = m > mrow[0][c] && m >= mrow[2][c];
                else {
                    const bool neg = (xs ^ ys) < 0;              // sgn = -1: compare with (N, x+1) and (S, x-1)
                    keep = m > (neg ? mrow[0][c + 1] : mrow[0][c - 1]) && m > (neg ? mrow[2][c - 1] : mrow[2][c + 1]);
                }
                if (keep) sv = m > high ? 2u : 1u;
            }
            packed |= sv << (8 * k);
        }
        *reinterpret_cast<uint32_t *>(&st[ty][4 * q]) = packed;
        if (packed) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if ((packed >> (8 * k)) & 255u) kq[atomicAdd(&s_q, 1)] = (uint16_t)(ty * CT_W + 4 * q + k);
        }
        if (iy < h) {
            const int ix0 = tx0 + 4 * q;
            uint8_t *dst = state + (size_t)frame * h * w + (size_t)iy * w + ix0;
            if (w4 && ix0 + 4 <= w) *reinterpret_cast<uint32_t *>(dst) = packed;
            else
                for (int k = 0; k < 4; k++)
                    if (ix0 + k < w) dst[k] = (uint8_t)(packed >> (8 * k));
        }
    }
    __syncthreads();
    // Tile-local connected components in shared memory: (1) label = leftmost pixel of the horizontal
    // run, (2) stitch runs to the row above (N, else NW / NE) with a shared-memory union-find,
    // (3) flatten and publish the GLOBAL index of the local root.  The global union-find
    // (k_ccl_merge) then only has to stitch across tile borders.
    // All three passes run over the list of kept pixels only (the per-pixel passes over the whole tile
    // were 62 % of this kernel's instructions for ~2 % of useful pixels).
    int *ssize = &mag[0][0], *sstrong = &sob[0][0];                // per local root: pixel count / has a strong pixel (mag, sob are dead)
    const int nq = s_q;
    for (int j = threadIdx.x; j < nq; j += 256) {
        const int i = kq[j], ty = i / CT_W, tx = i - ty * CT_W;
        int x0 = tx;
        while (x0 > 0 && st[ty][x0 - 1]) x0--;
        lab[i] = ty * CT_W + x0;
        ssize[i] = 0;
        sstrong[i] = 0;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nq; j += 256) {
        const int i = kq[j], ty = i / CT_W, tx = i - ty * CT_W;
        if (ty == 0) continue;
        const bool west = tx > 0 && st[ty][tx - 1];
        if (st[ty - 1][tx]) {
            if (!(west && st[ty - 1][tx - 1])) sm_union(lab, i, i - CT_W);       // the west pixel already links the same two runs
        } else {
            if (tx > 0 && st[ty - 1][tx - 1] && !west) sm_union(lab, i, i - CT_W - 1);
            if (tx < CT_W - 1 && st[ty - 1][tx + 1]) sm_union(lab, i, i - CT_W + 1);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nq; j += 256) {
        const int i = kq[j], ty = i / CT_W, tx = i - ty * CT_W;
        const int iy = ty0 + ty, ix = tx0 + tx;                     // inside the image: pixels outside are never kept
        int r = i;
        while (lab[r] != r) r = lab[r];
        label[(size_t)frame * h * w + (size_t)iy * w + ix] = (ty0 + r / CT_W) * w + tx0 + (r % CT_W);
        atomicAdd(&ssize[r], 1);
        if (st[ty][tx] == 2) sstrong[r] = 1;
        if (r == i) sroots[atomicAdd(&s_n, 1)] = (uint16_t)i;
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_n) s_base = atomicAdd(&nroots[frame], s_n);
    __syncthreads();
    for (int j = threadIdx.x; j < s_n; j += 256) {
        const int i = sroots[j];
        roots[(size_t)frame * roots_cap + s_base + j] =
            make_int2((ty0 + i / CT_W) * w + tx0 + (i % CT_W), ssize[i] | (sstrong[i] ? (int)0x80000000 : 0));
    }
}

__device__ __forceinline__ int uf_find(const int *L, int x)
{
    int p;   // L2 loads: labels are updated with atomics by other SMs (a stale L1 line would only cost extra hops)
    while ((p = (__ldcg(L + x) & LBL_MASK)) != x) x = p;
    return x;
}

__device__ __forceinline__ void uf_union(int *L, int a, int b)
{
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;
    }
}

__global__ void __launch_bounds__(256)
k_ccl_merge(const uint8_t *__restrict__ state, int h, int w, int *__restrict__ label)
{
    // one thread per pixel on a tile border: 64 of the top row, 15 + 15 of the left / right column
    const int frame = blockIdx.y;
    const uint8_t *s = state + (size_t)frame * h * w;
    int *L = label + (size_t)frame * h * w;
    const int tiles_x = (w + CT_W - 1) / CT_W, tiles = tiles_x * ((h + CT_H - 1) / CT_H);
    const int t = blockIdx.x * 256 + threadIdx.x, tile = t / 96, k = t - tile * 96;
    if (tile >= tiles || k >= CT_W + 2 * (CT_H - 1)) return;
    const int ty = k < CT_W ? 0 : (k < CT_W + CT_H - 1 ? k - (CT_W - 1) : k - (CT_W + CT_H - 2));
    const int tx = k < CT_W ? k : (k < CT_W + CT_H - 1 ? 0 : CT_W - 1);
    const int y = (tile / tiles_x) * CT_H + ty, x = (tile % tiles_x) * CT_W + tx;
    if (y >= h || x >= w) return;
    const int i = y * w + x;
    if (!s[i]) return;
    if (tx == 0 && x > 0 && s[i - 1]) uf_union(L, i, i - 1);
    if (y == 0) return;
    const bool north = s[i - w] != 0;
    if (ty == 0) {                                           // row above lies in another tile
        if (north) uf_union(L, i, i - w);
        else {
            if (x > 0 && s[i - w - 1]) uf_union(L, i, i - w - 1);
            if (x < w - 1 && s[i - w + 1]) uf_union(L, i, i - w + 1);
        }
    } else if (!north) {                                     // only the diagonal neighbour across the side border
        if (tx == 0 && x > 0 && s[i - w - 1]) uf_union(L, i, i - w - 1);
        if (tx == CT_W - 1 && x < w - 1 && s[i - w + 1]) uf_union(L, i, i - w + 1);
    }
}

// local roots with a strong pixel flag their global root (a root's own entry is only ever touched by
// atomics after the merge, so the flag survives; readers mask it)
__global__ void __launch_bounds__(256)
k_ccl_root_flag(const int2 *__restrict__ roots, const int *__restrict__ nroots, int roots_cap, int total, int *__restrict__ label)
{
    const int frame = blockIdx.y, n = nroots[frame];
    const int2 *R = roots + (size_t)frame * roots_cap;
    int *L = label + (size_t)frame * total;
    for (int j = blockIdx.x * 256 + threadIdx.x; j < n; j += gridDim.x * 256) {
        const int2 e = R[j];
        if (e.y >= 0) continue;                                // no strong pixel in this local component
        const int g = uf_find(L, e.x);
        if (!(__ldcg(L + g) & LBL_FLAG)) atomicOr(&L[g], LBL_FLAG);
    }
}

__global__ void __launch_bounds__(256)
k_ccl_root_count(const int2 *__restrict__ roots, const int *__restrict__ nroots, int roots_cap, int total,
                 const int *__restrict__ label, unsigned long long *__restrict__ counts)
{
    const int frame = blockIdx.y, n = nroots[frame];
    const int2 *R = roots + (size_t)frame * roots_cap;
    const int *L = label + (size_t)frame * total;
    int cnt = 0;
    for (int j = blockIdx.x * 256 + threadIdx.x; j < n; j += gridDim.x * 256) {
        const int2 e = R[j];
        const int g = uf_find(L, e.x);
        if (__ldcg(L + g) & LBL_FLAG) cnt += e.y & 0x7fffffff;
    }
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&counts[frame], (unsigned long long)cnt);
}

// every kept pixel points at its root; roots of components that contain a strong pixel get the flag
// (a root's own entry is never rewritten by the flatten, so the flag survives; readers mask it)
__global__ void __launch_bounds__(256)
k_ccl_flatten(const uint8_t *__restrict__ state, int total, int *__restrict__ label)
{
    const int frame = blockIdx.y;
    const uint8_t *s = state + (size_t)frame * total;
    int *L = label + (size_t)frame * total;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        const uint8_t st = s[i];
        if (!st) continue;
        const int r = uf_find(L, i);
        if (r != i) L[i] = r;
        if (st == 2 && !(__ldcg(L + r) & LBL_FLAG)) atomicOr(&L[r], LBL_FLAG);
    }
}

__global__ void __launch_bounds__(256)
k_ccl_count(const uint8_t *__restrict__ state, int total, const int *__restrict__ label,
            unsigned long long *__restrict__ counts, uint8_t *__restrict__ edges)
{
    const int frame = blockIdx.y;
    const uint8_t *s = state + (size_t)frame * total;
    const int *L = label + (size_t)frame * total;
    int cnt = 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        bool e = false;
        if (s[i]) e = (L[L[i] & LBL_MASK] & LBL_FLAG) != 0;
        cnt += e;
        if (edges) edges[(size_t)frame * total + i] = e ? 255 : 0;
    }
    cnt = warp_sum(cnt);
    if (counts && (threadIdx.x & 31) == 0 && cnt) atomicAdd(&counts[frame], (unsigned long long)cnt);
}

int run_canny(vqa_ctx *c, const uint8_t *gray, int n, int h, int w, unsigned long long *counts, uint8_t *edges_out)
{
    const size_t total = (size_t)h * w;
    if (total > (size_t)LBL_MASK) return set_err(c, VQA_E_UNSUPPORTED, "canny: frame larger than 2^30 pixels");
    VQA_BUF(c, state, uint8_t, "canny.state", total * n);
    VQA_BUF(c, label, int, "canny.label", total * n);
    const int tiles = cdiv(w, CT_W) * cdiv(h, CT_H), roots_cap = tiles * (CT_H * CT_W / 4);
    VQA_BUF(c, roots, int2, "canny.roots", (size_t)roots_cap * n);
    VQA_BUF(c, nroots, int, "canny.nroots", n);
    VQA_CUDA(c, cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * (size_t)n, c->stream));
    VQA_CUDA(c, cudaMemsetAsync(nroots, 0, sizeof(int) * (size_t)n, c->stream));
    dim3 g1(cdiv(w, CT_W), cdiv(h, CT_H), n);
    VQA_BYTES(c, 2.0 * total * n);
    VQA_LAUNCH(c, k_canny_nms, g1, 256, 0, gray, h, w, 100, 200, state, label, roots, nroots, roots_cap);
    VQA_BYTES(c, 96.0 * tiles * n);
    VQA_LAUNCH(c, k_ccl_merge, dim3(cdiv(tiles * 96, 256), n), 256, 0, state, h, w, label);
    const dim3 gR(std::min(64, cdiv(roots_cap, 256)), n);
    VQA_LAUNCH(c, k_ccl_root_flag, gR, 256, 0, roots, nroots, roots_cap, (int)total, label);
    VQA_LAUNCH(c, k_ccl_root_count, gR, 256, 0, roots, nroots, roots_cap, (int)total, label, counts);
    if (edges_out) {                                           // debug tap: paint the map with the per-pixel passes
        int bpf = cdiv((long)total, 256 * 8);
        if (bpf < 1) bpf = 1;
        dim3 g2(bpf, n);
        VQA_LAUNCH(c, k_ccl_flatten, g2, 256, 0, state, (int)total, label);
        VQA_LAUNCH(c, k_ccl_count, g2, 256, 0, state, (int)total, label, (unsigned long long *)nullptr, edges_out);
    }
    return VQA_OK;
}

}  // namespace vqa
