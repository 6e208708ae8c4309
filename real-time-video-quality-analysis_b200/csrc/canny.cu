// cv2.Canny(gray, 100, 200) edge-pixel count, bit-exact (complexity_metrics.py:503-504).
//
//   k_canny_nms     Sobel 3x3 (replicate border) + L1 magnitude (zero ring) + non-maximum
//                   suppression with OpenCV's TG22 fixed-point sectors -> 0 / 1 (weak) / 2 (strong)
//   k_ccl_merge     8-connected union-find over kept pixels (atomicMin on roots)
//   k_ccl_flatten   every kept pixel points at its root; roots of components with a strong pixel get a flag bit
//   k_ccl_count     count (and optionally paint) the kept pixels of flagged components
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
            int *__restrict__ label)
{
    __shared__ uint8_t px[CT_H + 4][CT_W + 4];
    __shared__ int mag[CT_H + 2][CT_W + 2];
    __shared__ uint8_t st[CT_H][CT_W];
    __shared__ int lab[CT_H * CT_W];                 // tile-local union-find (indices inside the tile)
    const int frame = blockIdx.z;
    const uint8_t *g = gray + (size_t)frame * h * w;
    const int tx0 = blockIdx.x * CT_W, ty0 = blockIdx.y * CT_H;
    for (int i = threadIdx.x; i < (CT_H + 4) * (CT_W + 4); i += 256) {
        int y = i / (CT_W + 4), x = i - y * (CT_W + 4);
        int gy = clampi(ty0 - 2 + y, 0, h - 1), gx = clampi(tx0 - 2 + x, 0, w - 1);
        px[y][x] = g[(size_t)gy * w + gx];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (CT_H + 2) * (CT_W + 2); i += 256) {
        int my = i / (CT_W + 2), mx = i - my * (CT_W + 2);
        int iy = ty0 - 1 + my, ix = tx0 - 1 + mx, m = 0;
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
            int cy = my + 1, cx = mx + 1;
            int dx = (px[cy - 1][cx + 1] + 2 * px[cy][cx + 1] + px[cy + 1][cx + 1]) -
                     (px[cy - 1][cx - 1] + 2 * px[cy][cx - 1] + px[cy + 1][cx - 1]);
            int dy = (px[cy + 1][cx - 1] + 2 * px[cy + 1][cx] + px[cy + 1][cx + 1]) -
                     (px[cy - 1][cx - 1] + 2 * px[cy - 1][cx] + px[cy - 1][cx + 1]);
            m = abs(dx) + abs(dy);
        }
        mag[my][mx] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CT_H * CT_W; i += 256) {
        int ty = i / CT_W, tx = i - ty * CT_W;
        int iy = ty0 + ty, ix = tx0 + tx;
        if (iy >= h || ix >= w) { st[ty][tx] = 0; continue; }      // outside the image: never an edge
        int my = ty + 1, mx = tx + 1, m = mag[my][mx];
        uint8_t s = 0;
        if (m > low) {
            int cy = ty + 2, cx = tx + 2;
            int xs = (px[cy - 1][cx + 1] + 2 * px[cy][cx + 1] + px[cy + 1][cx + 1]) -
                     (px[cy - 1][cx - 1] + 2 * px[cy][cx - 1] + px[cy + 1][cx - 1]);
            int ys = (px[cy + 1][cx - 1] + 2 * px[cy + 1][cx] + px[cy + 1][cx + 1]) -
                     (px[cy - 1][cx - 1] + 2 * px[cy - 1][cx] + px[cy - 1][cx + 1]);
            int ax = abs(xs), ay = abs(ys) << 15;
            int t = ax * 13573;
            bool keep;
            if (ay < t) keep = m > mag[my][mx - 1] && m >= mag[my][mx + 1];
            else if (ay > t + (ax << 16)) keep = m > mag[my - 1][mx] && m >= mag[my + 1][mx];
            else {
                int sgn = (xs ^ ys) < 0 ? -1 : 1;
                keep = m > mag[my - 1][mx - sgn] && m > mag[my + 1][mx + sgn];
            }
            if (keep) s = m > high ? 2 : 1;
        }
        st[ty][tx] = s;
        state[(size_t)frame * h * w + (size_t)iy * w + ix] = s;
    }
    __syncthreads();
    // Tile-local connected components in shared memory: (1) label = leftmost pixel of the horizontal
    // run, (2) stitch runs to the row above (N, else NW / NE) with a shared-memory union-find,
    // (3) flatten and publish the GLOBAL index of the local root.  The global union-find
    // (k_ccl_merge) then only has to stitch across tile borders.
    for (int i = threadIdx.x; i < CT_H * CT_W; i += 256) {
        int ty = i / CT_W, tx = i - ty * CT_W;
        int l = -1;
        if (st[ty][tx]) {
            int x0 = tx;
            while (x0 > 0 && st[ty][x0 - 1]) x0--;
            l = ty * CT_W + x0;
        }
        lab[i] = l;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CT_H * CT_W; i += 256) {
        int ty = i / CT_W, tx = i - ty * CT_W;
        if (ty == 0 || !st[ty][tx]) continue;
        const bool west = tx > 0 && st[ty][tx - 1];
        if (st[ty - 1][tx]) {
            if (!(west && st[ty - 1][tx - 1])) sm_union(lab, i, i - CT_W);       // the west pixel already links the same two runs
        } else {
            if (tx > 0 && st[ty - 1][tx - 1] && !west) sm_union(lab, i, i - CT_W - 1);
            if (tx < CT_W - 1 && st[ty - 1][tx + 1]) sm_union(lab, i, i - CT_W + 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CT_H * CT_W; i += 256) {
        int ty = i / CT_W, tx = i - ty * CT_W;
        int iy = ty0 + ty, ix = tx0 + tx;
        if (iy >= h || ix >= w || !st[ty][tx]) continue;
        int r = i;
        while (lab[r] != r) r = lab[r];
        label[(size_t)frame * h * w + (size_t)iy * w + ix] = (ty0 + r / CT_W) * w + tx0 + (r % CT_W);
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
    const int frame = blockIdx.y;
    const uint8_t *s = state + (size_t)frame * h * w;
    int *L = label + (size_t)frame * h * w;
    const int total = h * w;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        if (!s[i]) continue;
        const int y = i / w, x = i - y * w;
        const int tx = x % CT_W, ty = y % CT_H;
        if (ty != 0 && tx != 0 && tx != CT_W - 1) continue;     // interior pixels were stitched inside their tile
        if (tx == 0 && x > 0 && s[i - 1]) uf_union(L, i, i - 1);
        if (y == 0) continue;
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
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&counts[frame], (unsigned long long)cnt);
}

int run_canny(vqa_ctx *c, const uint8_t *gray, int n, int h, int w, unsigned long long *counts, uint8_t *edges_out)
{
    const size_t total = (size_t)h * w;
    if (total > (size_t)LBL_MASK) return set_err(c, VQA_E_UNSUPPORTED, "canny: frame larger than 2^30 pixels");
    VQA_BUF(c, state, uint8_t, "canny.state", total * n);
    VQA_BUF(c, label, int, "canny.label", total * n);
    VQA_CUDA(c, cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * (size_t)n, c->stream));
    dim3 g1(cdiv(w, CT_W), cdiv(h, CT_H), n);
    VQA_BYTES(c, 2.0 * total * n);
    VQA_LAUNCH(c, k_canny_nms, g1, 256, 0, gray, h, w, 100, 200, state, label);
    int bpf = cdiv((long)total, 256 * 8);
    if (bpf < 1) bpf = 1;
    dim3 g2(bpf, n);
    VQA_BYTES(c, 1.0 * total * n);
    VQA_LAUNCH(c, k_ccl_merge, g2, 256, 0, state, h, w, label);
    VQA_BYTES(c, 1.0 * total * n);
    VQA_LAUNCH(c, k_ccl_flatten, g2, 256, 0, state, (int)total, label);
    VQA_BYTES(c, 1.0 * total * n);
    VQA_LAUNCH(c, k_ccl_count, g2, 256, 0, state, (int)total, label, counts, edges_out);
    return VQA_OK;
}

}  // namespace vqa
