// 2-D DCT as a batched dense contraction on the 5th-generation tensor cores:
//     C = D_h * X * D_w^T      (cv2.dct of process_dct_frame / process_temporal_dct_frame,
//                               complexity_metrics.py:346-364, 543-579)
//
//   GEMM 1 (MODE 1):  T^T[w x h] = D_w[w x w] * X^T          A = D_w (bf16 hi + lo), B = X (exact in bf16)
//   GEMM 2 (MODE 2):  C  [h x w] = D_h[h x h] * T            A = D_h (hi + lo),      B = T^T (hi + lo)
//
// Precision: a single bf16 pass misses the 1e-4 bar, so D and the intermediate T are
// split into two bf16 terms and the cross terms are accumulated in the same fp32 TMEM accumulator:
// GEMM 1 = hi*X + lo*X, GEMM 2 = hi*hi + hi*lo + lo*hi (the lo*lo term is < 2^-16 relative).
//
// Kernel: persistent, warp-specialised, one CTA per SM, 192 threads:
//   warp 0    TMA producer   cp.async.bulk.tensor (SWIZZLE_128B, zero fill for M/N/K tails) -> smem ring
//   warp 1    MMA issuer     tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16, accumulators in TMEM
//                            (two accumulator buffers so the epilogue of tile i overlaps tile i+1)
//   warps 2-5 epilogue       tcgen05.ld 32x32b.x32 -> registers -> (MODE 1) bf16 hi/lo split of T^T,
//                            (MODE 2) fp32 coefficients + fused sum(C^2) energy
// Roofline: tensor pipe (bf16).  Issued flops per frame: 2*2*w*w*h + 3*2*h*h*w.
#include <cuda.h>
#include <cuda_bf16.h>

#include "vqa_common.cuh"

namespace vqa {

constexpr int UM_BM = 128, UM_BK = 64, UM_THREADS = 192;


// BK = K elements per shared-memory stage: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B: half the
// bytes per stage, so a wider N tile still has >= 4 stages in flight)
template <int BN, int MODE, int BK = UM_BK>
struct UmmaCfg {
    static constexpr int A_PARTS = 2;
    static constexpr int B_PARTS = MODE == 1 ? 1 : 2;
    static constexpr int A_BYTES = UM_BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_PARTS * A_BYTES + B_PARTS * B_BYTES;
    static constexpr int STAGES = (208 * 1024) / STAGE_BYTES;
    static constexpr int TMEM_COLS = 2 * BN <= 256 ? 256 : 512;      // two accumulators; allocations are powers of two
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

// ------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must become a launch failure, never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 6000000000LL) {
            printf("vqa dct_umma: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
                   threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 UMMA): start address >> 4,
// LBO = 0 (one 128-byte atom along K), SBO = 8 rows * 128 B = 1024 B, version 1, layout type 2.
// BK = 32: SWIZZLE_64B, SBO = 8 rows * 64 B = 512 B, layout type 4.
template <int BK = UM_BK>
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr)
{
    const uint32_t lo = (smem_addr & 0x3FFFFu) >> 4;
    const uint32_t hi = BK == 64 ? (64u | (1u << 14) | (2u << 29)) : (32u | (1u << 14) | (4u << 29));
    return ((uint64_t)hi << 32) | lo;
}

// instruction descriptor: D = F32, A = B = BF16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct UmmaOut {
    __nv_bfloat16 *hi, *lo;      // MODE 1: T^T split, [frame][M][ld]
    int ld;
    size_t frame_stride;         // elements
    float *C;                    // MODE 2: coefficients [frame][M][ldc]
    int ldc;
    size_t c_frame_stride;
    double *energy;              // MODE 2: [frame]
};

template <int BN, int MODE, int BK = UM_BK>
__global__ void __launch_bounds__(UM_THREADS, 1)
k_dct_umma(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
           const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1, int M, int N, int K,
           int nframes, UmmaOut out)
{
    using Cfg = UmmaCfg<BN, MODE, BK>;
    constexpr int S = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S * Cfg::STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + S, *tfull = bars + 2 * S, *tempty = bars + 2 * S + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * S + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
        for (int a = 0; a < 2; a++) { mbar_init(smem_u32(&tfull[a]), 1); mbar_init(smem_u32(&tempty[a]), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)Cfg::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int m_tiles = (M + UM_BM - 1) / UM_BM, n_tiles = (N + BN - 1) / BN, k_blocks = (K + BK - 1) / BK;
    const int tiles_per_frame = m_tiles * n_tiles, total = tiles_per_frame * nframes;
    const uint32_t smem_base = smem_u32(smem);

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        int it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const int frame = tile / tiles_per_frame, r = tile - frame * tiles_per_frame;
            const int nb = r / m_tiles, mb = r - nb * m_tiles;
            for (int kb = 0; kb < k_blocks; kb++, it++) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
                if (lane == 0) {
                    const uint32_t fb = smem_u32(&full[s]);
                    const uint32_t st = smem_base + (uint32_t)s * Cfg::STAGE_BYTES;
                    mbar_expect_tx(fb, Cfg::STAGE_BYTES);
                    tma_load_2d(st, &tmA0, fb, kb * BK, mb * UM_BM);
                    tma_load_2d(st + Cfg::A_BYTES, &tmA1, fb, kb * BK, mb * UM_BM);
                    tma_load_3d(st + 2 * Cfg::A_BYTES, &tmB0, fb, kb * BK, nb * BN, frame);
                    if (MODE == 2) tma_load_3d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &tmB1, fb, kb * BK, nb * BN, frame);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc(UM_BM, BN);
        int it = 0, tl = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, tl++) {
            const int a = tl & 1;
            const uint32_t aph = (uint32_t)(tl >> 1) & 1u;
            mbar_wait(smem_u32(&tempty[a]), aph ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
            for (int kb = 0; kb < k_blocks; kb++, it++) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                mbar_wait(smem_u32(&full[s]), ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t st = smem_base + (uint32_t)s * Cfg::STAGE_BYTES;
                    const uint64_t a0 = umma_desc<BK>(st), a1 = umma_desc<BK>(st + Cfg::A_BYTES);
                    const uint64_t b0 = umma_desc<BK>(st + 2 * Cfg::A_BYTES), b1 = umma_desc<BK>(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++) {
                        const uint64_t adv = (uint64_t)(k * 2);          // 16 bf16 = 32 bytes = 2 x 16-byte units
                        const uint32_t first = (kb | k) ? 1u : 0u;
                        umma_bf16(d_tmem, a0 + adv, b0 + adv, idesc, first);
                        if (MODE == 1) {
                            umma_bf16(d_tmem, a1 + adv, b0 + adv, idesc, 1u);
                        } else {
                            umma_bf16(d_tmem, a0 + adv, b1 + adv, idesc, 1u);
                            umma_bf16(d_tmem, a1 + adv, b0 + adv, idesc, 1u);
                        }
                    }
                    umma_commit(smem_u32(&empty[s]));                    // frees the smem stage when the MMAs retire
                    if (kb == k_blocks - 1) umma_commit(smem_u32(&tfull[a]));
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;                                          // TMEM lane quadrant of this warp
        int tl = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, tl++) {
            const int frame = tile / tiles_per_frame, r = tile - frame * tiles_per_frame;
            const int nb = r / m_tiles, mb = r - nb * m_tiles;
            const int a = tl & 1;
            const uint32_t aph = (uint32_t)(tl >> 1) & 1u;
            mbar_wait(smem_u32(&tfull[a]), aph);
            tc_fence_after();
            const int row = mb * UM_BM + q * 32 + lane;
            double e = 0;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c0), v);
                const int col0 = nb * BN + c0;
                if (row < M && col0 < N) {
                    if (MODE == 1) {
                        __nv_bfloat16 *ph_ = out.hi + (size_t)frame * out.frame_stride + (size_t)row * out.ld + col0;
                        __nv_bfloat16 *pl_ = out.lo + (size_t)frame * out.frame_stride + (size_t)row * out.ld + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint32_t hw[4], lw[4];
#pragma unroll
                            for (int t = 0; t < 4; t++) {
                                const float x0 = __uint_as_float(v[j + 2 * t]), x1 = __uint_as_float(v[j + 2 * t + 1]);
                                const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                                const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
                                const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
                                hw[t] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                                lw[t] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                            }
                            if (col0 + j + 8 <= N) {
                                *reinterpret_cast<uint4 *>(ph_ + j) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                                *reinterpret_cast<uint4 *>(pl_ + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                            } else {
                                for (int t = 0; t < 8; t++)
                                    if (col0 + j + t < N) {
                                        ph_[j + t] = __ushort_as_bfloat16((unsigned short)(hw[t >> 1] >> ((t & 1) * 16)));
                                        pl_[j + t] = __ushort_as_bfloat16((unsigned short)(lw[t >> 1] >> ((t & 1) * 16)));
                                    }
                            }
                        }
                    } else {
                        float *pc = out.C + (size_t)frame * out.c_frame_stride + (size_t)row * out.ldc + col0;
                        const bool vec = (out.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(out.C) & 15) == 0) &&
                                         ((out.c_frame_stride & 3) == 0);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float x0 = __uint_as_float(v[j]), x1 = __uint_as_float(v[j + 1]);
                            const float x2 = __uint_as_float(v[j + 2]), x3 = __uint_as_float(v[j + 3]);
                            if (vec && col0 + j + 4 <= N) {
                                *reinterpret_cast<float4 *>(pc + j) = make_float4(x0, x1, x2, x3);
                                e += (double)x0 * x0 + (double)x1 * x1 + (double)x2 * x2 + (double)x3 * x3;
                            } else {
                                const float xs[4] = {x0, x1, x2, x3};
                                for (int t = 0; t < 4; t++)
                                    if (col0 + j + t < N) { pc[j + t] = xs[t]; e += (double)xs[t] * xs[t]; }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            if (MODE == 2) {
                e = warp_sum(e);
                if (lane == 0 && e != 0) atomicAdd(&out.energy[frame], e);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tempty[a]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                     : "memory");
    }
}

// ---------------------------------------------------------------------- operand preparation
// D[k][i] split into bf16 hi + lo, row pitch ld
__global__ void k_dct_basis_split(int n, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo, int ld)
{
    const long total = (long)n * n;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int k = (int)(e / n), i = (int)(e - (long)k * n);
        double v;
        if (k == 0) v = sqrt(1.0 / n);
        else {
            long m = ((long)(2 * i + 1) * k) % (4L * n);
            v = sqrt(2.0 / n) * cospi((double)m / (2.0 * n));
        }
        const float f = (float)v;
        const __nv_bfloat16 h = __float2bfloat16_rn(f);
        hi[(size_t)k * ld + i] = h;
        lo[(size_t)k * ld + i] = __float2bfloat16_rn((float)(v - (double)__bfloat162float(h)));
    }
}

// Per-frame integer mean m = round(sum x / (h w)).  The contraction runs on x - m (still exact
// integers in bf16): every coefficient except DC is unchanged, and the DC path -- thousands of
// same-sign products whose truncating fp32 tensor-core accumulation biased the energy by 7.6e-5 at
// 4K -- now sums small mixed-sign terms.  k_dc_fix adds m * sqrt(h w) back afterwards.
__global__ void __launch_bounds__(256)
k_frame_sum(const uint8_t *__restrict__ x, long per_frame, unsigned long long *__restrict__ sums)
{
    const int frame = blockIdx.y;
    const uint8_t *p = x + (size_t)frame * per_frame;
    unsigned long long acc = 0;
    const long nvec = ((reinterpret_cast<uintptr_t>(p) & 15) == 0) ? per_frame / 16 : 0;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (long)gridDim.x * 256) {
        const uint4 v = ld_stream(reinterpret_cast<const uint4 *>(p) + i);
        unsigned s = __dp4a(v.x, 0x01010101u, 0u);
        s = __dp4a(v.y, 0x01010101u, s);
        s = __dp4a(v.z, 0x01010101u, s);
        s = __dp4a(v.w, 0x01010101u, s);
        acc += s;
    }
    for (long i = nvec * 16 + (long)blockIdx.x * 256 + threadIdx.x; i < per_frame; i += (long)gridDim.x * 256) acc += p[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&sums[frame], acc);
}

// uint8 [n][h][w] -> bf16 [n][h][ld] of (x - mean)   (integers in [-255, 255] are exact in bf16).
// Rows are converted 8 pixels per thread (64-bit load, 128-bit store) when w % 8 == 0.
__global__ void __launch_bounds__(256)
k_u8_to_bf16(const uint8_t *__restrict__ x, int h, int w, int ld, const unsigned long long *__restrict__ sums,
             __nv_bfloat16 *__restrict__ y)
{
    const int frame = blockIdx.y;
    const uint8_t *s = x + (size_t)frame * h * w;
    __nv_bfloat16 *d = y + (size_t)frame * h * ld;
    const long total = (long)h * w;
    const int m = (int)((sums[frame] + (unsigned long long)(total / 2)) / (unsigned long long)total);
    if ((w & 7) == 0 && (reinterpret_cast<uintptr_t>(s) & 7) == 0) {
        const int wq = w >> 3;
        const long nq = (long)h * wq;
        for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < nq; i += (long)gridDim.x * 256) {
            const int r = (int)(i / wq), cq = (int)(i - (long)r * wq);
            const uint2 v = *reinterpret_cast<const uint2 *>(s + (size_t)r * w + cq * 8);
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t word = k < 2 ? v.x : v.y;
                const int b0 = (int)((word >> ((k & 1) * 16)) & 255u) - m, b1 = (int)((word >> ((k & 1) * 16 + 8)) & 255u) - m;
                o[k] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn((float)b0)) |
                       ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn((float)b1)) << 16);
            }
            *reinterpret_cast<uint4 *>(d + (size_t)r * ld + cq * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        return;
    }
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const int r = (int)(i / w), c = (int)(i - (long)r * w);
        d[(size_t)r * ld + c] = __float2bfloat16_rn((float)((int)s[i] - m));
    }
}

__global__ void k_dc_fix(float *__restrict__ coef, size_t frame_stride, int h, int w, int n,
                         const unsigned long long *__restrict__ sums, double *__restrict__ energy)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const long total = (long)h * w;
    const int m = (int)((sums[f] + (unsigned long long)(total / 2)) / (unsigned long long)total);
    const double c_old = (double)coef[(size_t)f * frame_stride];
    const float c_new = (float)(c_old + (double)m * sqrt((double)total));
    coef[(size_t)f * frame_stride] = c_new;
    energy[f] += (double)c_new * (double)c_new - c_old * c_old;
}

// ------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct UmmaState {
    PFN_encodeTiled encode = nullptr;
    int basis_w = 0, basis_h = 0;
    bool attr_set = false;
};

static int get_state(vqa_ctx *c, UmmaState **out)
{
    if (!c->umma) {
        UmmaState *s = new UmmaState();
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
            delete s;
            return set_err(c, VQA_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
        }
        s->encode = (PFN_encodeTiled)fn;
        c->umma = s;
    }
    *out = (UmmaState *)c->umma;
    return VQA_OK;
}

void dct_umma_release(vqa_ctx *c)
{
    delete (UmmaState *)c->umma;
    c->umma = nullptr;
}

// rows x K bf16, row pitch ld elements; optional frame dimension.
static int make_map(vqa_ctx *c, UmmaState *s, CUtensorMap *m, const void *base, int K, int rows, int ld, int frames,
                    size_t frame_stride_elems, int box_rows, int bk = UM_BK)
{
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(frames > 0 ? frames : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)frame_stride_elems * 2};
    cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const int rank = frames > 0 ? 3 : 2;
    CUresult r = s->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(c, VQA_E_CUDA, "cuTensorMapEncodeTiled failed (%d): K=%d rows=%d ld=%d", (int)r, K, rows, ld);
    return VQA_OK;
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

constexpr bool DCT_GEMM2_WIDE = false;             // development build: GEMM 2 on wide tiles with 32-element K stages (see run_dct_umma)

int run_dct_umma(vqa_ctx *c, const uint8_t *x, int n, int h, int w, float *coef, double *energy,
                 const unsigned long long *pixel_sums)
{
    // N tile per GEMM (measured, 48 x 1080p): GEMM 1 with BN = 256 (3 stages of 64 KB) runs at
    // 1356 TFLOP/s issued vs 1110 with BN = 128; GEMM 2 with BN = 256 has room for 2 stages only
    // (96 KB each) and is no faster than BN = 128 with 3 stages (1097 vs 1128 TFLOP/s).
    constexpr int BN1 = 256, BN2 = 128;      // BN2 = 256 (2 smem stages of 96 KB) measured 4 % slower
    UmmaState *s;
    int rc = get_state(c, &s);
    if (rc) return rc;
#ifdef VQA_AB
    const bool gemm2_wide = getenv("VQA_DCT_G2W") ? atoi(getenv("VQA_DCT_G2W")) != 0 : DCT_GEMM2_WIDE;
#endif
    const int ldw = round_up(w, 64), ldh = round_up(h, 64);
    VQA_BUF(c, Dw_hi, __nv_bfloat16, "umma.Dw_hi", (size_t)w * ldw);
    VQA_BUF(c, Dw_lo, __nv_bfloat16, "umma.Dw_lo", (size_t)w * ldw);
    VQA_BUF(c, Dh_hi, __nv_bfloat16, "umma.Dh_hi", (size_t)h * ldh);
    VQA_BUF(c, Dh_lo, __nv_bfloat16, "umma.Dh_lo", (size_t)h * ldh);
    VQA_BUF(c, X, __nv_bfloat16, "umma.X", (size_t)n * h * ldw);
    VQA_BUF(c, Tt_hi, __nv_bfloat16, "umma.Tt_hi", (size_t)n * w * ldh);
    VQA_BUF(c, Tt_lo, __nv_bfloat16, "umma.Tt_lo", (size_t)n * w * ldh);
    if (s->basis_w != w || s->basis_h != h) {          // (buffers are reallocated only when they grow: regenerate on any change)
        VQA_LAUNCH(c, k_dct_basis_split, 296, 256, 0, w, Dw_hi, Dw_lo, ldw);
        VQA_LAUNCH(c, k_dct_basis_split, 296, 256, 0, h, Dh_hi, Dh_lo, ldh);
        s->basis_w = w;
        s->basis_h = h;
    }
    if (!s->attr_set) {
        VQA_CUDA(c, cudaFuncSetAttribute(k_dct_umma<BN1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<BN1, 1>::SMEM_BYTES));
        VQA_CUDA(c, cudaFuncSetAttribute(k_dct_umma<BN2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<BN2, 2>::SMEM_BYTES));
        s->attr_set = true;
    }
    VQA_BUF(c, sums_buf, unsigned long long, "umma.sums", n);
    const unsigned long long *sums = pixel_sums;            // the caller may already know them (gray histogram moments)
    int bpf = cdiv((long)h * w, 256 * 8);
    if (bpf < 1) bpf = 1;
    if (!sums) {
        VQA_CUDA(c, cudaMemsetAsync(sums_buf, 0, sizeof(unsigned long long) * (size_t)n, c->stream));
        int bps = cdiv((long)h * w, 256 * 16 * 4);
        if (bps < 1) bps = 1;
        VQA_BYTES(c, 1.0 * h * w * n);
        VQA_LAUNCH(c, k_frame_sum, dim3(bps, n), 256, 0, x, (long)h * w, sums_buf);
        sums = sums_buf;
    }
    VQA_BYTES(c, 3.0 * h * w * n);
    VQA_LAUNCH(c, k_u8_to_bf16, dim3(bpf, n), 256, 0, x, h, w, ldw, sums, X);

    alignas(64) CUtensorMap mA0, mA1, mB0, mB1;
    // GEMM 1: T^T[w x h] = Dw[w x w] * X^T ; A rows = w, K = w ; B = X rows = h, K = w, per frame
    if ((rc = make_map(c, s, &mA0, Dw_hi, w, w, ldw, 0, 0, UM_BM))) return rc;
    if ((rc = make_map(c, s, &mA1, Dw_lo, w, w, ldw, 0, 0, UM_BM))) return rc;
    if ((rc = make_map(c, s, &mB0, X, w, h, ldw, n, (size_t)h * ldw, BN1))) return rc;
    mB1 = mB0;
    UmmaOut o1{};
    o1.hi = Tt_hi; o1.lo = Tt_lo; o1.ld = ldh; o1.frame_stride = (size_t)w * ldh;
    {
        const int tiles = cdiv(w, UM_BM) * cdiv(h, BN1) * n;
        const int grid = tiles < c->sm_count ? tiles : c->sm_count;
        VQA_BYTES(c, ((double)h * ldw * 2 + 4.0 * w * h) * n);
        VQA_FLOPS(c, 2.0 * 2.0 * w * w * h * n);
        VQA_LAUNCH(c, (k_dct_umma<BN1, 1>), grid, UM_THREADS, (UmmaCfg<BN1, 1>::SMEM_BYTES), mA0, mA1, mB0, mB1, w, h, w, n, o1);
    }
    // GEMM 2: C[h x w] = Dh[h x h] * T ; A rows = h, K = h ; B = T^T rows = w, K = h, per frame
    UmmaOut o2{};
    o2.C = coef; o2.ldc = w; o2.c_frame_stride = (size_t)h * w; o2.energy = energy;
    VQA_BYTES(c, (4.0 * w * h + 4.0 * w * h) * n);
    VQA_FLOPS(c, 3.0 * 2.0 * h * h * w * n);
#ifdef VQA_AB                                                        // measured slower (profiles/r02_notes.md 14): development build only
    if (gemm2_wide) {
        // 128 x 256 tiles on 32-element K stages (SWIZZLE_64B): 48 KB per stage, 4 stages; 128 flop per staged byte instead of 96
        constexpr int BNW = 256, BKW = 32;
        VQA_CUDA(c, cudaFuncSetAttribute(k_dct_umma<BNW, 2, BKW>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<BNW, 2, BKW>::SMEM_BYTES));
        if ((rc = make_map(c, s, &mA0, Dh_hi, h, h, ldh, 0, 0, UM_BM, BKW))) return rc;
        if ((rc = make_map(c, s, &mA1, Dh_lo, h, h, ldh, 0, 0, UM_BM, BKW))) return rc;
        if ((rc = make_map(c, s, &mB0, Tt_hi, h, w, ldh, n, (size_t)w * ldh, BNW, BKW))) return rc;
        if ((rc = make_map(c, s, &mB1, Tt_lo, h, w, ldh, n, (size_t)w * ldh, BNW, BKW))) return rc;
        const int tiles = cdiv(h, UM_BM) * cdiv(w, BNW) * n;
        const int grid = tiles < c->sm_count ? tiles : c->sm_count;
        VQA_LAUNCH(c, (k_dct_umma<BNW, 2, BKW>), grid, UM_THREADS, (UmmaCfg<BNW, 2, BKW>::SMEM_BYTES), mA0, mA1, mB0, mB1, h, w, h, n, o2);
    } else
#endif
    {
        if ((rc = make_map(c, s, &mA0, Dh_hi, h, h, ldh, 0, 0, UM_BM))) return rc;
        if ((rc = make_map(c, s, &mA1, Dh_lo, h, h, ldh, 0, 0, UM_BM))) return rc;
        if ((rc = make_map(c, s, &mB0, Tt_hi, h, w, ldh, n, (size_t)w * ldh, BN2))) return rc;
        if ((rc = make_map(c, s, &mB1, Tt_lo, h, w, ldh, n, (size_t)w * ldh, BN2))) return rc;
        const int tiles = cdiv(h, UM_BM) * cdiv(w, BN2) * n;
        const int grid = tiles < c->sm_count ? tiles : c->sm_count;
        VQA_LAUNCH(c, (k_dct_umma<BN2, 2>), grid, UM_THREADS, (UmmaCfg<BN2, 2>::SMEM_BYTES), mA0, mA1, mB0, mB1, h, w, h, n, o2);
    }
    VQA_LAUNCH(c, k_dc_fix, cdiv(n, 128), 128, 0, coef, (size_t)h * w, h, w, n, sums, energy);
    return VQA_OK;
}

}  // namespace vqa
