// Framerate variation and the EWM-smoothed mean as device reductions.
//   k_framerate     process_frame_interval_for_parallel (complexity_metrics.py:150-165)
//   k_ewm_partial   np.mean(pd.Series(x).ewm(alpha, adjust=True).mean()) (complexity_metrics.py:114-125,
//                   301-310) as the weighted sum  sum_i c_i x_i,
//                   c_i = (1/T) sum_{t>=i} beta^(t-i) / D_t,  D_t = (1 - beta^(t+1)) / (1 - beta),  beta = 1 - alpha
//                   (SURVEY.md a10).  A shard passes its slice and global offset; partials add up.
#include "vqa_common.cuh"

namespace vqa {

__global__ void k_framerate(const double *__restrict__ ts, int n, double *__restrict__ fps)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += gridDim.x * blockDim.x) {
        double dt = (ts[i + 1] - ts[i]) / 1000.0;
        fps[i] = dt > 0 ? 1.0 / dt : 0.0;
    }
}

// One block.  Coefficients via the backward recurrence S_i = 1/D_i + beta * S_{i+1} evaluated per
// thread with a bounded look-ahead: beta^k underflows double after ~1100 terms for beta <= 0.5 and
// the series is truncated once the term is below 2^-80 of the head (exact to double rounding).
__global__ void k_ewm_partial(const double *__restrict__ x, int n_local, long long offset, long long total,
                              double alpha, double *__restrict__ out)
{
    __shared__ double red[32];
    const double beta = 1.0 - alpha;
    double acc = 0;
    for (int i = threadIdx.x; i < n_local; i += blockDim.x) {
        const long long gi = offset + i;
        double s = 0, bp = 1.0;                       // bp = beta^(t-gi)
        for (long long t = gi; t < total; t++) {
            double Dt = (beta == 1.0) ? (double)(t + 1) : (1.0 - pow(beta, (double)(t + 1))) / (1.0 - beta);
            s += bp / Dt;
            bp *= beta;
            if (bp < 1e-30) break;
        }
        acc += s * x[i];
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int i = 0; i < (blockDim.x >> 5); i++) s += red[i];
        *out = s / (double)total;
    }
}

int run_framerate(vqa_ctx *c, const double *ts_dev, int n, double *fps_dev)
{
    if (n < 2) return VQA_OK;
    VQA_LAUNCH(c, k_framerate, cdiv(n - 1, 256), 256, 0, ts_dev, n, fps_dev);
    return VQA_OK;
}

int run_ewm_partial(vqa_ctx *c, const double *x_dev, int n_local, long long offset, long long total, double alpha,
                    double *out_dev)
{
    VQA_LAUNCH(c, k_ewm_partial, 1, 256, 0, x_dev, n_local, offset, total, alpha, out_dev);
    return VQA_OK;
}

}  // namespace vqa
