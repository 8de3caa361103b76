// Streaming (HBM-bound) kernels around the network: window gather, median overlap aggregation,
// ground-truth unrolling, and the stand-alone row-wise Poincare distance / norm.
#include "common.cuh"

namespace hypad {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- utils/dataloader.py:139-222: out[w][j] = X[w + j] -------------------------------------------------
template <typename T>
__global__ void window_gather_kernel(const double* __restrict__ X, int64_t n_windows, int S, T* __restrict__ out) {
    const int64_t total = n_windows * S;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t w = e / S;
        const int j = (int)(e - w * S);
        out[e] = (T)X[w + j];
    }
}

// ---- utils/anomaly_detection_utils.py:908-910 ----------------------------------------------------------
template <typename T>
__global__ void true_from_signal_kernel(const T* __restrict__ x, int64_t n, int64_t stride_, int S, double* __restrict__ out) {
    const int64_t total = n + S - 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t src = i < n ? i * stride_ : (n - 1) * stride_ + (i - (n - 1));
        out[i] = (double)x[src];
    }
}

// ---- utils/anomaly_detection_utils.py:918-923: median over the anti-diagonal ---------------------------
// One warp per timestep i: values y_hat[i-j][j], j in [max(0,i-N+1), min(i,S-1)] (n <= 128).  Ranks by counting
// (stable on ties), np.median semantics for fp32: odd n -> middle, even n -> fp32 (a+b)/2 of the two middles.
// A CTA handles MED_WARPS consecutive timesteps so that the 128-byte lines it touches are shared through L1.
constexpr int MED_WARPS = 8;

__global__ void __launch_bounds__(MED_WARPS * 32) median_overlap_kernel(const float* __restrict__ y_hat, int64_t N, int S,
                                                                       float* __restrict__ pred) {
    __shared__ float sV[MED_WARPS][128];
    __shared__ float sMid[MED_WARPS][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* V = sV[warp];
    const int64_t T = N + S - 1;
    const int64_t nblk = (T + MED_WARPS - 1) / MED_WARPS;
    for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int64_t i = blk * MED_WARPS + warp;
        if (i < T) {
            const int lo = (int)(i - N + 1 > 0 ? i - N + 1 : 0);
            const int hi = (int)(i < S - 1 ? i : S - 1);
            const int n = hi - lo + 1;
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = lane + 32 * q;
                v[q] = 0.f;
                if (k < n) {
                    const int j = lo + k;
                    v[q] = y_hat[(i - j) * (int64_t)S + j];
                    V[k] = v[q];
                }
            }
            __syncwarp();
            const int r_lo = (n - 1) >> 1, r_hi = n >> 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = lane + 32 * q;
                if (k < n) {
                    int rank = 0;
                    const float mine = v[q];
                    for (int t = 0; t < n; ++t) {
                        const float o = V[t];
                        rank += (o < mine) || (o == mine && t < k);
                    }
                    if (rank == r_lo) sMid[warp][0] = mine;
                    if (rank == r_hi) sMid[warp][1] = mine;
                }
            }
            __syncwarp();
            if (lane == 0) {
                const float a = sMid[warp][0], b = sMid[warp][1];
                pred[i] = (n & 1) ? a : __fmul_rn(__fadd_rn(a, b), 0.5f);
            }
            __syncwarp();
        }
    }
}

// ---- utils/anomaly_detection_utils.py:58-66 and np.linalg.norm(axis=1) (:342) ---------------------------
__global__ void poincare_rowdist_kernel(const float* __restrict__ recons, const float* __restrict__ truth, int64_t n, int S,
                                        float* __restrict__ rec, float* __restrict__ unorm) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
        const float* h = recons + row * S;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int c = lane; c < S; c += 32) {
            const float hv = h[c];
            s2 += (double)__fmul_rn(hv, hv);
            if (truth) {
                const float hx = truth[row * S + c];
                const float d = __fsub_rn(hx, hv);
                s0 += (double)__fmul_rn(d, d);
                s1 += (double)__fmul_rn(hx, hx);
            }
        }
        s0 = warp_sum_d(s0);
        s1 = warp_sum_d(s1);
        s2 = warp_sum_d(s2);
        if (lane == 0) {
            const float sqdist = (float)s0, squnorm = (float)s1, sqvnorm = (float)s2;
            if (rec) {
                const float t = __fdiv_rn(__fmul_rn(2.0f, sqdist), __fmul_rn(__fsub_rn(1.0f, squnorm), __fsub_rn(1.0f, sqvnorm)));
                const float xt = __fadd_rn(__fadd_rn(1.0f, t), 1e-7f);
                rec[row] = (float)acosh((double)xt);
            }
            if (unorm) unorm[row] = sqrtf(sqvnorm);
        }
    }
}

// ---- np.linalg.norm(true - recons, axis=1), utils/anomaly_detection_utils.py:157 (Euclidean multivariate) -----------
// true rows are float64 (the dataloader's samples) or float32, the reconstruction float32: the difference and the norm are
// float64 like numpy's (whose pairwise summation order differs in the last bits only).
template <typename T>
__global__ void rowdiff_norm_kernel(const T* __restrict__ truth, const float* __restrict__ recons, int64_t n, int S, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
        double s = 0.0;
        for (int c = lane; c < S; c += 32) {
            const double d = (double)truth[row * S + c] - (double)recons[row * S + c];
            s += d * d;
        }
        s = warp_sum_d(s);
        if (lane == 0) out[row] = sqrt(s);
    }
}

static unsigned grid_for(int64_t items, int per_block, int dev_mult = 16) {
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want = ceil_div(items, per_block);
    int64_t cap = (int64_t)sms * dev_mult;
    if (want < 1) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

}  // namespace hypad

using namespace hypad;

extern "C" {

int hypad_window_gather(const double* X, int64_t n_windows, int S, void* out, int out_is_f64, void* stream) {
    HYPAD_REQUIRE(X && out, "hypad_window_gather: NULL argument");
    HYPAD_REQUIRE(S >= 1 && n_windows >= 0, "hypad_window_gather: bad shape");
    if (n_windows == 0) return HYPAD_OK;
    const unsigned grid = grid_for(n_windows * S, 256 * 4);
    if (out_is_f64) window_gather_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(X, n_windows, S, (double*)out);
    else window_gather_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(X, n_windows, S, (float*)out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_true_from_signal(const void* x, int x_is_f64, int64_t n, int64_t row_stride, int S, double* out, void* stream) {
    HYPAD_REQUIRE(x && out, "hypad_true_from_signal: NULL argument");
    HYPAD_REQUIRE(n >= 1 && S >= 1 && row_stride >= 1, "hypad_true_from_signal: bad shape");
    const unsigned grid = grid_for(n + S - 1, 256);
    if (x_is_f64) true_from_signal_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)x, n, row_stride, S, out);
    else true_from_signal_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, n, row_stride, S, out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_median_overlap(const float* y_hat, int64_t n, int S, float* pred, void* stream) {
    HYPAD_REQUIRE(y_hat && pred, "hypad_median_overlap: NULL argument");
    HYPAD_REQUIRE(n >= 1 && S >= 1 && S <= 128, "hypad_median_overlap: S=%d outside 1..128 or n<1", S);
    const unsigned grid = grid_for(n + S - 1, MED_WARPS, 8);
    median_overlap_kernel<<<grid, MED_WARPS * 32, 0, (cudaStream_t)stream>>>(y_hat, n, S, pred);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_poincare_rowdist(const float* recons, const float* truth, int64_t n, int S, float* out, void* stream) {
    HYPAD_REQUIRE(recons && truth && out, "hypad_poincare_rowdist: NULL argument");
    HYPAD_REQUIRE(n >= 0 && S >= 1, "hypad_poincare_rowdist: bad shape");
    if (n == 0) return HYPAD_OK;
    poincare_rowdist_kernel<<<grid_for(n, 8), 256, 0, (cudaStream_t)stream>>>(recons, truth, n, S, out, nullptr);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_rownorm(const float* x, int64_t n, int S, float* out, void* stream) {
    HYPAD_REQUIRE(x && out, "hypad_rownorm: NULL argument");
    HYPAD_REQUIRE(n >= 0 && S >= 1, "hypad_rownorm: bad shape");
    if (n == 0) return HYPAD_OK;
    poincare_rowdist_kernel<<<grid_for(n, 8), 256, 0, (cudaStream_t)stream>>>(x, nullptr, n, S, nullptr, out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_rowdiff_norm(const void* truth, int truth_is_f64, const float* recons, int64_t n, int S, double* out, void* stream) {
    HYPAD_REQUIRE(truth && recons && out, "hypad_rowdiff_norm: NULL argument");
    HYPAD_REQUIRE(n >= 0 && S >= 1, "hypad_rowdiff_norm: bad shape");
    if (n == 0) return HYPAD_OK;
    if (truth_is_f64) rowdiff_norm_kernel<double><<<grid_for(n, 8), 256, 0, (cudaStream_t)stream>>>((const double*)truth, recons, n, S, out);
    else rowdiff_norm_kernel<float><<<grid_for(n, 8), 256, 0, (cudaStream_t)stream>>>((const float*)truth, recons, n, S, out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

}  // extern "C"
