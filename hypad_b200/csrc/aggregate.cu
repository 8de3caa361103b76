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
// Timestep i takes the median of y_hat[i-j][j], j in [max(0,i-N+1), min(i,S-1)] (n <= 128 values): np.median semantics for
// fp32 -- odd n the middle value, even n the fp32 mean (a+b)/2 of the two middles, NaN if any value is NaN.
// An anti-diagonal is the worst possible access pattern for the row-major reconstruction (lanes S-1 floats apart, one sector
// per lane), and every element belongs to exactly one of them.  So a CTA takes MED_TILE consecutive timesteps, copies the
// MED_TILE + S - 1 reconstruction rows they touch into shared memory with coalesced 16-byte loads (neighbouring CTAs re-read
// the S - 1 rows they share from L2), and its warps cut the diagonals from there: lane k reads row (i - lo - k), column lo + k,
// S - 1 words apart -- odd for an even pitch, no bank conflict.  The median comes from a bitonic sort of the <= 128 values
// across the warp (4 per lane as order-preserving integer keys, padding sorts last): 28 compare-exchange stages instead of
// the n^2 = 10^4 comparisons of ranking by counting.
constexpr int MED_WARPS = 8;
constexpr int MED_TILE = 128;

__device__ __forceinline__ unsigned int med_key(float v) {
    if (v != v) return 0xffffffffu;  // NaN sorts last (and makes the result NaN)
    const unsigned int b = __float_as_uint(v);
    return (b >> 31) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float med_unkey(unsigned int k) { return __uint_as_float((k >> 31) ? (k & 0x7fffffffu) : ~k); }

__global__ void __launch_bounds__(MED_WARPS * 32) median_overlap_kernel(const float* __restrict__ y_hat, int64_t N, int S, int pitch,
                                                                       float* __restrict__ pred) {
    extern __shared__ __align__(16) float srows[];  // (MED_TILE + S - 1) rows x pitch
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t T = N + S - 1;
    const int64_t ntiles = (T + MED_TILE - 1) / MED_TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t t0 = tile * MED_TILE;
        const int64_t w_lo = t0 - (S - 1) > 0 ? t0 - (S - 1) : 0;
        const int64_t w_hi = t0 + MED_TILE - 1 < N - 1 ? t0 + MED_TILE - 1 : N - 1;  // inclusive
        const int rows = (int)(w_hi - w_lo + 1);
        __syncthreads();  // the previous tile's diagonals have been read
        if (rows > 0) {
            const float* src = y_hat + w_lo * (int64_t)S;
            const int total = rows * S;
            if (pitch == S && (S & 3) == 0) {
                const float4* s4 = reinterpret_cast<const float4*>(src);  // w_lo * S floats: a multiple of 4, 16-byte aligned
                float4* d4 = reinterpret_cast<float4*>(srows);
                for (int e = threadIdx.x; e < total / 4; e += MED_WARPS * 32) d4[e] = __ldg(s4 + e);
            } else {
                for (int e = threadIdx.x; e < total; e += MED_WARPS * 32) {
                    const int w = e / S;
                    srows[w * pitch + (e - w * S)] = __ldg(src + e);
                }
            }
        }
        __syncthreads();
        for (int tt = warp; tt < MED_TILE; tt += MED_WARPS) {
            const int64_t i = t0 + tt;
            if (i >= T) break;
            const int lo = (int)(i - N + 1 > 0 ? i - N + 1 : 0);
            const int hi = (int)(i < S - 1 ? i : S - 1);
            const int n = hi - lo + 1;
            unsigned int key[4];
            bool has_nan = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = lane + 32 * q;
                key[q] = 0xffffffffu;
                if (k < n) {
                    const int j = lo + k;
                    const float v = srows[(int)(i - j - w_lo) * pitch + j];
                    has_nan |= v != v;
                    key[q] = med_key(v);
                }
            }
            has_nan = __any_sync(0xffffffffu, has_nan);
            // bitonic sort, ascending, of the 128 keys at positions p = lane + 32 q
#pragma unroll
            for (int k = 2; k <= 128; k <<= 1) {
#pragma unroll
                for (int j = k >> 1; j > 0; j >>= 1) {
                    if (j >= 32) {  // partner in the same lane: registers q and q ^ (j / 32)
                        const int dq = j >> 5;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (!(q & dq)) {
                                const bool up = ((q * 32) & k) == 0;  // ascending block (lane bits are below 32 <= j < k)
                                const unsigned int a = key[q], b = key[q | dq];
                                const unsigned int mn = a < b ? a : b, mx = a < b ? b : a;
                                key[q] = up ? mn : mx;
                                key[q | dq] = up ? mx : mn;
                            }
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int p = lane + 32 * q;
                            const unsigned int other = __shfl_xor_sync(0xffffffffu, key[q], j);
                            const bool up = (p & k) == 0, lower = (lane & j) == 0;
                            const unsigned int mn = key[q] < other ? key[q] : other, mx = key[q] < other ? other : key[q];
                            key[q] = (up == lower) ? mn : mx;
                        }
                    }
                }
            }
            const int r_lo = (n - 1) >> 1, r_hi = n >> 1;
            unsigned int ka = 0, kb = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if ((r_lo >> 5) == q) ka = key[q];
                if ((r_hi >> 5) == q) kb = key[q];
            }
            ka = __shfl_sync(0xffffffffu, ka, r_lo & 31);
            kb = __shfl_sync(0xffffffffu, kb, r_hi & 31);
            if (lane == 0) {
                const float a = med_unkey(ka), b = med_unkey(kb);
                float m = (n & 1) ? a : __fmul_rn(__fadd_rn(a, b), 0.5f);
                if (has_nan) m = __uint_as_float(0x7fc00000u);
                pred[i] = m;
            }
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
    const int pitch = S + (S & 1);  // even pitch: the diagonal's lane stride pitch - 1 is odd (no shared-memory bank conflict)
    const size_t smem = (size_t)(MED_TILE + S - 1) * pitch * sizeof(float);
    static bool configured = false;  // idempotent, cheap: racing threads at worst set the attribute twice
    if (!configured) {
        HYPAD_CUDA_TRY(cudaFuncSetAttribute(median_overlap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (MED_TILE + 127) * 128 * 4));
        configured = true;
    }
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t ntiles = ceil_div(n + S - 1, MED_TILE);
    const int64_t cap = (int64_t)sms * 2;  // two resident CTAs per SM (91 KB of rows each at S = 100)
    const unsigned grid = (unsigned)(ntiles < cap ? ntiles : cap);
    median_overlap_kernel<<<grid, MED_WARPS * 32, smem, (cudaStream_t)stream>>>(y_hat, n, S, pitch, pred);
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
