// O(T) float64 finishing steps: critic z-score with quantile band (utils/anomaly_detection_utils.py:307-333),
// centred rolling mean (pandas rolling(center=True).mean(), :326-331 / :954-961), z-score + clip (:523-524),
// score combination (:336-362, :554-570) and the per-window statistics / run extraction of find_anomalies
// (:1098-1166).  All results stay on the device; scalars travel through a small block of the workspace.
#include "common.cuh"

namespace hypad {

constexpr int RB = 256;  // threads of the reduction kernels

__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) sh[0] = t;
    }
    __syncthreads();
    t = sh[0];
    __syncthreads();
    return t;
}

template <typename T>
__device__ __forceinline__ double ld(const T* p, int64_t i) { return (double)p[i]; }

// ---------------------------------------------------------------------------------------------------------
// order statistics by radix select on order-preserving 64-bit keys (4 ranks at once)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long dkey(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

constexpr int NQ = 4;  // simultaneous ranks

struct SelectState {  // lives in the workspace
    unsigned long long prefix[NQ];
    long long rank[NQ];
    unsigned int hist[8][NQ][256];
};

__global__ void select_init_kernel(SelectState* st, long long r0, long long r1, long long r2, long long r3) {
    const int t = threadIdx.x + blockIdx.x * blockDim.x;
    unsigned int* h = &st->hist[0][0][0];
    for (int e = t; e < 8 * NQ * 256; e += gridDim.x * blockDim.x) h[e] = 0;
    if (t == 0) {
        st->prefix[0] = st->prefix[1] = st->prefix[2] = st->prefix[3] = 0ull;
        st->rank[0] = r0; st->rank[1] = r1; st->rank[2] = r2; st->rank[3] = r3;
    }
}

// pass p handles bits [56-8p, 64-8p): histogram of the digit among keys whose higher bits equal prefix[q]
__global__ void __launch_bounds__(RB) select_hist_kernel(const double* __restrict__ x, int64_t len, int pass, SelectState* st) {
    __shared__ unsigned int sh[NQ][256];
    for (int e = threadIdx.x; e < NQ * 256; e += blockDim.x) (&sh[0][0])[e] = 0;
    __syncthreads();
    const int shift = 56 - 8 * pass;
    unsigned long long pre[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) pre[q] = st->prefix[q];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const unsigned long long k = dkey(x[i]);
        const unsigned long long hi = pass == 0 ? 0ull : (k >> (shift + 8));
        const unsigned int digit = (unsigned int)((k >> shift) & 255ull);
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            if (hi == pre[q]) atomicAdd(&sh[q][digit], 1u);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < NQ * 256; e += blockDim.x) {
        const unsigned int c = (&sh[0][0])[e];
        if (c) atomicAdd(&st->hist[pass][0][0] + e, c);
    }
}

// one block of NQ warps: locate the digit holding rank[q], extend the prefix
__global__ void __launch_bounds__(NQ * 32) select_pick_kernel(int pass, SelectState* st) {
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int* h = st->hist[pass][q];
    unsigned int c[8];
    unsigned int s = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        c[t] = h[lane * 8 + t];
        s += c[t];
    }
    unsigned int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const long long rank = st->rank[q];
    const long long before = (long long)incl - s;
    const bool mine = rank >= before && rank < (long long)incl;
    if (mine) {
        long long acc = before;
        int digit = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (rank >= acc && rank < acc + c[t]) {
                digit = lane * 8 + t;
                st->rank[q] = rank - acc;
                break;
            }
            acc += c[t];
        }
        st->prefix[q] = (st->prefix[q] << 8) | (unsigned long long)digit;
    }
}

// ---------------------------------------------------------------------------------------------------------
// _compute_critic_score scalars:  s[0]=q25 s[1]=q75 s[2]=mean(all) s[3]=mean(in band) s[4]=std(all, ddof 0)
// ---------------------------------------------------------------------------------------------------------
__global__ void quantile_finish_kernel(const SelectState* st, double g25, double g75, double* s) {
    // numpy _lerp: a + (b-a)*t, and b - (b-a)*(1-t) when t >= 0.5
    const double a0 = dunkey(st->prefix[0]), b0 = dunkey(st->prefix[1]);
    const double a1 = dunkey(st->prefix[2]), b1 = dunkey(st->prefix[3]);
    const double d0 = b0 - a0, d1 = b1 - a1;
    s[0] = g25 >= 0.5 ? b0 - d0 * (1.0 - g25) : a0 + d0 * g25;
    s[1] = g75 >= 0.5 ? b1 - d1 * (1.0 - g75) : a1 + d1 * g75;
}

// partial[b*3 + {0,1,2}] = sum(x), sum(x in band), count(in band)
__global__ void __launch_bounds__(RB) band_partial_kernel(const double* __restrict__ x, int64_t len, const double* s,
                                                          double* __restrict__ partial) {
    __shared__ double sh[32];
    const double lo = s[0], hi = s[1];
    double a = 0.0, b = 0.0, c = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const double v = x[i];
        a += v;
        if (v >= lo && v <= hi) {
            b += v;
            c += 1.0;
        }
    }
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    c = block_sum(c, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x * 3 + 0] = a;
        partial[blockIdx.x * 3 + 1] = b;
        partial[blockIdx.x * 3 + 2] = c;
    }
}
__global__ void __launch_bounds__(RB) band_final_kernel(const double* __restrict__ partial, int nblocks, int64_t len, double* s) {
    __shared__ double sh[32];
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
        a += partial[i * 3];
        b += partial[i * 3 + 1];
        c += partial[i * 3 + 2];
    }
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    c = block_sum(c, sh);
    if (threadIdx.x == 0) {
        s[2] = a / (double)len;
        s[3] = b / c;
    }
}
// generic: partial[b] = sum (x - *center)^2
template <typename T>
__global__ void __launch_bounds__(RB) sqdev_partial_kernel(const T* __restrict__ x, int64_t len, const double* center,
                                                           double* __restrict__ partial) {
    __shared__ double sh[32];
    const double m = *center;
    double a = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const double d = ld(x, i) - m;
        a += d * d;
    }
    a = block_sum(a, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = a;
}
template <typename T>
__global__ void __launch_bounds__(RB) sum_partial_kernel(const T* __restrict__ x, int64_t len, double* __restrict__ partial) {
    __shared__ double sh[32];
    double a = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) a += ld(x, i);
    a = block_sum(a, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = a;
}
// *dst = f(sum(partial)) with f = mean (mode 0) or sqrt(sum / (len - ddof)) (mode 1)
__global__ void __launch_bounds__(RB) scalar_final_kernel(const double* __restrict__ partial, int nblocks, int64_t len, int ddof,
                                                          int mode, double* dst) {
    __shared__ double sh[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) a += partial[i];
    a = block_sum(a, sh);
    if (threadIdx.x == 0) *dst = mode == 0 ? a / (double)len : sqrt(a / (double)(len - ddof));
}

// z = |x - s[3]| / s[4] + 1
__global__ void critic_z_kernel(const double* __restrict__ x, int64_t len, const double* s, double* __restrict__ z) {
    const double mu = s[3], sd = s[4];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) z[i] = fabs((x[i] - mu) / sd) + 1.0;
}

// out = clip((x - s[0]) / s[1], 0) + 1
template <typename T>
__global__ void zscore_clip_kernel(const T* __restrict__ x, int64_t len, const double* s, double* __restrict__ out) {
    const double mu = s[0], sd = s[1];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const double z = (ld(x, i) - mu) / sd;
        out[i] = (z > 0.0 ? z : (z != z ? z : 0.0)) + 1.0;
    }
}

// ---------------------------------------------------------------------------------------------------------
// centred rolling mean through a two-level fp64 prefix sum
// ---------------------------------------------------------------------------------------------------------
constexpr int SCAN_CHUNK = 2048;  // elements per CTA (256 threads x 8)

__global__ void __launch_bounds__(256) scan_local_kernel(const double* __restrict__ x, int64_t len, double* __restrict__ pre,
                                                         double* __restrict__ totals) {
    __shared__ double sh[256];
    const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + threadIdx.x * 8;
    double v[8];
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k] = base + k < len ? x[base + k] : 0.0;
        run += v[k];
        v[k] = run;
    }
    sh[threadIdx.x] = run;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {  // Hillis-Steele inclusive scan of the thread sums
        double t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0.0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    const double off = threadIdx.x ? sh[threadIdx.x - 1] : 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (base + k < len) pre[base + k] = v[k] + off;
    if (threadIdx.x == 255) totals[blockIdx.x] = sh[255];
}
// exclusive scan of the chunk totals in place (one warp, contiguous segments per lane)
__global__ void __launch_bounds__(32) scan_totals_kernel(double* totals, int64_t nchunks) {
    const int lane = threadIdx.x;
    const int64_t per = (nchunks + 31) / 32;
    const int64_t b = lane * per, e = b + per < nchunks ? b + per : nchunks;
    double s = 0.0;
    for (int64_t i = b; i < e; ++i) s += totals[i];
    double incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    double run = incl - s;
    for (int64_t i = b; i < e; ++i) {
        const double t = totals[i];
        totals[i] = run;
        run += t;
    }
}
__global__ void rolling_mean_kernel(const double* __restrict__ pre, const double* __restrict__ offs, int64_t len, int64_t window,
                                    int64_t min_periods, double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t back = window / 2, fwd = (window - 1) / 2;
    const int64_t need = min_periods > 1 ? min_periods : 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        int64_t a = i - back, b = i + fwd + 1;
        if (a < 0) a = 0;
        if (b > len) b = len;
        const int64_t cnt = b - a;
        if (window <= 0 || cnt < need) {
            out[i] = nan("");
            continue;
        }
        const double hi = pre[b - 1] + offs[(b - 1) / SCAN_CHUNK];
        const double lo = a > 0 ? pre[a - 1] + offs[(a - 1) / SCAN_CHUNK] : 0.0;
        out[i] = (hi - lo) / (double)cnt;
    }
}

// ---------------------------------------------------------------------------------------------------------
// combine_scores
// ---------------------------------------------------------------------------------------------------------
template <typename TR>
__global__ void combine_kernel(int mode, const double* __restrict__ c, const TR* __restrict__ r, const float* __restrict__ u,
                               double lambda_rec, int64_t n, double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double cv = c ? c[i] : 0.0;
        const double rv = r ? (double)r[i] : 0.0;
        const double uv = u ? (double)u[i] : 0.0;
        double o;
        switch (mode) {
            case 0: o = cv * rv; break;                                  // mult
            case 1: o = (cv * rv) * uv; break;                           // uncertainty
            case 2: o = 0.2 * cv + 0.8 * rv; break;                      // sum
            case 3: o = cv; break;                                       // critic
            case 4: o = cv * uv; break;                                  // critic_uncertainty
            case 5: o = (0.5 * cv) * uv + (0.5 * rv) * uv; break;        // sum_uncertainty
            case 6: o = rv; break;                                       // rec
            case 7: o = rv * uv; break;                                  // rec_uncertainty
            default: o = (1.0 - lambda_rec) * (cv - 1.0) + lambda_rec * (rv - 1.0); break;  // score_anomalies "sum"
        }
        out[i] = o;
    }
}

// ---------------------------------------------------------------------------------------------------------
// find_anomalies: per analysis window mean / std / threshold, dilated above-threshold runs, max_below
// ---------------------------------------------------------------------------------------------------------
constexpr int TW_THREADS = 1024;
constexpr int TW_MAXPAD = 512;

__device__ __forceinline__ unsigned long long dmax_key(double x) { return dkey(x); }

__global__ void __launch_bounds__(TW_THREADS) threshold_windows_kernel(const double* __restrict__ errors, int64_t len,
                                                                       int64_t window_size, int64_t step, int ddof, int pad,
                                                                       double* __restrict__ stats, double* __restrict__ runs,
                                                                       int32_t* __restrict__ n_runs, int max_runs,
                                                                       unsigned long long* __restrict__ run_max_keys) {
    __shared__ double sh[32];
    __shared__ int s_flags[TW_THREADS + 2 * TW_MAXPAD];
    __shared__ int s_scan[TW_THREADS];
    __shared__ int s_carry[2];  // [0] runs opened so far, [1] dil flag of the element before the tile
    __shared__ unsigned long long s_below;
    const int k = blockIdx.x;
    const int64_t w0 = (int64_t)k * step;
    const int64_t w1 = w0 + window_size < len ? w0 + window_size : len;
    const int64_t n = w1 - w0;
    const double* e = errors + w0;
    const int tid = threadIdx.x;
    // mean, std(ddof), threshold (_fixed_threshold, k = 4)
    double a = 0.0;
    for (int64_t i = tid; i < n; i += TW_THREADS) a += e[i];
    const double mean = block_sum(a, sh) / (double)n;
    a = 0.0;
    for (int64_t i = tid; i < n; i += TW_THREADS) {
        const double d = e[i] - mean;
        a += d * d;
    }
    const double sd = sqrt(block_sum(a, sh) / (double)(n - ddof));
    const double thr = mean + 4.0 * sd;
    double* rk = runs + (size_t)k * max_runs * 3;
    unsigned long long* rmax = run_max_keys + (size_t)k * max_runs;
    for (int i = tid; i < max_runs; i += TW_THREADS) rmax[i] = 0ull;
    if (tid == 0) {
        s_carry[0] = 0;
        s_carry[1] = 0;
        s_below = 0ull;
    }
    __syncthreads();
    bool any_below = false;
    unsigned long long below_key = 0ull;
    for (int64_t t0 = 0; t0 < n; t0 += TW_THREADS) {
        // flags of [t0 - pad, t0 + TW_THREADS + pad)
        for (int f = tid; f < TW_THREADS + 2 * pad; f += TW_THREADS) {
            const int64_t idx = t0 - pad + f;
            s_flags[f] = (idx >= 0 && idx < n && e[idx] > thr) ? 1 : 0;
        }
        __syncthreads();
        const int64_t i = t0 + tid;
        int dil = 0;
        if (i < n) {
            for (int f = tid; f <= tid + 2 * pad; ++f) dil |= s_flags[f];
        }
        const int prev_carry = s_carry[1];
        __syncthreads();
        s_flags[tid] = dil;  // reuse: dil flags of the tile (only [0, TW_THREADS) needed from here on)
        __syncthreads();
        const int prev = tid == 0 ? prev_carry : s_flags[tid - 1];
        const int is_start = (i < n && dil && !prev) ? 1 : 0;
        // inclusive scan of starts
        s_scan[tid] = is_start;
        __syncthreads();
        for (int o = 1; o < TW_THREADS; o <<= 1) {
            const int v = tid >= o ? s_scan[tid - o] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const int base = s_carry[0];
        if (i < n) {
            const double v = e[i];
            if (dil) {
                const int id = base + s_scan[tid] - 1;  // run index of this element
                if (id < max_runs) {
                    if (is_start) rk[id * 3 + 0] = (double)i;
                    const int next = (i + 1 < n) ? ((tid + 1 < TW_THREADS) ? s_flags[tid + 1] : -1) : 0;
                    int nd = next;
                    if (next < 0) {  // first element of the next tile: recompute its dilation
                        nd = 0;
                        for (int64_t j = i + 1 - pad; j <= i + 1 + pad; ++j)
                            if (j >= 0 && j < n && e[j] > thr) nd = 1;
                    }
                    if (!nd) rk[id * 3 + 1] = (double)i;
                    if (v > thr) atomicMax(&rmax[id], dmax_key(v));
                }
            } else {
                any_below = true;
                const unsigned long long kv = dmax_key(v);
                below_key = kv > below_key ? kv : below_key;
            }
        }
        __syncthreads();
        if (tid == TW_THREADS - 1) {
            s_carry[0] = base + s_scan[tid];
            s_carry[1] = dil;
        }
        __syncthreads();
    }
    if (any_below) atomicMax(&s_below, below_key);
    __syncthreads();
    if (tid == 0) {
        stats[k * 4 + 0] = mean;
        stats[k * 4 + 1] = sd;
        stats[k * 4 + 2] = thr;
        stats[k * 4 + 3] = s_below ? dunkey(s_below) : 0.0;  // `above.all()` -> max_below = 0 (:1154-1155)
        n_runs[k] = s_carry[0];
    }
    __syncthreads();
    const int nr = s_carry[0] < max_runs ? s_carry[0] : max_runs;
    for (int r = tid; r < nr; r += TW_THREADS) rk[r * 3 + 2] = dunkey(rmax[r]);
}

static unsigned red_grid(int64_t n) {
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want = ceil_div(n, RB * 8);
    if (want < 1) want = 1;
    const int64_t cap = (int64_t)sms * 4;
    return (unsigned)(want < cap ? want : cap);
}
static unsigned ew_grid(int64_t n) {
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want = ceil_div(n, 256);
    if (want < 1) want = 1;
    const int64_t cap = (int64_t)sms * 16;
    return (unsigned)(want < cap ? want : cap);
}

// workspace layout helpers -------------------------------------------------------------------------------
static inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

static int rolling_mean(hypad_ctx* ctx, const double* x, int64_t len, int64_t window, int64_t min_periods, double* out,
                        char* ws, cudaStream_t stream) {
    // ws: pre[len] | totals[nchunks]
    const int64_t nchunks = ceil_div(len, SCAN_CHUNK);
    double* pre = (double*)ws;
    double* totals = (double*)(ws + align256((size_t)len * 8));
    scan_local_kernel<<<(unsigned)nchunks, 256, 0, stream>>>(x, len, pre, totals);
    HYPAD_LAUNCH_CHECK();
    scan_totals_kernel<<<1, 32, 0, stream>>>(totals, nchunks);
    HYPAD_LAUNCH_CHECK();
    rolling_mean_kernel<<<ew_grid(len), 256, 0, stream>>>(pre, totals, len, window, min_periods, out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}
static size_t rolling_ws_bytes(int64_t len) { return align256((size_t)len * 8) + align256((size_t)ceil_div(len, SCAN_CHUNK) * 8); }

}  // namespace hypad

using namespace hypad;

extern "C" {

int hypad_rolling_mean_centered(hypad_ctx* ctx, const double* x, int64_t len, int64_t window, int64_t min_periods, double* out,
                                void* stream) {
    HYPAD_REQUIRE(ctx && x && out, "hypad_rolling_mean_centered: NULL argument");
    HYPAD_REQUIRE(len >= 0, "hypad_rolling_mean_centered: len < 0");
    if (len == 0) return HYPAD_OK;
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_workspace(ctx, rolling_ws_bytes(len));
    if (rc != HYPAD_OK) return rc;
    return rolling_mean(ctx, x, len, window, min_periods, out, (char*)ctx->workspace, (cudaStream_t)stream);
}

int hypad_critic_zscore_smooth(hypad_ctx* ctx, const double* kmax, int64_t len, int64_t smooth_window, double* out,
                               void* stream_) {
    HYPAD_REQUIRE(ctx && kmax && out, "hypad_critic_zscore_smooth: NULL argument");
    HYPAD_REQUIRE(len >= 1, "hypad_critic_zscore_smooth: len < 1");
    cudaStream_t stream = (cudaStream_t)stream_;
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    const unsigned rg = red_grid(len);
    // workspace: scalars[8] | SelectState | partial[3*rg] | z[len] | rolling scratch
    const size_t o_state = 256, o_part = o_state + align256(sizeof(SelectState)), o_z = o_part + align256((size_t)rg * 3 * 8);
    const size_t o_roll = o_z + align256((size_t)len * 8);
    int rc = ensure_workspace(ctx, o_roll + rolling_ws_bytes(len));
    if (rc != HYPAD_OK) return rc;
    char* ws = (char*)ctx->workspace;
    double* s = (double*)ws;
    SelectState* st = (SelectState*)(ws + o_state);
    double* partial = (double*)(ws + o_part);
    double* z = (double*)(ws + o_z);
    // np.quantile(method='linear'): virtual index q*(n-1); neighbours floor / floor+1 (clipped)
    const double v25 = 0.25 * (double)(len - 1), v75 = 0.75 * (double)(len - 1);
    const long long f25 = (long long)v25, f75 = (long long)v75;
    const long long c25 = f25 + 1 < len ? f25 + 1 : len - 1, c75 = f75 + 1 < len ? f75 + 1 : len - 1;
    select_init_kernel<<<8, 256, 0, stream>>>(st, f25, c25, f75, c75);
    HYPAD_LAUNCH_CHECK();
    for (int p = 0; p < 8; ++p) {
        select_hist_kernel<<<rg, RB, 0, stream>>>(kmax, len, p, st);
        HYPAD_LAUNCH_CHECK();
        select_pick_kernel<<<1, NQ * 32, 0, stream>>>(p, st);
        HYPAD_LAUNCH_CHECK();
    }
    quantile_finish_kernel<<<1, 1, 0, stream>>>(st, v25 - (double)f25, v75 - (double)f75, s);
    HYPAD_LAUNCH_CHECK();
    band_partial_kernel<<<rg, RB, 0, stream>>>(kmax, len, s, partial);
    HYPAD_LAUNCH_CHECK();
    band_final_kernel<<<1, RB, 0, stream>>>(partial, (int)rg, len, s);
    HYPAD_LAUNCH_CHECK();
    sqdev_partial_kernel<double><<<rg, RB, 0, stream>>>(kmax, len, s + 2, partial);
    HYPAD_LAUNCH_CHECK();
    scalar_final_kernel<<<1, RB, 0, stream>>>(partial, (int)rg, len, 0, 1, s + 4);
    HYPAD_LAUNCH_CHECK();
    critic_z_kernel<<<ew_grid(len), 256, 0, stream>>>(kmax, len, s, z);
    HYPAD_LAUNCH_CHECK();
    return rolling_mean(ctx, z, len, smooth_window, smooth_window / 2, out, ws + o_roll, stream);
}

int hypad_zscore_clip(hypad_ctx* ctx, const void* x, int x_is_f32, int64_t len, double* out, void* stream_) {
    HYPAD_REQUIRE(ctx && x && out, "hypad_zscore_clip: NULL argument");
    HYPAD_REQUIRE(len >= 1, "hypad_zscore_clip: len < 1");
    cudaStream_t stream = (cudaStream_t)stream_;
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    const unsigned rg = red_grid(len);
    int rc = ensure_workspace(ctx, 256 + align256((size_t)rg * 8));
    if (rc != HYPAD_OK) return rc;
    double* s = (double*)ctx->workspace;
    double* partial = (double*)((char*)ctx->workspace + 256);
    if (x_is_f32) sum_partial_kernel<float><<<rg, RB, 0, stream>>>((const float*)x, len, partial);
    else sum_partial_kernel<double><<<rg, RB, 0, stream>>>((const double*)x, len, partial);
    HYPAD_LAUNCH_CHECK();
    scalar_final_kernel<<<1, RB, 0, stream>>>(partial, (int)rg, len, 0, 0, s);
    HYPAD_LAUNCH_CHECK();
    if (x_is_f32) sqdev_partial_kernel<float><<<rg, RB, 0, stream>>>((const float*)x, len, s, partial);
    else sqdev_partial_kernel<double><<<rg, RB, 0, stream>>>((const double*)x, len, s, partial);
    HYPAD_LAUNCH_CHECK();
    scalar_final_kernel<<<1, RB, 0, stream>>>(partial, (int)rg, len, 0, 1, s + 1);
    HYPAD_LAUNCH_CHECK();
    if (x_is_f32) zscore_clip_kernel<float><<<ew_grid(len), 256, 0, stream>>>((const float*)x, len, s, out);
    else zscore_clip_kernel<double><<<ew_grid(len), 256, 0, stream>>>((const double*)x, len, s, out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_combine_scores(int mode, const double* critic_scores, const void* rec, int rec_is_f32, const float* unorm,
                         double lambda_rec, int64_t n, double* out, void* stream_) {
    HYPAD_REQUIRE(out != nullptr, "hypad_combine_scores: out is NULL");
    HYPAD_REQUIRE(mode >= 0 && mode <= 8, "hypad_combine_scores: unknown mode %d", mode);
    const bool need_c = mode == 0 || mode == 1 || mode == 2 || mode == 3 || mode == 4 || mode == 5 || mode == 8;
    const bool need_r = mode == 0 || mode == 1 || mode == 2 || mode == 5 || mode == 6 || mode == 7 || mode == 8;
    const bool need_u = mode == 1 || mode == 4 || mode == 5 || mode == 7;
    HYPAD_REQUIRE(!need_c || critic_scores, "hypad_combine_scores: mode %d needs critic_scores", mode);
    HYPAD_REQUIRE(!need_r || rec, "hypad_combine_scores: mode %d needs rec", mode);
    HYPAD_REQUIRE(!need_u || unorm, "hypad_combine_scores: mode %d needs unorm", mode);
    if (n <= 0) return n == 0 ? HYPAD_OK : HYPAD_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (rec_is_f32) combine_kernel<float><<<ew_grid(n), 256, 0, stream>>>(mode, critic_scores, (const float*)rec, unorm, lambda_rec, n, out);
    else combine_kernel<double><<<ew_grid(n), 256, 0, stream>>>(mode, critic_scores, (const double*)rec, unorm, lambda_rec, n, out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_threshold_windows(hypad_ctx* ctx, const double* errors, int64_t len, int64_t window_size, int64_t step,
                            int64_t n_analysis, int ddof, int anomaly_padding, double* stats, double* runs,
                            int32_t* n_runs, int max_runs, void* stream_) {
    HYPAD_REQUIRE(ctx && errors && stats && runs && n_runs, "hypad_threshold_windows: NULL argument");
    HYPAD_REQUIRE(len >= 1 && window_size >= 1 && step >= 1 && n_analysis >= 1 && max_runs >= 1, "hypad_threshold_windows: bad shape");
    HYPAD_REQUIRE((n_analysis - 1) * step < len, "hypad_threshold_windows: last window starts beyond the data");
    HYPAD_REQUIRE(anomaly_padding >= 0 && anomaly_padding <= TW_MAXPAD, "hypad_threshold_windows: padding %d outside 0..%d",
                  anomaly_padding, TW_MAXPAD);
    HYPAD_REQUIRE(ddof == 0 || ddof == 1, "hypad_threshold_windows: ddof must be 0 or 1");
    cudaStream_t stream = (cudaStream_t)stream_;
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_workspace(ctx, (size_t)n_analysis * max_runs * 8);
    if (rc != HYPAD_OK) return rc;
    threshold_windows_kernel<<<(unsigned)n_analysis, TW_THREADS, 0, stream>>>(errors, len, window_size, step, ddof, anomaly_padding,
                                                                              stats, runs, n_runs, max_runs,
                                                                              (unsigned long long*)ctx->workspace);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

}  // extern "C"
