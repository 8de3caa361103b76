// Global statistics of the finish in stages a caller can put an exchange between: every stage works on THIS GPU's slice of a
// per-position array and leaves a small record (a histogram, a few partial sums) that is combined with the other ranks'
// records -- gathered by the caller, torch.distributed / NCCL over NVLink -- by the next stage.  With one rank the stages run
// back to back on the local record.  hypad_critic_zscore_smooth (finish.cu) is exactly that chain.
//
//   _compute_critic_score (utils/anomaly_detection_utils.py:307-333):
//     q25, q75 = np.quantile(critics, .25 / .75)          -> radix select on order-preserving keys, 11 bits per pass
//     mean of the values inside [q25, q75], std of all     -> (hi, lo) partial sums, order-independent (dd.cuh)
//   zscore + clip of the multivariate reconstruction error (:177-178, :523-524): mean and std of all rows, same sums.
#include "common.cuh"
#include "dd.cuh"
#include "finish_common.cuh"

namespace hypad {

constexpr int RB = 256;
constexpr int NQ = 4;            // simultaneous order statistics: floor / ceil neighbours of the two quantiles
constexpr int SEL_BINS = 2048;   // 11-bit digits

struct FinState {
    unsigned long long prefix[NQ];
    long long rank[NQ];
    int rep[NQ];        // first query with the same prefix: its histogram serves this one too
    int key_bits;       // 32: the values are fp32-representable (KDE selections of fp32 critics); 64 otherwise
    int passes;
    double g[2];        // interpolation fractions of the two quantiles
    double s[8];        // 0 q25, 1 q75, 2 mean(all), 3 mean(in band), 4 std(all, ddof 0); 5 mean, 6 std of the z-score input
    unsigned int ticket;
    unsigned int pad;
};

__device__ __forceinline__ unsigned long long dkey64(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey64(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ unsigned long long dkey32(double x) {
    unsigned int b = __float_as_uint((float)x);
    return (unsigned long long)((b >> 31) ? ~b : (b | 0x80000000u));
}
__device__ __forceinline__ double dunkey32(unsigned long long k) {
    const unsigned int b = (unsigned int)k;
    return (double)__uint_as_float((b >> 31) ? (b & 0x7fffffffu) : ~b);
}

// digits from the top: 32-bit keys 11 | 11 | 10, 64-bit keys 11 | 11 | 11 | 11 | 10 | 10
__host__ __device__ inline void sel_digit(int key_bits, int pass, int* shift, int* width) {
    if (key_bits == 32) {
        const int sh[3] = {21, 10, 0}, w[3] = {11, 11, 10};
        *shift = sh[pass];
        *width = w[pass];
    } else {
        const int sh[6] = {53, 42, 31, 20, 10, 0}, w[6] = {11, 11, 11, 11, 10, 10};
        *shift = sh[pass];
        *width = w[pass];
    }
}

__device__ __forceinline__ void sel_begin_body(FinState* st, long long n_total, int key_bits);
__global__ void sel_begin_kernel(FinState* st, long long n_total, int key_bits) { sel_begin_body(st, n_total, key_bits); }
__device__ __forceinline__ void sel_begin_body(FinState* st, long long n_total, int key_bits) {
    // np.quantile(method='linear'): virtual index q (n-1); neighbours floor and floor+1 (clipped)
    const double v25 = 0.25 * (double)(n_total - 1), v75 = 0.75 * (double)(n_total - 1);
    const long long f25 = (long long)v25, f75 = (long long)v75;
    st->rank[0] = f25;
    st->rank[1] = f25 + 1 < n_total ? f25 + 1 : n_total - 1;
    st->rank[2] = f75;
    st->rank[3] = f75 + 1 < n_total ? f75 + 1 : n_total - 1;
    st->g[0] = v25 - (double)f25;
    st->g[1] = v75 - (double)f75;
    for (int q = 0; q < NQ; ++q) {
        st->prefix[q] = 0ull;
        st->rep[q] = 0;
    }
    st->key_bits = key_bits;
    st->passes = key_bits == 32 ? 3 : 6;
    st->ticket = 0u;
}

// hist[q][digit] += 1 for every local element whose higher key bits equal prefix[q] (only for q == rep[q]).
// Values cluster (critic outputs live in a narrow range), so a thread merges consecutive equal (digit, query set) pairs
// before touching the shared-memory histogram.
__global__ void __launch_bounds__(RB) sel_hist_kernel(const double* __restrict__ x, int64_t len, int pass, const FinState* st,
                                                      unsigned int* __restrict__ hist) {
    __shared__ unsigned int sh[NQ][SEL_BINS];
    for (int e = threadIdx.x; e < NQ * SEL_BINS; e += RB) (&sh[0][0])[e] = 0u;
    __syncthreads();
    const int key_bits = st->key_bits;
    int shift, width;
    sel_digit(key_bits, pass, &shift, &width);
    const int hi_shift = shift + width;
    unsigned long long pre[NQ];
    bool own[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        pre[q] = st->prefix[q];
        own[q] = st->rep[q] == q;
    }
    const unsigned int dmask = (1u << width) - 1u;
    unsigned int last = 0xffffffffu, cnt = 0u;
    const int64_t stride = (int64_t)gridDim.x * RB;
    for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < len; i += stride) {
        const unsigned long long k = key_bits == 32 ? dkey32(x[i]) : dkey64(x[i]);
        const unsigned long long hi = hi_shift >= key_bits ? 0ull : (k >> hi_shift);
        unsigned int cur = (unsigned int)(k >> shift) & dmask;
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            if (own[q] && hi == pre[q]) cur |= 1u << (12 + q);
        if (cur == last) {
            ++cnt;
        } else {
            if (cnt) {
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    if (last & (1u << (12 + q))) atomicAdd(&sh[q][last & 2047u], cnt);
            }
            last = cur;
            cnt = 1u;
        }
    }
    if (cnt) {
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            if (last & (1u << (12 + q))) atomicAdd(&sh[q][last & 2047u], cnt);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < NQ * SEL_BINS; e += RB) {
        const unsigned int c = (&sh[0][0])[e];
        if (c) atomicAdd(hist + e, c);
    }
}

// One CTA, 256 threads per order statistic: sum the ranks' histograms (8 consecutive bins per thread), locate the digit holding
// rank[q] with a block-wide prefix sum, extend the prefix; after the last pass turn the four order statistics into the two
// quantiles (numpy's _lerp).
// (NQ * 256 threads take part; st and hists may live in shared or global memory; wsum: shared scratch)
__device__ __forceinline__ void sel_pick_body(const unsigned int* __restrict__ hists, int world, int pass, FinState* st, long long (*wsum)[8],
                                              int tid) {
    const int q = tid >> 8, t = tid & 255, lane = tid & 31, w = t >> 5;
    int shift, width;
    sel_digit(st->key_bits, pass, &shift, &width);
    const int bins = 1 << width;  // 1024 or 2048
    const int src = st->rep[q];
    const long long rank = st->rank[q];
    const unsigned long long prefix = st->prefix[q];
    long long c[8];
    long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) c[k] = 0;
    if (t * 8 < bins) {
        for (int r = 0; r < world; ++r) {
            const uint4* p = reinterpret_cast<const uint4*>(hists + ((size_t)r * NQ + src) * SEL_BINS + t * 8);
            const uint4 a = p[0], b = p[1];
            c[0] += a.x; c[1] += a.y; c[2] += a.z; c[3] += a.w; c[4] += b.x; c[5] += b.y; c[6] += b.z; c[7] += b.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s += c[k];
    }
    long long incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[q][w] = incl;
    __syncthreads();  // also: every thread has read the state before any thread updates it
    long long before = incl - s;
    for (int k = 0; k < w; ++k) before += wsum[q][k];
    if (rank >= before && rank < before + s) {
        long long acc = before;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (rank >= acc && rank < acc + c[k]) {
                st->rank[q] = rank - acc;
                st->prefix[q] = (prefix << width) | (unsigned long long)(t * 8 + k);
            }
            acc += c[k];
        }
    }
    __syncthreads();
    if (tid == 0) {
        for (int a = 0; a < NQ; ++a) {
            int rep = a;
            for (int b = a - 1; b >= 0; --b)
                if (st->prefix[b] == st->prefix[a]) rep = b;
            st->rep[a] = rep;
        }
        if (pass == st->passes - 1) {
            double v[NQ];
            for (int a = 0; a < NQ; ++a) v[a] = st->key_bits == 32 ? dunkey32(st->prefix[a]) : dunkey64(st->prefix[a]);
            // numpy _lerp: a + (b-a)*t, and b - (b-a)*(1-t) when t >= 0.5
            const double d0 = v[1] - v[0], d1 = v[3] - v[2], g25 = st->g[0], g75 = st->g[1];
            st->s[0] = g25 >= 0.5 ? v[1] - d0 * (1.0 - g25) : v[0] + d0 * g25;
            st->s[1] = g75 >= 0.5 ? v[3] - d1 * (1.0 - g75) : v[2] + d1 * g75;
        }
    }
}

__global__ void __launch_bounds__(NQ * 256) sel_pick_kernel(const unsigned int* __restrict__ hists, int world, int pass, FinState* st) {
    __shared__ long long wsum[NQ][8];
    sel_pick_body(hists, world, pass, st, wsum, threadIdx.x);
}

// Local partial sums: out[0..1] sum x, [2..3] sum x^2, [4..5] sum of the x inside [lo, hi], [6] their count -- (hi, lo) pairs.
// band: lo / hi come from st->s[0..1]; without it only the first two sums are taken (zscore moments).
// The last CTA to finish folds the per-CTA records in a fixed pattern (the result does not depend on scheduling).
template <typename T>
__global__ void __launch_bounds__(RB) moments_partial_kernel(const T* __restrict__ x, int64_t len, int band, FinState* st,
                                                             double* __restrict__ scratch, double* __restrict__ out) {
    __shared__ double sh[64];
    __shared__ bool last_cta;
    const double lo = band ? st->s[0] : 0.0, hi = band ? st->s[1] : 0.0;
    dd s1 = dd_make(0.0), s2 = dd_make(0.0), sb = dd_make(0.0);
    double nb = 0.0;
    const int64_t stride = (int64_t)gridDim.x * RB;
    for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < len; i += stride) {
        const double v = (double)x[i];
        s1 = dd_add(s1, v);
        s2 = dd_add(s2, dd_prod(v, v));
        if (band && v >= lo && v <= hi) {
            sb = dd_add(sb, v);
            nb += 1.0;
        }
    }
    s1 = dd_block_sum(s1, sh);
    s2 = dd_block_sum(s2, sh);
    sb = dd_block_sum(sb, sh);
    const dd nbs = dd_block_sum(dd_make(nb), sh);
    if (threadIdx.x == 0) {
        double* r = scratch + (size_t)blockIdx.x * 8;
        r[0] = s1.hi; r[1] = s1.lo; r[2] = s2.hi; r[3] = s2.lo; r[4] = sb.hi; r[5] = sb.lo; r[6] = nbs.hi; r[7] = 0.0;
        __threadfence();
        last_cta = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last_cta) return;
    __threadfence();
    const int t = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (t < 4) {  // warp t folds quantity t over the CTAs: lane-strided partial sums, then a fixed shuffle tree
        dd acc = dd_make(0.0);
        for (unsigned int b = lane; b < gridDim.x; b += 32) {
            const double* r = scratch + (size_t)b * 8;
            acc = dd_add(acc, t < 3 ? dd_make(__ldcg(r + 2 * t), __ldcg(r + 2 * t + 1)) : dd_make(__ldcg(r + 6)));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc = dd_add(acc, dd_shfl_xor(acc, o));
        if (lane == 0) {
            if (t < 3) {
                out[2 * t] = acc.hi;
                out[2 * t + 1] = acc.lo;
            } else {
                out[6] = acc.hi;
                out[7] = 0.0;
                st->ticket = 0u;
            }
        }
    }
}

// The ranks' records (rank order) -> mean(all), mean(in band), std(all, ddof): s[2], s[3], s[4]; for the z-score moments
// (band == 0) s[5] = mean, s[6] = std.
__device__ __forceinline__ void moments_finish(dd s1, dd s2, dd sb, double nb, long long n_total, int band, int ddof, FinState* st) {
    const double n = (double)n_total;
    const dd mean = dd_div(s1, n);
    // sum (x - mean)^2 = sum x^2 - (sum x)^2 / n, every term carried to ~1e-32
    dd ss = dd_add(s2, dd_neg(dd_div(dd_mul(s1, s1), n)));
    double var = dd_value(dd_div(ss, n - (double)ddof));
    var = var > 0.0 ? var : (var == var ? 0.0 : var);
    const double sd = sqrt(var);
    if (band) {
        st->s[2] = dd_value(mean);
        st->s[3] = dd_value(dd_div(sb, nb));
        st->s[4] = sd;
    } else {
        st->s[5] = dd_value(mean);
        st->s[6] = sd;
    }
}
__global__ void moments_final_kernel(const double* __restrict__ parts, int world, long long n_total, int band, int ddof, FinState* st) {
    dd s1 = dd_make(0.0), s2 = dd_make(0.0), sb = dd_make(0.0);
    double nb = 0.0;
    for (int r = 0; r < world; ++r) {
        const double* p = parts + (size_t)r * 8;
        s1 = dd_add(s1, dd_make(p[0], p[1]));
        s2 = dd_add(s2, dd_make(p[2], p[3]));
        sb = dd_add(sb, dd_make(p[4], p[5]));
        nb += p[6];
    }
    moments_finish(s1, s2, sb, nb, n_total, band, ddof, st);
}

// ---------------------------------------------------------------------------------------------------------
// Short signals: the whole critic-score step -- select, band mean / std, z-score, smoothing -- and the score combination in ONE
// launch of one CTA.  A signal of a few thousand positions is launch-bound: the 16 launches of the staged chain above cost an
// order of magnitude more than their arithmetic.  Same arithmetic (exact select; (hi, lo) sums, whose rounded results do not
// depend on the order; the shared element-wise definitions), so the results are those of the staged chain bit for bit.
// Smoothing: a thread owns a run of consecutive positions, forms the first window's sum directly and slides it -- add the
// entering value, subtract the leaving one, both in (hi, lo) arithmetic -- and keeps the index of the last value change for
// pandas' all-equal-window rule (finish.cu).
// ---------------------------------------------------------------------------------------------------------
constexpr int SMALL_THREADS = 1024;
constexpr int SMALL_MAX = 65536;  // positions

template <typename TR>
__global__ void __launch_bounds__(SMALL_THREADS) critic_small_kernel(const double* __restrict__ x, int n_pos, int smooth_window, int keys_f32,
                                                                     int combine_mode, const TR* __restrict__ rec, const float* __restrict__ unorm,
                                                                     int n_windows, double* __restrict__ cs_out, double* __restrict__ final_out,
                                                                     double* __restrict__ zbuf, FinState* gstate) {
    __shared__ unsigned int hist[NQ][SEL_BINS];
    __shared__ long long wsum[NQ][8];
    __shared__ double sh[64];
    __shared__ FinState st;
    const int tid = threadIdx.x;
    if (tid == 0) sel_begin_body(&st, n_pos, keys_f32 ? 32 : 64);
    __syncthreads();
    // ---- radix select ---------------------------------------------------------------------------------
    const int key_bits = keys_f32 ? 32 : 64, passes = keys_f32 ? 3 : 6;
    for (int pass = 0; pass < passes; ++pass) {
        for (int e = tid; e < NQ * SEL_BINS; e += SMALL_THREADS) (&hist[0][0])[e] = 0u;
        __syncthreads();
        int shift, width;
        sel_digit(key_bits, pass, &shift, &width);
        const int hi_shift = shift + width;
        const unsigned int dmask = (1u << width) - 1u;
        for (int i = tid; i < n_pos; i += SMALL_THREADS) {
            const unsigned long long k = key_bits == 32 ? dkey32(x[i]) : dkey64(x[i]);
            const unsigned long long hi = hi_shift >= key_bits ? 0ull : (k >> hi_shift);
            const unsigned int digit = (unsigned int)(k >> shift) & dmask;
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                if (st.rep[q] == q && hi == st.prefix[q]) atomicAdd(&hist[q][digit], 1u);
        }
        __syncthreads();
        sel_pick_body(&hist[0][0], 1, pass, &st, wsum, tid);
        __syncthreads();
    }
    // ---- band mean, std ---------------------------------------------------------------------------------
    {
        const double lo = st.s[0], hi = st.s[1];
        dd s1 = dd_make(0.0), s2 = dd_make(0.0), sb = dd_make(0.0);
        double nb = 0.0;
        for (int i = tid; i < n_pos; i += SMALL_THREADS) {
            const double v = x[i];
            s1 = dd_add(s1, v);
            s2 = dd_add(s2, dd_prod(v, v));
            if (v >= lo && v <= hi) {
                sb = dd_add(sb, v);
                nb += 1.0;
            }
        }
        s1 = dd_block_sum(s1, sh);
        s2 = dd_block_sum(s2, sh);
        sb = dd_block_sum(sb, sh);
        const dd nbs = dd_block_sum(dd_make(nb), sh);
        if (tid == 0) moments_finish(s1, s2, sb, nbs.hi, n_pos, 1, 0, &st);
        __syncthreads();
    }
    if (tid < 8 && gstate) gstate->s[tid] = st.s[tid];  // for hypad_stats_read
    // ---- z-score, smoothing, combination ------------------------------------------------------------------
    const double mu = st.s[3], sd = st.s[4];
    // z once per position into scratch (a double division each): the smoothing reads every value up to `window` times
    for (int i = tid; i < n_pos; i += SMALL_THREADS) zbuf[i] = critic_z(x[i], mu, sd);
    __syncthreads();
    const int window = smooth_window, back = window / 2, fwd = (window - 1) / 2;
    const int need = window / 2 > 1 ? window / 2 : 1;  // min_periods = window // 2, at least one sample
    const int per = (n_pos + SMALL_THREADS - 1) / SMALL_THREADS;
    const int i0 = tid * per, i1 = i0 + per < n_pos ? i0 + per : n_pos;
    dd S = dd_make(0.0);
    int a = 0, b = 0;        // the current window [a, b)
    bool have = false;       // S, a, b, zlast, last_change describe the previous position's window
    double zlast = 0.0;      // z at b - 1
    long long last_change = -1;
    for (int i = i0; i < i1; ++i) {
        double out;
        int na = i - back, nb = i + fwd + 1;
        if (na < 0) na = 0;
        if (nb > n_pos) nb = n_pos;
        if (window <= 0 || nb - na < need) {
            out = nan("");
            have = false;
        } else {
            if (!have) {  // the window's sum from scratch
                S = dd_make(0.0);
                last_change = -1;
                zlast = na > 0 ? zbuf[na - 1] : 0.0;
                for (int j = na; j < nb; ++j) {
                    const double z = zbuf[j];
                    S = dd_add(S, z);
                    if (j == 0 || z != zlast) last_change = j;
                    zlast = z;
                }
                a = na;
                b = nb;
                have = true;
            } else {
                for (; b < nb; ++b) {  // values entering on the right
                    const double z = zbuf[b];
                    S = dd_add(S, z);
                    if (z != zlast) last_change = b;
                    zlast = z;
                }
                for (; a < na; ++a) S = dd_add(S, -zbuf[a]);  // values leaving on the left
            }
            // x[a .. b-1] all equal: pandas returns the value, not sum / count
            out = last_change <= a ? zlast : S.hi / (double)(b - a);
        }
        cs_out[i] = out;
        if (i < n_windows) {
            const TR rv = rec ? rec[i] : (TR)0;
            const double uv = unorm ? (double)unorm[i] : 0.0;
            final_out[i] = combine_value<TR>(combine_mode, out, rv, uv, 0.5);
        }
    }
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

int ensure_fin_state(hypad_ctx* ctx) {
    if (ctx->fin_state) return HYPAD_OK;
    // FinState | per-CTA scratch of the moments kernels (4 x SMs records of 8 doubles) | a local record / histogram for one rank
    const size_t bytes = 1024 + (size_t)kNumSMs * 8 * 4 * 8 * 2 + NQ * SEL_BINS * 4 + 256;
    HYPAD_CUDA_TRY(cudaMalloc(&ctx->fin_state, bytes));
    HYPAD_CUDA_TRY(cudaMemset(ctx->fin_state, 0, bytes));
    return HYPAD_OK;
}
FinState* fin_state(hypad_ctx* ctx) { return (FinState*)ctx->fin_state; }
double* fin_scalars(hypad_ctx* ctx) { return ((FinState*)ctx->fin_state)->s; }
static double* fin_scratch(hypad_ctx* ctx) { return (double*)((char*)ctx->fin_state + 1024); }
double* fin_local_record(hypad_ctx* ctx) { return (double*)((char*)ctx->fin_state + 1024 + (size_t)kNumSMs * 8 * 4 * 8 * 2); }

}  // namespace hypad

using namespace hypad;

extern "C" {

int hypad_stats_select_passes(int keys_f32) { return keys_f32 ? 3 : 6; }

int hypad_stats_select_begin(hypad_ctx* ctx, int64_t n_total, int keys_f32, void* stream) {
    HYPAD_REQUIRE(ctx && n_total >= 1, "hypad_stats_select_begin: bad argument");
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_fin_state(ctx);
    if (rc != HYPAD_OK) return rc;
    sel_begin_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(fin_state(ctx), (long long)n_total, keys_f32 ? 32 : 64);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_stats_select_hist(hypad_ctx* ctx, const double* x, int64_t len, int pass, uint32_t* hist, void* stream_) {
    HYPAD_REQUIRE(ctx && ctx->fin_state && hist && (x || len == 0) && len >= 0 && pass >= 0 && pass < 6,
                  "hypad_stats_select_hist: bad argument (hypad_stats_select_begin first)");
    cudaStream_t stream = (cudaStream_t)stream_;
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    HYPAD_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)HYPAD_SELECT_HIST_WORDS * 4, stream));
    if (len > 0) {
        sel_hist_kernel<<<red_grid(len), RB, 0, stream>>>(x, len, pass, fin_state(ctx), hist);
        HYPAD_LAUNCH_CHECK();
    }
    return HYPAD_OK;
}

int hypad_stats_select_pick(hypad_ctx* ctx, const uint32_t* hists, int world, int pass, void* stream) {
    HYPAD_REQUIRE(ctx && ctx->fin_state && hists && world >= 1 && pass >= 0 && pass < 6, "hypad_stats_select_pick: bad argument");
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    sel_pick_kernel<<<1, NQ * 256, 0, (cudaStream_t)stream>>>(hists, world, pass, fin_state(ctx));
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_stats_moments_partial(hypad_ctx* ctx, const void* x, int x_is_f32, int64_t len, int band, double* record, void* stream_) {
    HYPAD_REQUIRE(ctx && record && (x || len == 0) && len >= 0, "hypad_stats_moments_partial: bad argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_fin_state(ctx);
    if (rc != HYPAD_OK) return rc;
    if (len == 0) {
        HYPAD_CUDA_TRY(cudaMemsetAsync(record, 0, 64, stream));
        return HYPAD_OK;
    }
    const unsigned g = red_grid(len);
    if (x_is_f32) moments_partial_kernel<float><<<g, RB, 0, stream>>>((const float*)x, len, band, fin_state(ctx), fin_scratch(ctx), record);
    else moments_partial_kernel<double><<<g, RB, 0, stream>>>((const double*)x, len, band, fin_state(ctx), fin_scratch(ctx), record);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_stats_moments_final(hypad_ctx* ctx, const double* records, int world, int64_t n_total, int band, int ddof, void* stream) {
    HYPAD_REQUIRE(ctx && ctx->fin_state && records && world >= 1 && n_total >= 1, "hypad_stats_moments_final: bad argument");
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    moments_final_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(records, world, (long long)n_total, band, ddof, fin_state(ctx));
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_critic_small_max(void) { return SMALL_MAX; }

int hypad_critic_combine_small(hypad_ctx* ctx, const double* kmax, int64_t n_pos, int64_t smooth_window, int keys_f32, int combine_mode,
                               const float* rec, const float* unorm, int64_t n_windows, double* critic_scores, double* final,
                               void* stream) {
    HYPAD_REQUIRE(ctx && kmax && critic_scores && final, "hypad_critic_combine_small: NULL argument");
    HYPAD_REQUIRE(n_pos >= 1 && n_pos <= SMALL_MAX && n_windows >= 0 && n_windows <= n_pos, "hypad_critic_combine_small: %lld positions "
                  "outside 1..%d", (long long)n_pos, SMALL_MAX);
    HYPAD_REQUIRE(combine_mode >= 0 && combine_mode <= 8, "hypad_critic_combine_small: unknown mode %d", combine_mode);
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_fin_state(ctx);
    if (rc != HYPAD_OK) return rc;
    if ((rc = ensure_workspace(ctx, (size_t)n_pos * 8)) != HYPAD_OK) return rc;
    critic_small_kernel<float><<<1, SMALL_THREADS, 0, (cudaStream_t)stream>>>(kmax, (int)n_pos, (int)smooth_window, keys_f32, combine_mode, rec,
                                                                            unorm, (int)n_windows, critic_scores, final,
                                                                            (double*)ctx->workspace, fin_state(ctx));
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_stats_read(hypad_ctx* ctx, double* host8, void* stream) {
    HYPAD_REQUIRE(ctx && ctx->fin_state && host8, "hypad_stats_read: bad argument");
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    HYPAD_CUDA_TRY(cudaMemcpyAsync(host8, fin_scalars(ctx), 64, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    HYPAD_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return HYPAD_OK;
}

}  // extern "C"
