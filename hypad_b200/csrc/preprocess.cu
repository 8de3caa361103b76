// The step in front of the scoring path: utils/dataloader.py:83-137 of the reference on the device.
//   time_segments_aggregate(method="mean")  :99-137   rows sorted by timestamp, segment k = rows with
//                                                     start_k <= t <= start_k + interval - 1 (pandas label slicing is
//                                                     inclusive), value = NaN-skipping mean, NaN for an empty segment
//   SimpleImputer() (mean)                  :86-87    NaN -> mean of the valid entries
//   MinMaxScaler(feature_range=(lo, hi))    :88-89    x * scale + (lo - min * scale), scale = (hi - lo) / (max - min)
// The reference walks the segments in a Python loop (`while start_ts <= max_ts`, one pandas slice per segment): ~2 s per
// 10^4 segments.  Here a thread owns a segment and finds its rows by binary search in the sorted timestamps.
#include "common.cuh"

namespace hypad {

// numpy's float64 add.reduce over a contiguous run (what pandas' nanmean calls after zeroing the NaNs): a plain
// left-to-right loop below 8 elements, 8 interleaved accumulators up to 128, recursive halving (to a multiple of 8) beyond.
__device__ __forceinline__ double np_sum_leaf(const double* __restrict__ v, int64_t n) {  // n <= 128
    auto val = [&](int64_t i) {
        const double x = v[i];
        return x != x ? 0.0 : x;
    };
    if (n < 8) {
        double r = 0.0;
        for (int64_t i = 0; i < n; ++i) r += val(i);
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = val(j);
    int64_t i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += val(i + j);
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += val(i);
    return res;
}

// The halving above 128 elements as an explicit post-order walk (a recursive device function would need a stack sized for the
// longest segment): frame = {offset, length, left sum, state}; the depth is < log2(n / 64) <= 40 for any n < 2^46.
__device__ double np_pairwise_sum(const double* __restrict__ v, int64_t n) {
    constexpr int DEPTH = 40;
    int64_t off[DEPTH], len[DEPTH];
    double left[DEPTH];
    signed char state[DEPTH];
    int sp = 0;
    bool returning = false;
    double ret = 0.0;
    off[0] = 0, len[0] = n, state[0] = 0;
    while (sp >= 0) {
        const int64_t m = len[sp];
        int64_t n2 = m / 2;
        n2 -= n2 % 8;
        if (!returning) {
            if (m <= 128 || sp == DEPTH - 1) {  // (the depth guard cannot trigger for addressable n)
                ret = np_sum_leaf(v + off[sp], m <= 128 ? m : 128);
                returning = true;
                --sp;
            } else {
                state[sp] = 1;
                off[sp + 1] = off[sp], len[sp + 1] = n2, state[sp + 1] = 0;
                ++sp;
            }
        } else if (state[sp] == 1) {
            left[sp] = ret;
            state[sp] = 2;
            off[sp + 1] = off[sp] + n2, len[sp + 1] = m - n2, state[sp + 1] = 0;
            ++sp;
            returning = false;
        } else {
            ret = left[sp] + ret;
            --sp;
        }
    }
    return ret;
}

__global__ void segments_aggregate_kernel(const double* __restrict__ ts, const double* __restrict__ values, int64_t n_rows,
                                          const double* __restrict__ seg_start, double interval, int64_t n_segments,
                                          double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_segments; k += stride) {
        const double s = seg_start[k];
        const double last = __dadd_rn(__dadd_rn(s, interval), -1.0);  // end_ts - 1, the inclusive upper label (:130-131)
        int64_t lo = 0, hi = n_rows;  // first row with ts >= s
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (ts[mid] < s) lo = mid + 1;
            else hi = mid;
        }
        const int64_t first = lo;
        hi = n_rows;  // first row with ts > last
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (ts[mid] <= last) lo = mid + 1;
            else hi = mid;
        }
        const int64_t m = lo - first;
        int64_t cnt = 0;
        for (int64_t i = first; i < lo; ++i) cnt += values[i] == values[i];
        out[k] = cnt ? np_pairwise_sum(values + first, m) / (double)cnt : __longlong_as_double(0x7ff8000000000000ll);
    }
}

// column statistics of the aggregated signal: {sum of valid, count of valid, min, max} -> st[0..3]
__global__ void __launch_bounds__(256) column_stats_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ partial) {
    __shared__ double sh[4][8];
    double s = 0.0, c = 0.0, mn = __longlong_as_double(0x7ff0000000000000ll), mx = -mn;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = x[i];
        if (v == v) {
            s += v;
            c += 1.0;
            mn = fmin(mn, v);
            mx = fmax(mx, v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) sh[0][w] = s, sh[1][w] = c, sh[2][w] = mn, sh[3][w] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) {
            s += sh[0][q];
            c += sh[1][q];
            mn = fmin(mn, sh[2][q]);
            mx = fmax(mx, sh[3][q]);
        }
        double* p = partial + 4 * blockIdx.x;
        p[0] = s, p[1] = c, p[2] = mn, p[3] = mx;
    }
}

// st: {mean of valid, scale, offset}; one thread folds the per-CTA partials in order (deterministic)
__global__ void column_stats_final_kernel(const double* __restrict__ partial, int n_blocks, double lo, double hi, double* __restrict__ st) {
    double s = 0.0, c = 0.0, mn = __longlong_as_double(0x7ff0000000000000ll), mx = -mn;
    for (int b = 0; b < n_blocks; ++b) {
        s += partial[4 * b];
        c += partial[4 * b + 1];
        mn = fmin(mn, partial[4 * b + 2]);
        mx = fmax(mx, partial[4 * b + 3]);
    }
    const double mean = s / c;  // NaN when nothing is valid, like the reference's imputer dropping the column
    // min / max of the imputed column are those of the valid entries (the mean lies between them)
    const double range = mx - mn;
    const double scale = (hi - lo) / (range != 0.0 ? range : 1.0);  // sklearn _handle_zeros_in_scale
    st[0] = mean;
    st[1] = scale;
    st[2] = __dadd_rn(lo, -__dmul_rn(mn, scale));
}

__global__ void impute_scale_kernel(const double* __restrict__ x, int64_t n, const double* __restrict__ st, double* __restrict__ out) {
    const double mean = st[0], scale = st[1], off = st[2];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = x[i];
        v = v == v ? v : mean;
        out[i] = __dadd_rn(__dmul_rn(v, scale), off);  // X *= scale_; X += min_  (two roundings, like sklearn)
    }
}

// ---- scipy.signal.detrend(type="linear") of the YAHOO branch (utils/dataloader.py:36-38, :66): subtract the least-squares line
// over t_i = (i + 1) / n.  Closed-form normal equations on centred abscissae (u_i = t_i - (n + 1) / (2 n), sum u_i^2 =
// (n^2 - 1) / (12 n)); scipy solves the same 2-column problem with LAPACK gelsd, the two agree to a few ulps of max|v|.
__global__ void __launch_bounds__(256) detrend_stats_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ partial) {
    __shared__ double sh[2][8];
    const double dn = (double)n, ubar = (dn + 1.0) / (2.0 * dn);
    double s = 0.0, suv = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = v[i];
        s += x;
        suv = fma((double)(i + 1) / dn - ubar, x, suv);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        suv += __shfl_xor_sync(0xffffffffu, suv, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) sh[0][w] = s, sh[1][w] = suv;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) s += sh[0][q], suv += sh[1][q];
        partial[2 * blockIdx.x] = s;
        partial[2 * blockIdx.x + 1] = suv;
    }
}

// st: {slope, intercept} of the fitted line slope * t + intercept
__global__ void detrend_final_kernel(const double* __restrict__ partial, int n_blocks, int64_t n, double* __restrict__ st) {
    double s = 0.0, suv = 0.0;
    for (int b = 0; b < n_blocks; ++b) s += partial[2 * b], suv += partial[2 * b + 1];
    const double dn = (double)n, ubar = (dn + 1.0) / (2.0 * dn);
    const double suu = (dn * dn - 1.0) / (12.0 * dn);
    const double slope = n > 1 ? suv / suu : 0.0;  // one sample: the minimum-norm fit reproduces it, the residual is 0
    st[0] = slope;
    st[1] = s / dn - slope * ubar;
}

__global__ void detrend_apply_kernel(const double* __restrict__ v, int64_t n, const double* __restrict__ st, double* __restrict__ out) {
    const double slope = st[0], icpt = st[1], dn = (double)n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = v[i] - fma(slope, (double)(i + 1) / dn, icpt);
}

}  // namespace hypad

using namespace hypad;

extern "C" {

int hypad_segments_aggregate(const double* ts_sorted, const double* values, int64_t n_rows, const double* seg_start, double interval,
                             int64_t n_segments, double* out, void* stream) {
    HYPAD_REQUIRE(ts_sorted && values && seg_start && out, "hypad_segments_aggregate: NULL argument");
    HYPAD_REQUIRE(n_rows >= 1 && n_segments >= 1 && interval > 0.0, "hypad_segments_aggregate: bad shape");
    const unsigned grid = (unsigned)((n_segments + 127) / 128 < 148 * 16 ? (n_segments + 127) / 128 : 148 * 16);
    segments_aggregate_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(ts_sorted, values, n_rows, seg_start, interval, n_segments, out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_impute_minmax(hypad_ctx* ctx, const double* x, int64_t n, double lo, double hi, double* out, void* stream_) {
    HYPAD_REQUIRE(ctx && x && out, "hypad_impute_minmax: NULL argument");
    HYPAD_REQUIRE(n >= 1 && hi > lo, "hypad_impute_minmax: bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    const int blocks = (int)((n + 256 * 8 - 1) / (256 * 8) < 592 ? (n + 256 * 8 - 1) / (256 * 8) : 592);
    int rc = ensure_workspace(ctx, (size_t)(4 * blocks + 4) * sizeof(double));
    if (rc != HYPAD_OK) return rc;
    double* partial = (double*)ctx->workspace;
    double* st = partial + 4 * blocks;
    column_stats_kernel<<<blocks, 256, 0, stream>>>(x, n, partial);
    HYPAD_LAUNCH_CHECK();
    column_stats_final_kernel<<<1, 1, 0, stream>>>(partial, blocks, lo, hi, st);
    HYPAD_LAUNCH_CHECK();
    impute_scale_kernel<<<blocks, 256, 0, stream>>>(x, n, st, out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_detrend_linear(hypad_ctx* ctx, const double* x, int64_t n, double* out, void* stream_) {
    HYPAD_REQUIRE(ctx && x && out, "hypad_detrend_linear: NULL argument");
    HYPAD_REQUIRE(n >= 1, "hypad_detrend_linear: bad length");
    cudaStream_t stream = (cudaStream_t)stream_;
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    const int blocks = (int)((n + 256 * 8 - 1) / (256 * 8) < 592 ? (n + 256 * 8 - 1) / (256 * 8) : 592);
    int rc = ensure_workspace(ctx, (size_t)(2 * blocks + 2) * sizeof(double));
    if (rc != HYPAD_OK) return rc;
    double* partial = (double*)ctx->workspace;
    double* st = partial + 2 * blocks;
    detrend_stats_kernel<<<blocks, 256, 0, stream>>>(x, n, partial);
    HYPAD_LAUNCH_CHECK();
    detrend_final_kernel<<<1, 1, 0, stream>>>(partial, blocks, n, st);
    HYPAD_LAUNCH_CHECK();
    detrend_apply_kernel<<<blocks, 256, 0, stream>>>(x, n, st, out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

}  // extern "C"
