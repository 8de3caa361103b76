// Reconstruction errors of the Euclidean path: DTW (utils/anomaly_detection_utils.py:815-863), point-wise
// (:761-777) and area (:780-812) errors, all float64 like the reference.
//
// _dtw_error: L = (score_window/2)*2+1 (11 for the hard-coded score_window=10), both series zero-padded by
// h = L/2 on each side; for p in [0, len-L): dist_p = DTW(y_pad[p:p+L], yhat_pad[p:p+L]); the result is
// [0]*h + dists + [0]*(len - count - h), i.e. out[p+h] = dist_p and zeros elsewhere.
// pyts.metrics.dtw defaults (dist='square', method='classic'):  cost (a_r - b_j)^2,
//   acc[0][j] = acc[0][j-1] + cost ; acc[r][0] = acc[r-1][0] + cost ;
//   acc[r][j] = cost + min(acc[r-1][j-1], acc[r-1][j], acc[r][j-1]) ;  dtw = sqrt(acc[L-1][L-1]).
//
// The DP is 11 x 11: one thread per position keeps the rolling DP row and both 11-sample windows in
// registers (fully unrolled), so every lane is busy and there is no synchronisation at all -- a
// warp-per-pair anti-diagonal wavefront would idle 21 of 32 lanes on an 11-wide front.  For other window
// lengths (L <= 129) the generic kernel runs the same recurrence from local arrays.  Inputs are read through
// L1/L2 (each sample is touched by L neighbouring threads); the kernel is bound by its fp64 min/add chain.
#include "common.cuh"

namespace hypad {

template <typename TH>
__device__ __forceinline__ double padded(const TH* __restrict__ v, int64_t idx, int64_t len) {
    return (idx >= 0 && idx < len) ? (double)v[idx] : 0.0;
}

template <int L, typename TH>
__global__ void __launch_bounds__(256) dtw_fixed_kernel(const double* __restrict__ y, const TH* __restrict__ yh, int64_t len,
                                                        double* __restrict__ out) {
    constexpr int h = L / 2;
    const int64_t cnt = len - L > 0 ? len - L : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const int64_t p = i - h;  // window start in padded coordinates
        if (p < 0 || p >= cnt) {
            out[i] = 0.0;
            continue;
        }
        double a[L], b[L], row[L];
#pragma unroll
        for (int k = 0; k < L; ++k) {
            a[k] = padded(y, p + k - h, len);
            b[k] = padded(yh, p + k - h, len);
        }
        {
            double run = 0.0;
#pragma unroll
            for (int j = 0; j < L; ++j) {
                const double d = a[0] - b[j];
                run = __dadd_rn(run, __dmul_rn(d, d));
                row[j] = run;
            }
        }
#pragma unroll
        for (int r = 1; r < L; ++r) {
            double diag = row[0];
            {
                const double d = a[r] - b[0];
                row[0] = __dadd_rn(row[0], __dmul_rn(d, d));
            }
#pragma unroll
            for (int j = 1; j < L; ++j) {
                const double up = row[j];
                const double d = a[r] - b[j];
                row[j] = __dadd_rn(__dmul_rn(d, d), fmin(diag, fmin(up, row[j - 1])));
                diag = up;
            }
        }
        out[i] = sqrt(row[L - 1]);
    }
}

constexpr int DTW_MAXL = 129;

template <typename TH>
__global__ void __launch_bounds__(128) dtw_generic_kernel(const double* __restrict__ y, const TH* __restrict__ yh, int64_t len,
                                                          int L, double* __restrict__ out) {
    const int h = L / 2;
    const int64_t cnt = len - L > 0 ? len - L : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const int64_t p = i - h;
        if (p < 0 || p >= cnt) {
            out[i] = 0.0;
            continue;
        }
        double b[DTW_MAXL], row[DTW_MAXL];
        for (int k = 0; k < L; ++k) b[k] = padded(yh, p + k - h, len);
        const double a0 = padded(y, p - h, len);
        double run = 0.0;
        for (int j = 0; j < L; ++j) {
            const double d = a0 - b[j];
            run = __dadd_rn(run, __dmul_rn(d, d));
            row[j] = run;
        }
        for (int r = 1; r < L; ++r) {
            const double ar = padded(y, p + r - h, len);
            double diag = row[0];
            {
                const double d = ar - b[0];
                row[0] = __dadd_rn(row[0], __dmul_rn(d, d));
            }
            for (int j = 1; j < L; ++j) {
                const double up = row[j];
                const double d = ar - b[j];
                row[j] = __dadd_rn(__dmul_rn(d, d), fmin(diag, fmin(up, row[j - 1])));
                diag = up;
            }
        }
        out[i] = sqrt(row[L - 1]);
    }
}

template <typename TH>
__global__ void point_error_kernel(const double* __restrict__ y, const TH* __restrict__ yh, int64_t len, double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) out[i] = fabs(y[i] - (double)yh[i]);
}

// |rolling trapz(y) - rolling trapz(yhat)|, window sw centred ([i - sw/2, i + (sw-1)/2] clipped), min_periods sw/2,
// trapz with unit spacing = sum_k (v_k + v_{k+1}) / 2.
template <typename TH>
__global__ void area_error_kernel(const double* __restrict__ y, const TH* __restrict__ yh, int64_t len, int sw,
                                  double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int minp = sw / 2 > 1 ? sw / 2 : 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        int64_t a = i - sw / 2, b = i + (sw - 1) / 2 + 1;
        if (a < 0) a = 0;
        if (b > len) b = len;
        if (b - a < minp) {
            out[i] = nan("");
            continue;
        }
        double ty = 0.0, th = 0.0;
        for (int64_t k = a; k + 1 < b; ++k) {
            ty += (y[k + 1] + y[k]) / 2.0;
            th += ((double)yh[k + 1] + (double)yh[k]) / 2.0;
        }
        out[i] = fabs(ty - th);
    }
}

static unsigned ew_grid(int64_t n, int block) {
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want = ceil_div(n, block);
    if (want < 1) want = 1;
    const int64_t cap = (int64_t)sms * 16;
    return (unsigned)(want < cap ? want : cap);
}

}  // namespace hypad

using namespace hypad;

extern "C" {

int hypad_dtw_error(const double* y, const void* y_hat, int y_hat_is_f32, int64_t len, int score_window, double* out,
                    void* stream_) {
    HYPAD_REQUIRE(y && y_hat && out, "hypad_dtw_error: NULL argument");
    HYPAD_REQUIRE(len >= 0 && score_window >= 0, "hypad_dtw_error: bad shape");
    const int L = (score_window / 2) * 2 + 1;
    HYPAD_REQUIRE(L <= DTW_MAXL, "hypad_dtw_error: window length %d > %d", L, DTW_MAXL);
    if (len == 0) return HYPAD_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (L == 11) {
        if (y_hat_is_f32) dtw_fixed_kernel<11, float><<<ew_grid(len, 256), 256, 0, stream>>>(y, (const float*)y_hat, len, out);
        else dtw_fixed_kernel<11, double><<<ew_grid(len, 256), 256, 0, stream>>>(y, (const double*)y_hat, len, out);
    } else {
        if (y_hat_is_f32) dtw_generic_kernel<float><<<ew_grid(len, 128), 128, 0, stream>>>(y, (const float*)y_hat, len, L, out);
        else dtw_generic_kernel<double><<<ew_grid(len, 128), 128, 0, stream>>>(y, (const double*)y_hat, len, L, out);
    }
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_point_error(const double* y, const void* y_hat, int y_hat_is_f32, int64_t len, double* out, void* stream_) {
    HYPAD_REQUIRE(y && y_hat && out, "hypad_point_error: NULL argument");
    if (len <= 0) return len == 0 ? HYPAD_OK : HYPAD_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (y_hat_is_f32) point_error_kernel<float><<<ew_grid(len, 256), 256, 0, stream>>>(y, (const float*)y_hat, len, out);
    else point_error_kernel<double><<<ew_grid(len, 256), 256, 0, stream>>>(y, (const double*)y_hat, len, out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_area_error(const double* y, const void* y_hat, int y_hat_is_f32, int64_t len, int score_window, double* out,
                     void* stream_) {
    HYPAD_REQUIRE(y && y_hat && out, "hypad_area_error: NULL argument");
    HYPAD_REQUIRE(score_window >= 1, "hypad_area_error: score_window < 1");
    if (len <= 0) return len == 0 ? HYPAD_OK : HYPAD_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (y_hat_is_f32) area_error_kernel<float><<<ew_grid(len, 256), 256, 0, stream>>>(y, (const float*)y_hat, len, score_window, out);
    else area_error_kernel<double><<<ew_grid(len, 256), 256, 0, stream>>>(y, (const double*)y_hat, len, score_window, out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

}  // extern "C"
