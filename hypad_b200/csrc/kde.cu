// Overlap aggregation of critic scores by Gaussian-KDE arg-max (utils/anomaly_detection_utils.py:372-400,
// twin at :471-503), float64 like scipy.stats.gaussian_kde which the reference calls once per timestep.
//
// Timestep i aggregates the critic value of every window covering it: windows hi = min(i, N-1) down to
// lo = max(0, i-S+1) (the reference's j loop yields them in that, descending, order), n = hi-lo+1 <= S <= 128.
//   n == 1                      -> the value itself (np.median of one element)
//   variance == 0 (LinAlgError) -> np.median of identical values = the value
//   else                        -> V[argmax_j sum_i w * exp(-(P_i - P_j)^2 / 2) * norm],  P = V / cho,
//                                  cho = sqrt(var_ddof1) * n^(-1/5), w = 1/n, norm = (2 pi)^(-1/2) / cho,
//                                  first maximum wins (np.argmax).
// One warp owns one timestep; lanes hold points j = lane + 32 q.
//
// kde_exhaustive_kernel evaluates all n^2 kernel values in fp64, accumulating over i in scipy's order.
// kde_screened_kernel (the product path) first evaluates all densities in fp32 with ex2.approx on centred,
// bandwidth-scaled values, every unordered pair once; every j whose fp32 density is within 2.5e-5 + 1e-6 max|D|
// (relative) of the fp32 maximum is a candidate -- four times the worst-case fp32 error derived at the threshold below.
// Candidates are re-evaluated in fp32 from fp64 differences with expf (error < 1e-6); those within 1e-5 of that maximum
// survive, and only if several distinct values survive are they decided in fp64, in ascending j (first maximum wins),
// which equals the arg-max over all j.
#include <math_constants.h>

#include "common.cuh"

namespace hypad {

constexpr int KDE_WARPS = 4;
constexpr int KDE_MAXPTS = 128;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct KdeArgs {
    const float* critic;
    int64_t critic_offset, critic_len, n_windows, t0, t_count;
    double* out;
    int S;
};

// Loads the n points of timestep i into v[q] (lane + 32q), returns n.  Points beyond n are left 0.
__device__ __forceinline__ int load_points(const KdeArgs& a, int64_t i, int lane, double (&v)[4]) {
    const int64_t hi = i < a.n_windows - 1 ? i : a.n_windows - 1;
    const int64_t lo = i - a.S + 1 > 0 ? i - a.S + 1 : 0;
    const int n = (int)(hi - lo + 1);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = lane + 32 * q;
        v[q] = j < n ? (double)a.critic[hi - j - a.critic_offset] : 0.0;
    }
    return n;
}

// mean / ddof-1 variance of the warp's points (fp64)
__device__ __forceinline__ void point_stats(const double (&v)[4], int n, int lane, double& mean, double& var) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) s += v[q];  // padded entries are 0
    mean = warp_sum(s) / (double)n;
    double ss = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (lane + 32 * q < n) {
            const double d = v[q] - mean;
            ss += d * d;
        }
    var = warp_sum(ss) / (double)(n - 1);
}

// (best, j) arg-max across the warp with first-index tie-break; returns the winning j to every lane
__device__ __forceinline__ int warp_argmax_first(double best, int j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, j, o);
        if (ob > best || (ob == best && oj < j)) {
            best = ob;
            j = oj;
        }
    }
    return j;
}

// Scott factor n^(-1/5) for n = 1..128, computed once per CTA (a double-precision pow per timestep is ~10 % of the kernel)
__device__ __forceinline__ void fill_scott(double* tbl) {
    for (int j = threadIdx.x; j <= KDE_MAXPTS; j += blockDim.x) tbl[j] = j ? pow((double)j, -0.2) : 0.0;
    __syncthreads();
}

__global__ void __launch_bounds__(KDE_WARPS * 32) kde_exhaustive_kernel(const KdeArgs a) {
    __shared__ double sP[KDE_WARPS][KDE_MAXPTS];
    __shared__ double sScott[KDE_MAXPTS + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* P = sP[warp];
    fill_scott(sScott);
    const int64_t stride = (int64_t)gridDim.x * KDE_WARPS;
    for (int64_t it = (int64_t)blockIdx.x * KDE_WARPS + warp; it < a.t_count; it += stride) {
        const int64_t i = a.t0 + it;
        double v[4];
        const int n = load_points(a, i, lane, v);
        if (n == 1) {
            if (lane == 0) a.out[it] = v[0];
            continue;
        }
        double mean, var;
        point_stats(v, n, lane, mean, var);
        if (!(var > 0.0)) {  // scipy: Cholesky of [[0]] raises LinAlgError -> np.median of identical values
            if (lane == 0) a.out[it] = v[0];
            continue;
        }
        const double cho = sqrt(var) * sScott[n];
        const double norm = 0.3989422804014327 / cho;  // (2 pi)^(-1/2) / cho
        const double w = 1.0 / (double)n;
        double p[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            p[q] = v[q] / cho;
            if (lane + 32 * q < n) P[lane + 32 * q] = p[q];
        }
        __syncwarp();
        double est[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < n; ++k) {
            const double pk = P[k];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double r = pk - p[q];
                const double e = __dmul_rn(exp(-__dmul_rn(r, r) / 2.0), norm);
                est[q] = __dadd_rn(est[q], __dmul_rn(w, e));
            }
        }
        double best = -1.0;
        int bj = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = lane + 32 * q;
            if (j < n && est[q] > best) {
                best = est[q];
                bj = j;
            }
        }
        bj = warp_argmax_first(best, bj);
        // the winner's value: lane bj%32 holds it in v[bj/32]
        double val = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (bj == lane + 32 * q) val = v[q];
        val = __shfl_sync(0xffffffffu, val, bj & 31);
        if (lane == 0) a.out[it] = val;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(KDE_WARPS * 32, 8) kde_screened_kernel(const KdeArgs a) {
    __shared__ double sP[KDE_WARPS][KDE_MAXPTS];
    __shared__ float sD[KDE_WARPS][2 * KDE_MAXPTS];
    __shared__ double sScott[KDE_MAXPTS + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* P = sP[warp];
    float* D = sD[warp];
    fill_scott(sScott);
    const int64_t stride = (int64_t)gridDim.x * KDE_WARPS;
    // exp(-r^2/2) = 2^(-(s r)^2) with s = sqrt(log2(e)/2): the scale is folded into the stored values, so one kernel
    // evaluation is FADD, FMUL (negated), MUFU.EX2, FADD
    const double kScale = 0.8493218002880191;
    for (int64_t it = (int64_t)blockIdx.x * KDE_WARPS + warp; it < a.t_count; it += stride) {
        const int64_t i = a.t0 + it;
        double v[4];
        const int n = load_points(a, i, lane, v);
        if (n == 1) {
            if (lane == 0) a.out[it] = v[0];
            continue;
        }
        double mean, var;
        point_stats(v, n, lane, mean, var);
        if (!(var > 0.0)) {
            if (lane == 0) a.out[it] = v[0];
            continue;
        }
        const double cho = sqrt(var) * sScott[n];
        const double rcho = 1.0 / cho;
        const double dscale = kScale * rcho;
        float d[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = lane + 32 * q;
            d[q] = (float)((v[q] - mean) * dscale);
            if (j < n) P[j] = v[q] * rcho;  // scipy divides; a last-bit difference of the scaled points cannot change the arg-max
        }
        // ---- fp32 screening: all n^2 kernel values with ex2.approx -------------------------------------
        float e32[4] = {0.f, 0.f, 0.f, 0.f};
        if (n == 100) {
            // The interior case of the reference's window length.  K(a,b) = K(b,a): every unordered pair is evaluated ONCE and
            // credited to both points.  Points 0..95 form three groups of 32 (lane l owns point l of each group).  For a pair of
            // groups (X, Y), step s pairs x_l with y_(l+s): the value goes to lane l's own sum for x_l, and into a running sum R
            // that is handed from lane l+1 to lane l before every step, so that after 32 steps it has collected all 32
            // contributions of one y and one more rotation delivers it to its owner -- one shuffle instead of one MUFU.EX2.
            // Inside a group, steps 1..15 cover every unordered pair once and step 16 pairs l with l+16 from both sides.
            // The groups are stored twice back to back so that the rotated read is a constant offset from a per-lane base.
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                D[64 * q + lane] = d[q];
                D[64 * q + 32 + lane] = d[q];
            }
            if (lane < 4) D[192 + lane] = d[3];
            __syncwarp();
            const unsigned full = 0xffffffffu;
            const int nxt = (lane + 1) & 31;
            float e[3] = {1.f, 1.f, 1.f};  // the point's own kernel value
#pragma unroll
            for (int X = 0; X < 3; ++X) {
                {   // inside group X: the running sum is handed on after every step
                    const float* px = D + 64 * X + lane;
                    float R = 0.f;
#pragma unroll
                    for (int s = 1; s < 16; ++s) {
                        const float r = px[s] - d[X];
                        const float val = ex2_approx(-(r * r));
                        e[X] += val;
                        R = __shfl_sync(full, R + val, nxt);
                    }
                    e[X] += __shfl_sync(full, R, (lane + 16) & 31);  // after 15 hand-overs lane l holds the sum of point l+16
                    const float r = px[16] - d[X];
                    e[X] += ex2_approx(-(r * r));
                }
#pragma unroll
                for (int Y = X + 1; Y < 3; ++Y) {
                    const float* py = D + 64 * Y + lane;
                    float R = 0.f;
#pragma unroll
                    for (int s0 = 0; s0 < 32; s0 += 8) {
                        float y[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) y[u] = py[s0 + u];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float r = y[u] - d[X];
                            const float val = ex2_approx(-(r * r));
                            e[X] += val;
                            R = __shfl_sync(full, R + val, nxt);
                        }
                    }
                    e[Y] += R;  // 32 hand-overs: back at the owner
                }
            }
            // the last 4 points against the 96 (credited to both sides) and against each other
            float L[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float dt = D[192 + t];
                L[t] = 0.f;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float r = dt - d[q];
                    const float val = ex2_approx(-(r * r));
                    e[q] += val;
                    L[t] += val;
                }
            }
            {
                const float r = D[192 + ((lane >> 2) & 3)] - D[192 + (lane & 3)];
                const float val = lane < 16 ? ex2_approx(-(r * r)) : 0.f;  // includes the own value (1) on the diagonal
#pragma unroll
                for (int t = 0; t < 4; ++t) L[t] += (lane >> 2) == t ? val : 0.f;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) L[t] += __shfl_xor_sync(full, L[t], o);
                if (lane == t) e32[3] = L[t];
            }
            e32[0] = e[0];
            e32[1] = e[1];
            e32[2] = e[2];
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (lane + 32 * q < n) D[lane + 32 * q] = d[q];
            __syncwarp();
            for (int k = 0; k < n; ++k) {
                const float dk = D[k];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float r = dk - d[q];
                    e32[q] += ex2_approx(-(r * r));
                }
            }
        }
        float m32 = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (lane + 32 * q < n) m32 = fmaxf(m32, e32[q]);
        m32 = warp_max(m32);
        // Stage-1 margin.  Worst-case relative error of an fp32 density: ~90 sequential fp32 additions of positive terms
        // (5.4e-6), ex2.approx (2.4e-7), and the rounding of the centred operands, their difference r and r^2, which moves
        // a term's exponent by 2^-24 (2 r (|D_k| + |D_j| + r) + r^2) -- weighted by the term itself (r 2^(-r^2) <= 0.52,
        // r^2 2^(-r^2) <= 0.53) at most 4.1e-8 (2.1 Dmax + 1.6): 6.3e-6 + 8.6e-8 Dmax in all, Dmax = max |D|.  Two such
        // errors separate a true maximum from its fp32 rank; the margin is twice that again.
        float dmax = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (lane + 32 * q < n) dmax = fmaxf(dmax, fabsf(d[q]));
        dmax = warp_max(dmax);
        const float thr = m32 * (1.0f - (2.5e-5f + 1e-6f * dmax));
        unsigned cands[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) cands[q] = __ballot_sync(0xffffffffu, (lane + 32 * q < n) && e32[q] >= thr);
        // ---- stage 2: the candidates again, in fp32 but from the fp64 differences and with expf -----------------
        // Relative error of such a density: expf 2 ulp (2.4e-7), the rounded difference and its square perturb a term's
        // exponent b = r^2/2 by <= 1.8e-7 b, i.e. the term by <= 0.37 x 1.8e-7 whatever b is, (n/32 + 5) fp32 additions of
        // positive terms <= 6e-7: < 1e-6 in all.  Every candidate within 1e-5 of the stage-2 maximum survives (5x the
        // margin two such errors need) -- usually exactly one distinct value, which then IS the arg-max and needs no fp64.
        // A candidate whose value repeats a recently evaluated one has bitwise the same density and loses the tie to the
        // earlier index, so it is skipped (periodic and plateau signals produce many exact repeats).
        // Lane c keeps the c-th distinct candidate; more than 32 of them (a flat-topped density) fall back to fp64 for all.
        int ncand = 0, myj = 0;
        float myest = -1.0f;
        {
            // the distinct candidates first (lane c keeps the c-th): a single one needs no evaluation at all
            double seen0 = CUDART_NAN, seen1 = CUDART_NAN, seen2 = CUDART_NAN, seen3 = CUDART_NAN;
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                unsigned cand = q == 0 ? cands[0] : q == 1 ? cands[1] : q == 2 ? cands[2] : cands[3];
                while (cand) {
                    const int src = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const int j = src + 32 * q;
                    const double pj = P[j];
                    if (pj == seen0 || pj == seen1 || pj == seen2 || pj == seen3) continue;
                    seen3 = seen2;
                    seen2 = seen1;
                    seen1 = seen0;
                    seen0 = pj;
                    if (lane == ncand) myj = j;
                    ++ncand;
                }
            }
            if (ncand > 1) {
                const int ne = ncand < 32 ? ncand : 32;
#pragma unroll 1
                for (int c = 0; c < ne; ++c) {
                    const double pj = P[__shfl_sync(0xffffffffu, myj, c)];
                    float part = 0.0f;
                    for (int k = lane; k < n; k += 32) {
                        const float r = (float)(P[k] - pj);
                        part += expf(-0.5f * (r * r));
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    if (lane == c) myest = part;
                }
            } else if (lane == 0) {
                myest = 1.0f;  // the only candidate survives
            }
        }
        int bj = 0;
        unsigned keep[4] = {cands[0], cands[1], cands[2], cands[3]};
        bool decided = false;
        if (ncand <= 32) {
            const float m2 = warp_max(myest);
            unsigned surv = __ballot_sync(0xffffffffu, lane < ncand && myest >= m2 * (1.0f - 1e-5f));
            if (__popc(surv) == 1) {
                bj = __shfl_sync(0xffffffffu, myj, __ffs(surv) - 1);
                decided = true;
            } else {
                keep[0] = keep[1] = keep[2] = keep[3] = 0u;
                while (surv) {
                    const int c = __ffs(surv) - 1;
                    surv &= surv - 1;
                    const int j = __shfl_sync(0xffffffffu, myj, c);
                    const unsigned bit = 1u << (j & 31);
                    if ((j >> 5) == 0) keep[0] |= bit;
                    else if ((j >> 5) == 1) keep[1] |= bit;
                    else if ((j >> 5) == 2) keep[2] |= bit;
                    else keep[3] |= bit;
                }
            }
        }
        // ---- fp64 decision among the survivors, ascending j (first maximum wins) ----------------------------------
        // (the positive constants w = 1/n and norm = (2 pi)^(-1/2)/cho multiply every density alike: left out)
        if (!decided) {
            double best = -1.0;
            double seen0 = CUDART_NAN, seen1 = CUDART_NAN, seen2 = CUDART_NAN, seen3 = CUDART_NAN;
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                unsigned cand = q == 0 ? keep[0] : q == 1 ? keep[1] : q == 2 ? keep[2] : keep[3];
                while (cand) {
                    const int src = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const int j = src + 32 * q;
                    const double pj = P[j];
                    if (pj == seen0 || pj == seen1 || pj == seen2 || pj == seen3) continue;
                    seen3 = seen2;
                    seen2 = seen1;
                    seen1 = seen0;
                    seen0 = pj;
                    double part = 0.0;
                    for (int k = lane; k < n; k += 32) {
                        const double r = P[k] - pj;
                        part += exp(-0.5 * (r * r));
                    }
                    const double tot = warp_sum(part);  // xor tree: bitwise identical in every lane
                    if (tot > best) {
                        best = tot;
                        bj = j;
                    }
                }
            }
        }
        double val = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (bj == lane + 32 * q) val = v[q];
        val = __shfl_sync(0xffffffffu, val, bj & 31);
        if (lane == 0) a.out[it] = val;
        __syncwarp();
    }
}

static int check_args(const float* critic, int64_t critic_offset, int64_t critic_len, int64_t n_windows, int S, int64_t t0,
                      int64_t t_count, double* kmax) {
    HYPAD_REQUIRE(critic && kmax, "kde: NULL argument");
    HYPAD_REQUIRE(S >= 1 && S <= KDE_MAXPTS, "kde: S=%d outside 1..%d", S, KDE_MAXPTS);
    HYPAD_REQUIRE(n_windows >= 1, "kde: n_windows < 1");
    HYPAD_REQUIRE(t0 >= 0 && t_count >= 0 && t0 + t_count <= n_windows + S - 1, "kde: timestep range outside [0, N+S-1)");
    if (t_count > 0) {
        const int64_t first = t0 - S + 1 > 0 ? t0 - S + 1 : 0;
        const int64_t last = t0 + t_count - 1 < n_windows - 1 ? t0 + t_count - 1 : n_windows - 1;
        HYPAD_REQUIRE(critic_offset <= first && critic_offset + critic_len > last,
                      "kde: critic slice [%lld,%lld) does not cover windows [%lld,%lld]", (long long)critic_offset,
                      (long long)(critic_offset + critic_len), (long long)first, (long long)last);
    }
    return HYPAD_OK;
}

template <typename K>
static int launch_kde(K kernel, const float* critic, int64_t critic_offset, int64_t critic_len, int64_t n_windows, int S,
                      int64_t t0, int64_t t_count, double* kmax, cudaStream_t stream) {
    int rc = check_args(critic, critic_offset, critic_len, n_windows, S, t0, t_count, kmax);
    if (rc != HYPAD_OK || t_count == 0) return rc;
    KdeArgs a{critic, critic_offset, critic_len, n_windows, t0, t_count, kmax, S};
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = ceil_div(t_count, KDE_WARPS);
    const int64_t cap = (int64_t)sms * 8;  // 8 CTAs of 4 warps per SM, grid-stride beyond
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    kernel<<<grid, KDE_WARPS * 32, 0, stream>>>(a);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

}  // namespace hypad

extern "C" {

int hypad_kde_argmax_overlap(const float* critic, int64_t critic_offset, int64_t critic_len, int64_t n_windows, int S,
                             int64_t t0, int64_t t_count, double* kmax, void* stream) {
    return hypad::launch_kde(hypad::kde_screened_kernel, critic, critic_offset, critic_len, n_windows, S, t0, t_count, kmax,
                             (cudaStream_t)stream);
}

int hypad_kde_argmax_overlap_exhaustive(const float* critic, int64_t critic_offset, int64_t critic_len, int64_t n_windows,
                                        int S, int64_t t0, int64_t t_count, double* kmax, void* stream) {
    return hypad::launch_kde(hypad::kde_exhaustive_kernel, critic, critic_offset, critic_len, n_windows, S, t0, t_count, kmax,
                             (cudaStream_t)stream);
}

}  // extern "C"
