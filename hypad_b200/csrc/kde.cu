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
constexpr int KDE_CTAS = 7;  // resident CTAs per SM of the screened kernel (register budget 65536 / (7 x 128) = 73)

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

// The raw fp32 points of timestep i (0 beyond n): issued one iteration ahead, so the global-load latency hides behind the
// previous timestep's kernel evaluations.
__device__ __forceinline__ int fetch_points(const KdeArgs& a, int64_t i, int lane, float (&f)[4]) {
    const int64_t hi = i < a.n_windows - 1 ? i : a.n_windows - 1;
    const int64_t lo = i - a.S + 1 > 0 ? i - a.S + 1 : 0;
    const int n = (int)(hi - lo + 1);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = lane + 32 * q;
        f[q] = j < n ? a.critic[hi - j - a.critic_offset] : 0.0f;
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

// fp32 screening densities of up to 32 Q points, lane l owning points l, l + 32, ... (d[q], a sentinel far away where the
// timestep has no point).  K(a, b) = K(b, a): every unordered pair is evaluated ONCE and credited to both points.  Step s pairs
// the lane's Q points with the Q points of lane l + s (read from shared memory, where every group of 32 is stored twice back to
// back so that the rotated read is a constant offset from a per-lane base): Q^2 kernel values per step for Q loads and Q
// shuffles -- the values go to the lane's own sums, and into Q running sums that are handed from lane l + 1 to lane l after
// every step, so that they travel with the partner index; after 15 steps lane l holds the complete sums of lane l + 16's points.
// Steps 1..15 cover every unordered pair of lanes once, step 16 pairs l with l + 16 from both sides (own sums only).
// The special-function pipe (one MUFU.EX2 per pair) shares its issue path with shared-memory loads and shuffles: with one
// point per lane and step every pair cost one of each and the three pipes took turns (61 % XU, 60 % LSU busy); a Q x Q
// register tile amortises the load and the shuffle over Q pairs.
template <int Q>
__device__ __forceinline__ void screen_pairs(const float (&d)[4], const float* D, int lane, float (&e32)[4]) {
    const unsigned full = 0xffffffffu;
    const int nxt = (lane + 1) & 31;
    float e[Q], R[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        e[q] = 1.f;  // the point's own kernel value
        R[q] = 0.f;
    }
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int p = q + 1; p < Q; ++p) {  // the lane's own points against each other
            const float r = d[p] - d[q];
            const float val = ex2_approx(-(r * r));
            e[q] += val;
            e[p] += val;
        }
    const float* base = D + lane;
#pragma unroll
    for (int s = 1; s < 16; ++s) {
        float y[Q];
#pragma unroll
        for (int p = 0; p < Q; ++p) y[p] = base[64 * p + s];
#pragma unroll
        for (int p = 0; p < Q; ++p) {
            float acc = R[p];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float r = y[p] - d[q];
                const float val = ex2_approx(-(r * r));
                e[q] += val;
                acc += val;
            }
            R[p] = __shfl_sync(full, acc, nxt);
        }
    }
    {
        float y[Q];
#pragma unroll
        for (int p = 0; p < Q; ++p) y[p] = base[64 * p + 16];
#pragma unroll
        for (int p = 0; p < Q; ++p) {
            // after 15 hand-overs lane l holds the sums of lane l + 16's points: its owner fetches them
            e[p] += __shfl_sync(full, R[p], (lane + 16) & 31);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float r = y[p] - d[q];
                e[q] += ex2_approx(-(r * r));
            }
        }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) e32[q] = e[q];
}

__global__ void __launch_bounds__(KDE_WARPS * 32, KDE_CTAS) kde_screened_kernel(const KdeArgs a) {
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
    float nf[4];
    int nn = 0;
    {
        const int64_t it0 = (int64_t)blockIdx.x * KDE_WARPS + warp;
        if (it0 < a.t_count) nn = fetch_points(a, a.t0 + it0, lane, nf);
    }
    for (int64_t it = (int64_t)blockIdx.x * KDE_WARPS + warp; it < a.t_count; it += stride) {
        float cur[4];  // the raw points: what is compared and returned at the end (the float64 copies die after the set-up)
        const int n = nn;
#pragma unroll
        for (int q = 0; q < 4; ++q) cur[q] = nf[q];
        if (it + stride < a.t_count) nn = fetch_points(a, a.t0 + it + stride, lane, nf);
        if (n == 1) {
            if (lane == 0) a.out[it] = (double)cur[0];
            continue;
        }
        double mean, var, rcho;
        float d[4];
        const float kFar = 1e18f;  // sentinel for the slots beyond n: 2^-(1e36) = 0 against every real point
        {
            double v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = (double)cur[q];
            point_stats(v, n, lane, mean, var);
            if (!(var > 0.0)) {
                if (lane == 0) a.out[it] = v[0];
                continue;
            }
            const double cho = sqrt(var) * sScott[n];
            rcho = 1.0 / cho;
            const double dscale = kScale * rcho;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = lane + 32 * q;
                d[q] = j < n ? (float)((v[q] - mean) * dscale) : kFar;
                if (j < n) P[j] = v[q] * rcho;  // scipy divides; a last-bit difference of the scaled points cannot change the arg-max
            }
        }
        // ---- fp32 screening: every unordered pair once, with ex2.approx ----------------------------------
        float e32[4] = {0.f, 0.f, 0.f, 0.f};
        const int Q = n == 100 ? 3 : (n + 31) >> 5;  // n = 100: three full groups, the last four points separately
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q < Q) {
                D[64 * q + lane] = d[q];
                D[64 * q + 32 + lane] = d[q];
            }
        if (n == 100 && lane < 4) D[192 + lane] = d[3];
        __syncwarp();
        if (Q == 3) screen_pairs<3>(d, D, lane, e32);
        else if (Q == 4) screen_pairs<4>(d, D, lane, e32);
        else if (Q == 2) screen_pairs<2>(d, D, lane, e32);
        else screen_pairs<1>(d, D, lane, e32);
        if (n == 100) {
            // the interior case of the reference's window length: the last 4 points against the 96 (credited to both sides)
            // and against each other
            const unsigned full = 0xffffffffu;
            float L[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float dt = D[192 + t];
                L[t] = 0.f;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float r = dt - d[q];
                    const float val = ex2_approx(-(r * r));
                    e32[q] += val;
                    L[t] += val;
                }
            }
            {
                const float r = D[192 + ((lane >> 2) & 3)] - D[192 + (lane & 3)];
                const float val = lane < 16 ? ex2_approx(-(r * r)) : 0.f;  // includes the own value (1) on the diagonal
#pragma unroll
                for (int t = 0; t < 4; ++t) L[t] += (lane >> 2) == t ? val : 0.f;
            }
            // four warp sums for six shuffles: lanes trade halves (16), then quarters (8) of the four partial sums and finish the
            // one they are left with; lane l ends up with the total of point 96 + ((l >> 3) & 3)
            {
                const bool up16 = lane & 16, up8 = lane & 8;
                const float s0 = __shfl_xor_sync(full, up16 ? L[0] : L[2], 16), s1 = __shfl_xor_sync(full, up16 ? L[1] : L[3], 16);
                const float a0 = (up16 ? L[2] : L[0]) + s0, a1 = (up16 ? L[3] : L[1]) + s1;  // sums of points (2|0) and (3|1)
                const float s2 = __shfl_xor_sync(full, up8 ? a0 : a1, 8);
                float tot = (up8 ? a1 : a0) + s2;
                tot += __shfl_xor_sync(full, tot, 4);
                tot += __shfl_xor_sync(full, tot, 2);
                tot += __shfl_xor_sync(full, tot, 1);
                // lane bits (16, 8) -> point: (0,0) 0, (0,1) 1, (1,0) 2, (1,1) 3; lane t < 4 fetches point t's total
                const float mine = __shfl_sync(full, tot, ((lane & 2) << 3) | ((lane & 1) << 3));
                e32[3] = lane < 4 ? mine : 0.f;
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
        // Usually every candidate carries the same value: one candidate, or exact repeats of it (periodic and plateau signals
        // put the same critic value under a timestep more than once).  Equal values have bitwise equal densities and the first
        // index wins: decided without evaluating anything.
        bool all_same = false;
        int first_j = 0;
        {
            const int f0 = cands[0] ? __ffs(cands[0]) - 1 : cands[1] ? 31 + __ffs(cands[1]) : cands[2] ? 63 + __ffs(cands[2]) : 95 + __ffs(cands[3]);
            first_j = f0;
            float v0 = 0.0f;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if ((f0 >> 5) == q) v0 = cur[q];
            v0 = __shfl_sync(0xffffffffu, v0, f0 & 31);
            bool differs = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) differs |= ((cands[q] >> lane) & 1u) && cur[q] != v0;
            all_same = !__any_sync(0xffffffffu, differs);
        }
        if (!all_same) {
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
        int bj = first_j;
        unsigned keep[4] = {cands[0], cands[1], cands[2], cands[3]};
        bool decided = all_same;
        if (!all_same && ncand <= 32) {
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
        float val = 0.0f;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (bj == lane + 32 * q) val = cur[q];
        val = __shfl_sync(0xffffffffu, val, bj & 31);
        if (lane == 0) a.out[it] = (double)val;
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
    const int64_t cap = (int64_t)sms * KDE_CTAS;  // resident CTAs of 4 warps per SM, grid-stride beyond
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
