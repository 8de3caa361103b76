// O(T) float64 finishing steps: critic z-score with quantile band (utils/anomaly_detection_utils.py:307-333),
// centred rolling mean (pandas rolling(center=True).mean(), :326-331 / :954-961), z-score + clip (:523-524),
// score combination (:336-362, :554-570) and the per-window statistics / run extraction of find_anomalies
// (:1098-1166).  All results stay on the device; scalars travel through a small block of the workspace.
#include <cooperative_groups.h>

#include "common.cuh"
#include "dd.cuh"
#include "finish_common.cuh"

namespace hypad {

constexpr int RB = 256;  // threads of the reduction kernels

// critic_stats.cu
int ensure_fin_state(hypad_ctx* ctx);
double* fin_scalars(hypad_ctx* ctx);
double* fin_local_record(hypad_ctx* ctx);

// barrier of the first RB threads of a CTA (all of a 256-thread CTA; the statistics group of the fused thresholding kernel)
__device__ __forceinline__ void group_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp] = v;
    group_sync();
    double t = 0.0;
    if (warp == 0) {
        t = lane < (RB >> 5) ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) sh[0] = t;
    }
    group_sync();
    t = sh[0];
    group_sync();
    return t;
}

template <typename T>
__device__ __forceinline__ double ld(const T* p, int64_t i) { return (double)p[i]; }

__device__ __forceinline__ unsigned long long dkey(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// out = clip((x - s[5]) / s[6], 0) + 1   (the scalars of critic_stats.cu)
template <typename T>
__global__ void zscore_clip_kernel(const T* __restrict__ x, int64_t len, const double* s, double* __restrict__ out) {
    const double mu = s[5], sd = s[6];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const double z = (ld(x, i) - mu) / sd;
        out[i] = (z > 0.0 ? z : (z != z ? z : 0.0)) + 1.0;
    }
}

// ---------------------------------------------------------------------------------------------------------
// centred rolling mean (pandas rolling(window, center=True, min_periods).mean()) through a two-level prefix sum.
// pandas keeps a Kahan-compensated running sum and, when every value in the window is the same, returns that value itself
// (roll_mean / calc_mean in pandas/_libs/window/aggregations.pyx: num_consecutive_same_value >= nobs).  Both matter to the
// thresholding downstream: a flat stretch of the signal gives a flat stretch of scores, and a residue of a few ulps there
// decides whether `errors > mean + 4 std` holds on a window whose std is itself a few ulps.  So the prefix sums are carried as
// unevaluated (hi, lo) pairs (error-free TwoSum: the window sum is the correctly rounded difference of two exact prefixes up
// to ~1e-32 of their magnitude) and a max-scan carries the index of the last position whose value differs from its
// predecessor, which answers "is the window constant" in O(1).
// ---------------------------------------------------------------------------------------------------------
constexpr int SCAN_CHUNK = 2048;  // elements per CTA (256 threads x 8)

struct ScanBufs {
    double* pre_hi;      // [len] chunk-local inclusive prefix
    double* pre_lo;
    long long* lc;       // [len] chunk-local index of the last value change at or before i (-1: none in this chunk so far)
    double* tot_hi;      // [nchunks] chunk totals -> exclusive prefix of the totals
    double* tot_lo;
    long long* tot_lc;   // [nchunks] last change inside the chunk -> last change before the chunk
};

// the scanned value: x itself, or the critic z-score |x - s[3]| / s[4] + 1 (:322-325) taken on the fly
template <bool Z>
__device__ __forceinline__ double scan_value(const double* __restrict__ x, int64_t i, double mu, double sd) {
    return Z ? critic_z(x[i], mu, sd) : x[i];
}

template <bool Z>
__global__ void __launch_bounds__(256) scan_local_kernel(const double* __restrict__ x, int64_t len, ScanBufs sb, const double* scal) {
    const double mu = Z ? scal[3] : 0.0, sd = Z ? scal[4] : 1.0;
    __shared__ double sh_hi[256], sh_lo[256];
    __shared__ long long sh_lc[256];
    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + tid * 8;
    dd v[8];
    long long c[8];
    dd run;
    run.hi = run.lo = 0.0;
    long long last = -1;
    double prev = base > 0 && base - 1 < len ? scan_value<Z>(x, base - 1, mu, sd) : 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int64_t i = base + k;
        const double xv = i < len ? scan_value<Z>(x, i, mu, sd) : 0.0;
        run = dd_add(run, xv);
        v[k] = run;
        if (i < len && (i == 0 || xv != prev)) last = i;
        c[k] = last;
        prev = xv;
    }
    sh_hi[tid] = run.hi;
    sh_lo[tid] = run.lo;
    sh_lc[tid] = last;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {  // Hillis-Steele inclusive scan of the thread sums / last changes
        dd t;
        t.hi = t.lo = 0.0;
        long long l = -1;
        if (tid >= o) {
            t.hi = sh_hi[tid - o];
            t.lo = sh_lo[tid - o];
            l = sh_lc[tid - o];
        }
        __syncthreads();
        if (tid >= o) {
            dd m;
            m.hi = sh_hi[tid];
            m.lo = sh_lo[tid];
            m = dd_add(t, m);
            sh_hi[tid] = m.hi;
            sh_lo[tid] = m.lo;
            sh_lc[tid] = l > sh_lc[tid] ? l : sh_lc[tid];
        }
        __syncthreads();
    }
    dd off;
    off.hi = tid ? sh_hi[tid - 1] : 0.0;
    off.lo = tid ? sh_lo[tid - 1] : 0.0;
    const long long loff = tid ? sh_lc[tid - 1] : -1;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (base + k < len) {
            const dd r = dd_add(off, v[k]);
            sb.pre_hi[base + k] = r.hi;
            sb.pre_lo[base + k] = r.lo;
            sb.lc[base + k] = c[k] > loff ? c[k] : loff;
        }
    if (tid == 255) {
        sb.tot_hi[blockIdx.x] = sh_hi[255];
        sb.tot_lo[blockIdx.x] = sh_lo[255];
        sb.tot_lc[blockIdx.x] = sh_lc[255];
    }
}
// exclusive scan of the chunk totals in place (one warp, contiguous segments per lane)
__global__ void __launch_bounds__(32) scan_totals_kernel(ScanBufs sb, int64_t nchunks) {
    const int lane = threadIdx.x;
    const int64_t per = (nchunks + 31) / 32;
    const int64_t b = lane * per, e = b + per < nchunks ? b + per : nchunks;
    dd s;
    s.hi = s.lo = 0.0;
    long long l = -1;
    for (int64_t i = b; i < e; ++i) {
        dd t;
        t.hi = sb.tot_hi[i];
        t.lo = sb.tot_lo[i];
        s = dd_add(s, t);
        l = sb.tot_lc[i] > l ? sb.tot_lc[i] : l;
    }
    dd run;  // exclusive prefix over the lanes
    run.hi = run.lo = 0.0;
    long long lrun = -1;
    for (int src = 0; src < 31; ++src) {
        dd t;
        t.hi = __shfl_sync(0xffffffffu, s.hi, src);
        t.lo = __shfl_sync(0xffffffffu, s.lo, src);
        const long long tl = __shfl_sync(0xffffffffu, l, src);
        if (lane > src) {
            run = dd_add(run, t);
            lrun = tl > lrun ? tl : lrun;
        }
    }
    for (int64_t i = b; i < e; ++i) {
        dd t;
        t.hi = sb.tot_hi[i];
        t.lo = sb.tot_lo[i];
        const long long tl = sb.tot_lc[i];
        sb.tot_hi[i] = run.hi;
        sb.tot_lo[i] = run.lo;
        sb.tot_lc[i] = lrun;
        run = dd_add(run, t);
        lrun = tl > lrun ? tl : lrun;
    }
}
__device__ __forceinline__ dd scan_prefix(const ScanBufs& sb, int64_t i) {  // exact-ish sum of x[0..i]
    dd p, o;
    p.hi = sb.pre_hi[i];
    p.lo = sb.pre_lo[i];
    o.hi = sb.tot_hi[i / SCAN_CHUNK];
    o.lo = sb.tot_lo[i / SCAN_CHUNK];
    return dd_add(o, p);
}
// x / sb cover the global positions [ext0, ext0 + ext_len) of an array of n_total positions; output for the `count` positions
// from p0 on (the caller guarantees that their windows, clipped to [0, n_total), lie inside the covered range).
template <bool Z>
__global__ void rolling_mean_kernel(const double* __restrict__ x, ScanBufs sb, const double* scal, int64_t ext0, int64_t n_total,
                                    int64_t p0, int64_t count, int64_t window, int64_t min_periods, double* __restrict__ out) {
    const double mu = Z ? scal[3] : 0.0, sd = Z ? scal[4] : 1.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t back = window / 2, fwd = (window - 1) / 2;
    const int64_t need = min_periods > 1 ? min_periods : 1;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) {
        const int64_t i = p0 + j;
        int64_t a = i - back, b = i + fwd + 1;
        if (a < 0) a = 0;
        if (b > n_total) b = n_total;
        const int64_t cnt = b - a;
        if (window <= 0 || cnt < need) {
            out[j] = nan("");
            continue;
        }
        a -= ext0;
        b -= ext0;
        const long long lc = sb.lc[b - 1], lo_c = sb.tot_lc[(b - 1) / SCAN_CHUNK];
        if ((lc > lo_c ? lc : lo_c) <= a) {  // x[a..b-1] all equal: pandas returns the value, not sum / count
            out[j] = scan_value<Z>(x, b - 1, mu, sd);
            continue;
        }
        dd s = scan_prefix(sb, b - 1);
        if (a > 0) s = dd_add(s, dd_neg(scan_prefix(sb, a - 1)));
        out[j] = s.hi / (double)cnt;
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
        const TR rv = r ? r[i] : (TR)0;
        const double uv = u ? (double)u[i] : 0.0;
        const double o = combine_value<TR>(mode, cv, rv, uv, lambda_rec);
        out[i] = o;
    }
}

// ---------------------------------------------------------------------------------------------------------
// find_anomalies: per analysis window mean / std / threshold, dilated above-threshold runs, max_below.
// Everything is tile-parallel: two-stage sums for the statistics, then one pass that classifies every element
// (above threshold -> inside a padded run or not) from a bit-packed flag window in shared memory and emits the run
// starts / ends as unordered events; a second pass attributes every above-threshold value to its run
// (the run with the largest start <= its position) for the run maximum; a tiny per-window kernel orders the few runs.
// ---------------------------------------------------------------------------------------------------------
constexpr int TW_CHUNK = 4096;   // elements per CTA of the two-stage sums
constexpr int TW_TILE = 1024;    // elements per CTA of the event pass
constexpr int TW_MAXPAD = 512;

struct TwArgs {
    const double* errors;
    int64_t len, window_size, step;
    int n_analysis, ddof, pad, max_runs, n_slices;
    int stats_f32;          // mean / std / threshold rounded to fp32 and combined in fp32 (find_anomalies on an fp32 torch tensor)
    int64_t k0;             // index of the first analysis window handled (outputs are indexed from 0)
    double* stats;          // [n_analysis][4]
    double* runs;           // [n_analysis][max_runs][3]
    int32_t* n_runs;        // [n_analysis]
    // workspace
    double* partial;        // [n_analysis][n_slices]
    double* mean;           // [n_analysis]
    int* cnt;               // [n_analysis][2] starts, ends
    unsigned long long* below;  // [n_analysis]
    long long* starts;      // [n_analysis][max_runs]
    long long* ends;        // [n_analysis][max_runs]
    unsigned long long* rmax;   // [n_analysis][max_runs]
    // block summaries and work list of the product path
    double* bsum1;              // [blocks] sum(x - c)
    double* bsum2;              // [blocks] sum((x - c)^2)
    unsigned long long* bmax;   // [blocks] order-preserving key of the block maximum
    int* work_cnt;
    longlong2* work;            // (window, block) pairs that need the element-wise pass
    // One rank's share of an array sharded by contiguous, block-aligned ranges (hypad_tw_shard_*): `errors` is then a VIRTUAL
    // base pointer -- only the positions [lim_lo, lim_hi) (the own ones and a halo of pad + 1 either side) exist -- the
    // statistics read block summaries, block centres and window-edge elements assembled from every rank's record, and only
    // the own blocks [own_b0, own_b1) are classified and walked.
    int64_t lim_lo, lim_hi;     // readable positions of `errors` (0, len when the whole array is here)
    int64_t own_b0, own_b1;     // blocks this rank walks (0, number of blocks when the whole array is here)
    const double* cblk;         // [blocks] first element of every block (shard view; else read from `errors`)
    const double* edge;         // [n_analysis][2][TW_TILE] elements of every window's two ragged edges (shard view)
    unsigned long long* lead;   // [n_analysis] maximum of the above-threshold values in front of the first own run start
};

__device__ __forceinline__ void tw_window(const TwArgs& a, int k, const double*& e, int64_t& n) {
    const int64_t w0 = (int64_t)(a.k0 + k) * a.step;
    const int64_t w1 = w0 + a.window_size < a.len ? w0 + a.window_size : a.len;
    e = a.errors + w0;
    n = w1 - w0;
}

// mode 0: partial sums of x ; mode 1: partial sums of (x - mean)^2
__global__ void __launch_bounds__(256) tw_partial_kernel(const TwArgs a, int mode) {
    __shared__ double sh[32];
    const int k = blockIdx.y, s = blockIdx.x;
    const double* e;
    int64_t n;
    tw_window(a, k, e, n);
    const int64_t b0 = (int64_t)s * TW_CHUNK, b1 = b0 + TW_CHUNK < n ? b0 + TW_CHUNK : n;
    const double m = mode ? a.mean[k] : 0.0;
    double acc = 0.0;
    for (int64_t i = b0 + threadIdx.x; i < b1; i += blockDim.x) {
        const double d = e[i] - m;
        acc += mode ? d * d : d;
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) a.partial[(size_t)k * a.n_slices + s] = acc;
}

// mode 0: mean ; mode 1: std, threshold, and reset of the event state
__global__ void __launch_bounds__(32) tw_final_kernel(const TwArgs a, int mode) {
    const int k = blockIdx.x, lane = threadIdx.x;
    const double* e;
    int64_t n;
    tw_window(a, k, e, n);
    double acc = 0.0;
    for (int s = lane; s < a.n_slices; s += 32) acc += a.partial[(size_t)k * a.n_slices + s];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (mode == 0) {
        if (lane == 0) a.mean[k] = acc / (double)n;
    } else {
        if (lane == 0) {
            double mean = a.mean[k];
            double sd = sqrt(acc / (double)(n - a.ddof));
            double thr = mean + 4.0 * sd;  // _fixed_threshold, k = 4 (:1098-1114)
            if (a.stats_f32) {
                mean = (double)(float)mean;
                sd = (double)(float)sd;
                thr = (double)__fadd_rn((float)mean, __fmul_rn(4.0f, (float)sd));
            }
            a.stats[k * 4 + 0] = mean;
            a.stats[k * 4 + 1] = sd;
            a.stats[k * 4 + 2] = thr;
            a.cnt[k * 2] = 0;
            a.cnt[k * 2 + 1] = 0;
            a.below[k] = 0ull;
        }
        for (int r = lane; r < a.max_runs; r += 32) a.rmax[(size_t)k * a.max_runs + r] = 0ull;
    }
}

// any flag bit set in [lo, hi] of the packed window
__device__ __forceinline__ bool any_bits(const unsigned* w, int lo, int hi) {
    const int w0 = lo >> 5, w1 = hi >> 5;
    const unsigned m0 = 0xffffffffu << (lo & 31), m1 = 0xffffffffu >> (31 - (hi & 31));
    if (w0 == w1) return (w[w0] & m0 & m1) != 0u;
    unsigned acc = (w[w0] & m0) | (w[w1] & m1);
    for (int i = w0 + 1; i < w1; ++i) acc |= w[i];
    return acc != 0u;
}

// One tile [t0, t0+len) (window-relative, len <= TW_TILE) of window k: flags, dilation, run starts / ends, max outside runs.
__device__ __forceinline__ void tw_events_tile(const TwArgs& a, int k, int64_t t0, int len, unsigned* s_bits, unsigned char* s_dil,
                                               unsigned long long* s_below) {
    const int tid = threadIdx.x;
    const double* e;
    int64_t n;
    tw_window(a, k, e, n);
    const double thr = a.stats[k * 4 + 2];
    const int pad = a.pad;
    const int region = TW_TILE + 2 * pad + 2;       // positions t0-pad-1 .. t0+TILE+pad
    const int64_t r0 = t0 - pad - 1;
    if (tid == 0) *s_below = 0ull;
    bool any_flag = false;
    for (int f = tid; f < ((region + 31) & ~31); f += TW_TILE) {
        const int64_t q = r0 + f;
        const int64_t gq = (int64_t)(a.k0 + k) * a.step + q;  // global position: only [lim_lo, lim_hi) is readable
        const bool flag = f < region && q >= 0 && q < n && gq >= a.lim_lo && gq < a.lim_hi && e[q] > thr;
        const unsigned word = __ballot_sync(0xffffffffu, flag);
        any_flag |= word != 0u;
        if ((tid & 31) == 0) s_bits[f >> 5] = word;
    }
    // Almost every tile of almost every window has nothing above the threshold in reach: then no position is in a run
    // and the tile only contributes its maximum to max_below.  (The barrier also publishes s_bits and s_below.)
    const bool quiet = !__syncthreads_or(any_flag);
    const int64_t i = t0 + tid;
    const bool mine = tid < len && i < n;
    unsigned long long below_key = 0ull;
    if (quiet) {
        below_key = mine ? dkey(e[i]) : 0ull;
    } else {
        // dilation of positions t0-1 .. t0+TILE: bit range [j, j+2pad] of the packed window
        for (int j = tid; j < TW_TILE + 2; j += TW_TILE) {
            const int64_t p = t0 - 1 + j;
            s_dil[j] = (p >= 0 && p < n && any_bits(s_bits, j, j + 2 * pad)) ? 1 : 0;
        }
        __syncthreads();
        if (mine) {
            const bool dil = s_dil[tid + 1] != 0;
            if (dil) {
                if (!s_dil[tid]) {  // run start (shift(1).fillna(False), :1151-1160)
                    const int slot = atomicAdd(&a.cnt[k * 2], 1);
                    if (slot < a.max_runs) a.starts[(size_t)k * a.max_runs + slot] = i;
                }
                if (!s_dil[tid + 2]) {  // run end (the last element closes an open run, :1163-1164)
                    const int slot = atomicAdd(&a.cnt[k * 2 + 1], 1);
                    if (slot < a.max_runs) a.ends[(size_t)k * a.max_runs + slot] = i;
                }
            } else {
                below_key = dkey(e[i]);
            }
        }
    }
    // block max of the not-in-any-run values -> one atomic per CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long v = __shfl_xor_sync(0xffffffffu, below_key, o);
        below_key = v > below_key ? v : below_key;
    }
    if ((tid & 31) == 0 && below_key) atomicMax(s_below, below_key);
    __syncthreads();
    if (tid == 0 && *s_below) atomicMax(&a.below[k], *s_below);
    __syncthreads();  // the shared state is reused by the caller's next tile
}

// exhaustive version: every tile of every window
__global__ void __launch_bounds__(TW_TILE) tw_events_kernel(const TwArgs a) {
    constexpr int REGION = TW_TILE + 2 * TW_MAXPAD + 2;
    __shared__ unsigned s_bits[(REGION + 31) / 32 + 1];
    __shared__ unsigned char s_dil[TW_TILE + 2];
    __shared__ unsigned long long s_below;
    const int k = blockIdx.y;
    const double* e;
    int64_t n;
    tw_window(a, k, e, n);
    const int64_t t0 = (int64_t)blockIdx.x * TW_TILE;
    if (t0 >= n) return;
    tw_events_tile(a, k, t0, TW_TILE, s_bits, s_dil, &s_below);
}

// the run a position belongs to: the one with the largest start <= the position
__device__ __forceinline__ void tw_runmax_elem(const TwArgs& a, int k, const double* e, int64_t i) {
    const double v = e[i];
    if (!(v > a.stats[k * 4 + 2])) return;
    const int R = a.cnt[k * 2] < a.max_runs ? a.cnt[k * 2] : a.max_runs;
    const long long* st = a.starts + (size_t)k * a.max_runs;
    long long best = -1;
    int arg = -1;
    for (int r = 0; r < R; ++r) {
        const long long s = st[r];
        if (s <= i && s > best) {
            best = s;
            arg = r;
        }
    }
    if (arg >= 0) atomicMax(&a.rmax[(size_t)k * a.max_runs + arg], dkey(v));
    else if (a.lead) atomicMax(&a.lead[k], dkey(v));  // its run started on an earlier rank
}

// every above-threshold value goes to the run whose start is the largest start <= its position
__global__ void __launch_bounds__(TW_TILE) tw_runmax_kernel(const TwArgs a) {
    const int k = blockIdx.y;
    const double* e;
    int64_t n;
    tw_window(a, k, e, n);
    const int64_t i = (int64_t)blockIdx.x * TW_TILE + threadIdx.x;
    if (i >= n) return;
    tw_runmax_elem(a, k, e, i);
}

// ---------------------------------------------------------------------------------------------------------
// The product path reads the array ONCE.  The analysis windows overlap ten-fold and nearly all of their elements are far
// from any anomaly, so per aligned block of TW_TILE elements the kernel below keeps sum(x-c_b), sum((x-c_b)^2) and the maximum,
// centred on the block's own first element c_b; a window's statistics are the sums of its inner blocks plus its two ragged
// edges, and a block whose own and neighbouring maxima are all below the window's threshold cannot touch a run: it
// contributes its maximum to max_below and nothing else.  Only the remaining (window, block) pairs -- the blocks near
// anomalies and the window edges -- go through the element-wise tile code above, from a work list.
// The local centres are what makes flat stretches come out right: there every difference is exactly 0 (or a few ulps), the
// window mean is its first element plus an exactly-small correction, and the variance is the sum of squared deviations from
// THAT mean with no cancellation -- a constant window gets mean = the value, std = 0, threshold = the value, nothing above it,
// like numpy's two-pass mean / std do (a far-away centre left a residue of ~1e-16 in the mean with the variance clamped to 0,
// i.e. a threshold below the constant and one run spanning the whole window).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tw_blocks_body(const TwArgs& a, int64_t b, double* sh, unsigned long long* shm) {
    const int64_t i0 = b * TW_TILE, i1 = i0 + TW_TILE < a.len ? i0 + TW_TILE : a.len;
    const double c = a.errors[i0];
    double s1 = 0.0, s2 = 0.0;
    unsigned long long m = 0ull;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += RB) {
        const double x = a.errors[i], d = x - c;
        s1 += d;
        s2 += d * d;
        const unsigned long long key = dkey(x);
        m = key > m ? key : m;
    }
    s1 = block_sum(s1, sh);
    s2 = block_sum(s2, sh);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long v = __shfl_xor_sync(0xffffffffu, m, o);
        m = v > m ? v : m;
    }
    if ((threadIdx.x & 31) == 0) shm[threadIdx.x >> 5] = m;
    group_sync();
    if (threadIdx.x == 0) {
        for (int w = 1; w < RB / 32; ++w) m = shm[w] > m ? shm[w] : m;
        a.bsum1[b] = s1;
        a.bsum2[b] = s2;
        a.bmax[b] = m;
        if (a.cblk) const_cast<double*>(a.cblk)[b] = c;
    }
    group_sync();  // shm is reused by the caller's next block
}
__global__ void __launch_bounds__(RB) tw_blocks_kernel(const TwArgs a) {
    __shared__ double sh[32];
    __shared__ unsigned long long shm[RB / 32];
    tw_blocks_body(a, blockIdx.x, sh, shm);
}

// one CTA per window: statistics, reset of the event state, classification of the window's blocks
// SHARD: the array is not here; block centres and edge elements come from the assembled records (same values, same arithmetic)
template <bool SHARD>
__device__ __forceinline__ void tw_window_body(const TwArgs& a, int k, double* sh, unsigned long long* s_quiet_p) {
    unsigned long long& s_quiet = *s_quiet_p;
    const int tid = threadIdx.x;
    const int64_t w0 = (int64_t)(a.k0 + k) * a.step, w1 = w0 + a.window_size < a.len ? w0 + a.window_size : a.len, n = w1 - w0;
    const int64_t bf0 = (w0 + TW_TILE - 1) / TW_TILE, bf1 = w1 / TW_TILE;  // blocks [bf0, bf1) lie fully inside
    const bool blocks = bf0 < bf1;
    const int64_t e0 = blocks ? bf0 * TW_TILE : w1, e1 = blocks ? bf1 * TW_TILE : w1;  // ragged edges [w0, e0) and [e1, w1)
    const double* eL = SHARD ? a.edge + (size_t)k * 2 * TW_TILE - w0 : a.errors;             // element i of the left edge: eL[i]
    const double* eR = SHARD ? a.edge + ((size_t)k * 2 + 1) * TW_TILE - e1 : a.errors;       // element i of the right edge: eR[i]
    auto centre = [&](int64_t b) { return SHARD ? a.cblk[b] : a.errors[b * TW_TILE]; };
    const double c0 = SHARD ? (w0 < e0 ? eL[w0] : a.cblk[bf0]) : a.errors[w0];  // the window's own centre: its first element
    // mean = c0 + sum(x - c0) / n, with sum over a block = TW_TILE (c_b - c0) + sum(x - c_b)
    double s1 = 0.0;
    if (blocks)
        for (int64_t b = bf0 + tid; b < bf1; b += RB) s1 += (double)TW_TILE * (centre(b) - c0) + a.bsum1[b];
    for (int64_t i = w0 + tid; i < e0; i += RB) s1 += eL[i] - c0;
    for (int64_t i = e1 + tid; i < w1; i += RB) s1 += eR[i] - c0;
    s1 = block_sum(s1, sh);
    const double dm = s1 / (double)n;  // mean - c0
    // sum((x - mean)^2); over a block, with e = c_b - mean: sum((x-c_b)^2) + 2 e sum(x-c_b) + TW_TILE e^2
    double s2 = 0.0;
    if (blocks)
        for (int64_t b = bf0 + tid; b < bf1; b += RB) {
            const double e = (centre(b) - c0) - dm;
            s2 += a.bsum2[b] + 2.0 * e * a.bsum1[b] + (double)TW_TILE * e * e;
        }
    for (int64_t i = w0 + tid; i < e0; i += RB) {
        const double d = (eL[i] - c0) - dm;
        s2 += d * d;
    }
    for (int64_t i = e1 + tid; i < w1; i += RB) {
        const double d = (eR[i] - c0) - dm;
        s2 += d * d;
    }
    s2 = block_sum(s2, sh);
    double var = s2 / (double)(n - a.ddof);
    var = var > 0.0 ? var : 0.0;
    double mean = c0 + dm, sd = sqrt(var), thr = mean + 4.0 * sd;  // _fixed_threshold, k = 4 (:1098-1114)
    if (a.stats_f32) {  // torch fp32 tensor: errors.mean(), errors.std() and mean + 4 * std are fp32 values
        mean = (double)(float)mean;
        sd = (double)(float)sd;
        thr = (double)__fadd_rn((float)mean, __fmul_rn(4.0f, (float)sd));
    }
    if (tid == 0) {
        a.stats[k * 4 + 0] = mean;
        a.stats[k * 4 + 1] = sd;
        a.stats[k * 4 + 2] = thr;
        a.cnt[k * 2] = 0;
        a.cnt[k * 2 + 1] = 0;
        if (a.lead) a.lead[k] = 0ull;
        s_quiet = 0ull;
    }
    for (int r = tid; r < a.max_runs; r += RB) a.rmax[(size_t)k * a.max_runs + r] = 0ull;
    group_sync();
    // blocks overlapping the window: quiet ones give their maximum, the others go to the work list
    const int64_t nb = (a.len + TW_TILE - 1) / TW_TILE;
    const int64_t b_lo = w0 / TW_TILE, b_hi = (w1 - 1) / TW_TILE;
    unsigned long long quiet = 0ull;
    for (int64_t b = b_lo + tid; b <= b_hi; b += RB) {
        if (b < a.own_b0 || b >= a.own_b1) continue;  // another rank's block
        const bool inner = b >= bf0 && b < bf1;
        bool hot = !inner || a.bmax[b] > dkey(thr);
        if (b > 0) hot |= a.bmax[b - 1] > dkey(thr);
        if (b + 1 < nb) hot |= a.bmax[b + 1] > dkey(thr);
        if (hot) {
            const int slot = atomicAdd(a.work_cnt, 1);
            a.work[slot] = make_longlong2((long long)k, (long long)b);
        } else {
            quiet = a.bmax[b] > quiet ? a.bmax[b] : quiet;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long v = __shfl_xor_sync(0xffffffffu, quiet, o);
        quiet = v > quiet ? v : quiet;
    }
    if ((tid & 31) == 0 && quiet) atomicMax(&s_quiet, quiet);
    group_sync();
    if (tid == 0) a.below[k] = s_quiet;
    group_sync();  // s_quiet is reused by the caller's next window
}
__global__ void __launch_bounds__(RB) tw_window_kernel(const TwArgs a) {
    __shared__ double sh[32];
    __shared__ unsigned long long s_quiet;
    tw_window_body<false>(a, blockIdx.x, sh, &s_quiet);
}

// the work list's (window, block) pairs through the element-wise tile code
__global__ void __launch_bounds__(TW_TILE) tw_events_work_kernel(const TwArgs a) {
    constexpr int REGION = TW_TILE + 2 * TW_MAXPAD + 2;
    __shared__ unsigned s_bits[(REGION + 31) / 32 + 1];
    __shared__ unsigned char s_dil[TW_TILE + 2];
    __shared__ unsigned long long s_below;
    const int total = *a.work_cnt;
    for (int j = blockIdx.x; j < total; j += gridDim.x) {
        const longlong2 kb = a.work[j];
        const int k = (int)kb.x;
        const int64_t w0 = (int64_t)(a.k0 + k) * a.step, w1 = w0 + a.window_size < a.len ? w0 + a.window_size : a.len;
        const int64_t g0 = kb.y * TW_TILE > w0 ? kb.y * TW_TILE : w0, g1 = (kb.y + 1) * TW_TILE < w1 ? (kb.y + 1) * TW_TILE : w1;
        tw_events_tile(a, k, g0 - w0, (int)(g1 - g0), s_bits, s_dil, &s_below);
    }
}

__global__ void __launch_bounds__(TW_TILE) tw_runmax_work_kernel(const TwArgs a) {
    const int total = *a.work_cnt;
    for (int j = blockIdx.x; j < total; j += gridDim.x) {
        const longlong2 kb = a.work[j];
        const int k = (int)kb.x;
        const int64_t w0 = (int64_t)(a.k0 + k) * a.step, w1 = w0 + a.window_size < a.len ? w0 + a.window_size : a.len;
        const int64_t g = kb.y * TW_TILE + threadIdx.x;
        if (g >= w0 && g < w1) tw_runmax_elem(a, k, a.errors + w0, g - w0);
    }
}

// per window: order the runs by start, pair the r-th start with the r-th end, emit (start, end, max)
__device__ __forceinline__ void tw_emit_body(const TwArgs& a, int k, int tid, int nthreads) {
    const int total = a.cnt[k * 2];
    const int R = total < a.max_runs ? total : a.max_runs;
    const long long* st = a.starts + (size_t)k * a.max_runs;
    const long long* en = a.ends + (size_t)k * a.max_runs;
    double* out = a.runs + (size_t)k * a.max_runs * 3;
    for (int r = tid; r < R; r += nthreads) {
        const long long s = st[r], t = en[r];
        int rank_s = 0, rank_e = 0;
        for (int q = 0; q < R; ++q) {
            rank_s += st[q] < s;
            rank_e += en[q] < t;
        }
        out[rank_s * 3 + 0] = (double)s;
        out[rank_s * 3 + 2] = dunkey(a.rmax[(size_t)k * a.max_runs + r]);
        out[rank_e * 3 + 1] = (double)t;
    }
    if (tid == 0) {
        a.n_runs[k] = total;
        const unsigned long long b = a.below[k];
        a.stats[k * 4 + 3] = b ? dunkey(b) : 0.0;  // `above.all()` -> max_below = 0 (:1154-1155)
    }
}
__global__ void __launch_bounds__(256) tw_emit_kernel(const TwArgs a) { tw_emit_body(a, blockIdx.x, threadIdx.x, blockDim.x); }

// Short arrays: the five phases above in ONE cooperative launch, a grid-wide barrier between them (a signal of a few thousand
// positions is launch-bound: five launches cost more than their work).  The same device code, phase by phase, so the results are
// those of the separate launches.  The statistics phases use the first 256 threads of each CTA (the shape their reductions
// were written for), the element-wise phases all 1024.
__global__ void __launch_bounds__(TW_TILE) tw_fused_kernel(const TwArgs a, int nb) {
    constexpr int REGION = TW_TILE + 2 * TW_MAXPAD + 2;
    __shared__ unsigned s_bits[(REGION + 31) / 32 + 1];
    __shared__ unsigned char s_dil[TW_TILE + 2];
    __shared__ unsigned long long s_below, s_quiet, shm[RB / 32];
    __shared__ double sh[32];
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int tid = threadIdx.x;
    if (blockIdx.x == 0 && tid == 0) *a.work_cnt = 0;
    if (tid < RB)
        for (int64_t b = blockIdx.x; b < nb; b += gridDim.x) tw_blocks_body(a, b, sh, shm);
    grid.sync();
    if (tid < RB)
        for (int k = blockIdx.x; k < a.n_analysis; k += gridDim.x) tw_window_body<false>(a, k, sh, &s_quiet);
    grid.sync();
    const int total = *a.work_cnt;
    for (int j = blockIdx.x; j < total; j += gridDim.x) {
        const longlong2 kb = a.work[j];
        const int k = (int)kb.x;
        const int64_t w0 = (int64_t)(a.k0 + k) * a.step, w1 = w0 + a.window_size < a.len ? w0 + a.window_size : a.len;
        const int64_t g0 = kb.y * TW_TILE > w0 ? kb.y * TW_TILE : w0, g1 = (kb.y + 1) * TW_TILE < w1 ? (kb.y + 1) * TW_TILE : w1;
        tw_events_tile(a, k, g0 - w0, (int)(g1 - g0), s_bits, s_dil, &s_below);
    }
    grid.sync();
    for (int j = blockIdx.x; j < total; j += gridDim.x) {
        const longlong2 kb = a.work[j];
        const int k = (int)kb.x;
        const int64_t w0 = (int64_t)(a.k0 + k) * a.step, w1 = w0 + a.window_size < a.len ? w0 + a.window_size : a.len;
        const int64_t g = kb.y * TW_TILE + tid;
        if (g >= w0 && g < w1) tw_runmax_elem(a, k, a.errors + w0, g - w0);
    }
    grid.sync();
    if (tid < RB)
        for (int k = blockIdx.x; k < a.n_analysis; k += gridDim.x) tw_emit_body(a, k, tid, RB);
}

// ---------------------------------------------------------------------------------------------------------
// find_anomalies on an array sharded over several GPUs by contiguous, block-aligned ranges (hypad_tw_shard_pack / _runs).
// Nothing of the array's total length is gathered: a rank contributes a record -- the summaries (first element, centred sums,
// maximum) of its blocks, the elements of the window edges that fall into its range (at most two partial blocks per analysis
// window) and its first / last pad + 1 values for the neighbours' dilation halo -- and, with everybody's records, computes
// every window's statistics itself (the single-GPU arithmetic on the same values, tw_window_body<true>) and extracts the run
// fragments of ITS positions.  The fragments are gathered and joined on the host.
// ---------------------------------------------------------------------------------------------------------
constexpr int TW_MAXWORLD = 64;
struct TwShardMap {
    int world, rank;
    long long blk_start[TW_MAXWORLD + 1];  // first block of every rank (blocks are dealt out in order)
};

// record: cblk[bmax] | s1[bmax] | s2[bmax] | maxkey[bmax] | edge[n_analysis][2][TW_TILE] | strips[2][hp]
__global__ void __launch_bounds__(RB) tw_shard_edges_kernel(const double* __restrict__ local, int64_t first, int64_t count, int64_t n_total,
                                                            int64_t window_size, int64_t step, int hp, double* __restrict__ edge,
                                                            double* __restrict__ strips) {
    const int k = blockIdx.x, side = blockIdx.y;
    if (k == gridDim.x - 1) {  // the extra row of CTAs copies the halo strips: first hp values left-aligned, last hp right-aligned
        const int64_t m = count < hp ? count : hp;
        for (int j = threadIdx.x; j < m; j += RB) strips[side * hp + (side ? hp - m + j : j)] = side ? local[count - m + j] : local[j];
        return;
    }
    const int64_t w0 = (int64_t)k * step, w1 = w0 + window_size < n_total ? w0 + window_size : n_total;
    const int64_t bf0 = (w0 + TW_TILE - 1) / TW_TILE, bf1 = w1 / TW_TILE;
    const int64_t g0 = side ? bf1 * TW_TILE : w0, g1 = side ? w1 : bf0 * TW_TILE;  // the edge [g0, g1) lies inside one block
    if (g0 >= g1 || g0 < first || g0 >= first + count) return;                     // no edge, or another rank's
    double* dst = edge + ((size_t)k * 2 + side) * TW_TILE;
    for (int64_t j = threadIdx.x; j < g1 - g0; j += RB) dst[j] = local[g0 - first + j];
}

// every rank's record -> dense arrays over ALL blocks / windows (what tw_window_body<true> reads)
__global__ void tw_shard_assemble_kernel(const double* __restrict__ records, size_t rec_len, TwShardMap map, int bmax_per_rank, int64_t nb_total,
                                         int n_analysis, int64_t window_size, int64_t step, int64_t n_total, double* __restrict__ cblk,
                                         double* __restrict__ bsum1, double* __restrict__ bsum2, unsigned long long* __restrict__ bmaxk,
                                         double* __restrict__ edge) {
    auto owner = [&](int64_t b) {
        int r = 0;
        while (r + 1 < map.world && b >= map.blk_start[r + 1]) ++r;
        return r;
    };
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t b = gid; b < nb_total; b += stride) {
        const int r = owner(b);
        const double* rec = records + (size_t)r * rec_len;
        const int64_t lb = b - map.blk_start[r];
        cblk[b] = rec[lb];
        bsum1[b] = rec[bmax_per_rank + lb];
        bsum2[b] = rec[2 * (size_t)bmax_per_rank + lb];
        bmaxk[b] = (unsigned long long)__double_as_longlong(rec[3 * (size_t)bmax_per_rank + lb]);
    }
    const int64_t total = (int64_t)n_analysis * 2 * TW_TILE;
    for (int64_t e = gid; e < total; e += stride) {
        const int64_t ks = e / TW_TILE;
        const int k = (int)(ks >> 1), side = (int)(ks & 1);
        const int64_t w0 = (int64_t)k * step, w1 = w0 + window_size < n_total ? w0 + window_size : n_total;
        const int64_t bf0 = (w0 + TW_TILE - 1) / TW_TILE, bf1 = w1 / TW_TILE;
        const int64_t g0 = side ? bf1 * TW_TILE : w0, g1 = side ? w1 : bf0 * TW_TILE;
        double v = 0.0;
        if (g0 < g1) v = records[(size_t)owner(g0 / TW_TILE) * rec_len + 4 * (size_t)bmax_per_rank + e];
        edge[e] = v;
    }
}

__global__ void __launch_bounds__(RB) tw_shard_window_kernel(const TwArgs a) {
    __shared__ double sh[32];
    __shared__ unsigned long long s_quiet;
    tw_window_body<true>(a, blockIdx.x, sh, &s_quiet);
}

// per window: [n_starts, n_ends, lead key, below key, mean, std, threshold, 0] | starts (ascending) | their maxima | ends (ascending)
__global__ void __launch_bounds__(256) tw_shard_emit_kernel(const TwArgs a, double* __restrict__ out) {
    const int k = blockIdx.x;
    const int ns = a.cnt[k * 2], ne = a.cnt[k * 2 + 1];
    const int Rs = ns < a.max_runs ? ns : a.max_runs, Re = ne < a.max_runs ? ne : a.max_runs;
    const long long* st = a.starts + (size_t)k * a.max_runs;
    const long long* en = a.ends + (size_t)k * a.max_runs;
    double* o = out + (size_t)k * (8 + 3 * (size_t)a.max_runs);
    for (int r = threadIdx.x; r < Rs; r += blockDim.x) {
        const long long s = st[r];
        int rank_s = 0;
        for (int q = 0; q < Rs; ++q) rank_s += st[q] < s;
        o[8 + rank_s] = (double)s;
        const unsigned long long m = a.rmax[(size_t)k * a.max_runs + r];
        o[8 + a.max_runs + rank_s] = m ? dunkey(m) : -1.0 / 0.0;  // a start whose flagged values all lie on the next rank: -inf
    }
    for (int r = threadIdx.x; r < Re; r += blockDim.x) {
        const long long t = en[r];
        int rank_e = 0;
        for (int q = 0; q < Re; ++q) rank_e += en[q] < t;
        o[8 + 2 * a.max_runs + rank_e] = (double)t;
    }
    if (threadIdx.x == 0) {
        o[0] = (double)ns;
        o[1] = (double)ne;
        o[2] = __longlong_as_double((long long)a.lead[k]);
        o[3] = __longlong_as_double((long long)a.below[k]);
        o[4] = a.stats[k * 4 + 0];
        o[5] = a.stats[k * 4 + 1];
        o[6] = a.stats[k * 4 + 2];
        o[7] = 0.0;
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

static size_t rolling_ws_bytes(int64_t len) {
    return 3 * align256((size_t)len * 8) + 3 * align256((size_t)ceil_div(len, SCAN_CHUNK) * 8);
}
// x covers global positions [ext0, ext0 + ext_len) of n_total; out[j] = smoothed value at position p0 + j, j < count.
// zscore: smooth the critic z-score of x (scalars of critic_stats.cu) instead of x.
static int rolling_mean(hypad_ctx* ctx, const double* x, int64_t ext_len, int64_t ext0, int64_t n_total, int64_t p0, int64_t count,
                        int64_t window, int64_t min_periods, bool zscore, double* out, char* ws, cudaStream_t stream) {
    // ws: pre_hi[len] | pre_lo[len] | lc[len] | tot_hi[nchunks] | tot_lo[nchunks] | tot_lc[nchunks]
    const int64_t nchunks = ceil_div(ext_len, SCAN_CHUNK);
    const size_t la = align256((size_t)ext_len * 8), ca = align256((size_t)nchunks * 8);
    ScanBufs sb;
    sb.pre_hi = (double*)ws;
    sb.pre_lo = (double*)(ws + la);
    sb.lc = (long long*)(ws + 2 * la);
    sb.tot_hi = (double*)(ws + 3 * la);
    sb.tot_lo = (double*)(ws + 3 * la + ca);
    sb.tot_lc = (long long*)(ws + 3 * la + 2 * ca);
    const double* scal = zscore ? fin_scalars(ctx) : nullptr;
    if (zscore) scan_local_kernel<true><<<(unsigned)nchunks, 256, 0, stream>>>(x, ext_len, sb, scal);
    else scan_local_kernel<false><<<(unsigned)nchunks, 256, 0, stream>>>(x, ext_len, sb, scal);
    HYPAD_LAUNCH_CHECK();
    scan_totals_kernel<<<1, 32, 0, stream>>>(sb, nchunks);
    HYPAD_LAUNCH_CHECK();
    if (zscore) rolling_mean_kernel<true><<<ew_grid(count), 256, 0, stream>>>(x, sb, scal, ext0, n_total, p0, count, window, min_periods, out);
    else rolling_mean_kernel<false><<<ew_grid(count), 256, 0, stream>>>(x, sb, scal, ext0, n_total, p0, count, window, min_periods, out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

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
    return rolling_mean(ctx, x, len, 0, len, 0, len, window, min_periods, false, out, (char*)ctx->workspace, (cudaStream_t)stream);
}

int hypad_critic_smooth_shard(hypad_ctx* ctx, const double* kmax_ext, int64_t ext_len, int64_t ext0, int64_t n_total, int64_t p0,
                              int64_t count, int64_t smooth_window, double* out, void* stream) {
    HYPAD_REQUIRE(ctx && ctx->fin_state && kmax_ext && out, "hypad_critic_smooth_shard: NULL argument (hypad_stats_* first)");
    HYPAD_REQUIRE(ext_len >= 1 && ext0 >= 0 && ext0 + ext_len <= n_total && p0 >= ext0 && count >= 0 && p0 + count <= ext0 + ext_len,
                  "hypad_critic_smooth_shard: the slice [%lld, %lld) does not hold the positions [%lld, %lld)", (long long)ext0,
                  (long long)(ext0 + ext_len), (long long)p0, (long long)(p0 + count));
    if (count == 0) return HYPAD_OK;
    if (smooth_window > 0) {
        const int64_t need_lo = p0 - smooth_window / 2 > 0 ? p0 - smooth_window / 2 : 0;
        const int64_t hi = p0 + count - 1 + (smooth_window - 1) / 2 + 1, need_hi = hi < n_total ? hi : n_total;
        HYPAD_REQUIRE(ext0 <= need_lo && ext0 + ext_len >= need_hi, "hypad_critic_smooth_shard: the slice lacks the smoothing halo "
                      "(needs [%lld, %lld), holds [%lld, %lld))", (long long)need_lo, (long long)need_hi, (long long)ext0,
                      (long long)(ext0 + ext_len));
    }
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_workspace(ctx, rolling_ws_bytes(ext_len));
    if (rc != HYPAD_OK) return rc;
    return rolling_mean(ctx, kmax_ext, ext_len, ext0, n_total, p0, count, smooth_window, smooth_window / 2, true, out,
                        (char*)ctx->workspace, (cudaStream_t)stream);
}

int hypad_rolling_mean_shard(hypad_ctx* ctx, const double* x_ext, int64_t ext_len, int64_t ext0, int64_t n_total, int64_t p0,
                             int64_t count, int64_t window, int64_t min_periods, double* out, void* stream) {
    HYPAD_REQUIRE(ctx && x_ext && out, "hypad_rolling_mean_shard: NULL argument");
    HYPAD_REQUIRE(ext_len >= 1 && ext0 >= 0 && ext0 + ext_len <= n_total && p0 >= ext0 && count >= 0 && p0 + count <= ext0 + ext_len,
                  "hypad_rolling_mean_shard: the slice [%lld, %lld) does not hold the positions [%lld, %lld)", (long long)ext0,
                  (long long)(ext0 + ext_len), (long long)p0, (long long)(p0 + count));
    if (count == 0) return HYPAD_OK;
    if (window > 0) {
        const int64_t need_lo = p0 - window / 2 > 0 ? p0 - window / 2 : 0;
        const int64_t hi = p0 + count - 1 + (window - 1) / 2 + 1, need_hi = hi < n_total ? hi : n_total;
        HYPAD_REQUIRE(ext0 <= need_lo && ext0 + ext_len >= need_hi, "hypad_rolling_mean_shard: the slice lacks the smoothing halo "
                      "(needs [%lld, %lld), holds [%lld, %lld))", (long long)need_lo, (long long)need_hi, (long long)ext0,
                      (long long)(ext0 + ext_len));
    }
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_workspace(ctx, rolling_ws_bytes(ext_len));
    if (rc != HYPAD_OK) return rc;
    return rolling_mean(ctx, x_ext, ext_len, ext0, n_total, p0, count, window, min_periods, false, out, (char*)ctx->workspace,
                        (cudaStream_t)stream);
}

int hypad_critic_scores(hypad_ctx* ctx, const double* kmax, int64_t len, int64_t smooth_window, int keys_f32, double* out,
                        void* stream) {
    HYPAD_REQUIRE(ctx && kmax && out, "hypad_critic_scores: NULL argument");
    HYPAD_REQUIRE(len >= 1, "hypad_critic_scores: len < 1");
    // the one-rank chain of the staged statistics (critic_stats.cu)
    int rc = hypad_stats_select_begin(ctx, len, keys_f32, stream);
    if (rc != HYPAD_OK) return rc;
    uint32_t* hist = (uint32_t*)fin_local_record(ctx);
    for (int p = 0; p < hypad_stats_select_passes(keys_f32); ++p) {
        if ((rc = hypad_stats_select_hist(ctx, kmax, len, p, hist, stream)) != HYPAD_OK) return rc;
        if ((rc = hypad_stats_select_pick(ctx, hist, 1, p, stream)) != HYPAD_OK) return rc;
    }
    double* rec = fin_local_record(ctx);
    if ((rc = hypad_stats_moments_partial(ctx, kmax, 0, len, 1, rec, stream)) != HYPAD_OK) return rc;
    if ((rc = hypad_stats_moments_final(ctx, rec, 1, len, 1, 0, stream)) != HYPAD_OK) return rc;
    return hypad_critic_smooth_shard(ctx, kmax, len, 0, len, 0, len, smooth_window, out, stream);
}

int hypad_critic_zscore_smooth(hypad_ctx* ctx, const double* kmax, int64_t len, int64_t smooth_window, double* out,
                               void* stream) {
    return hypad_critic_scores(ctx, kmax, len, smooth_window, 0, out, stream);  // arbitrary doubles: 64-bit keys
}

int hypad_zscore_clip_apply(hypad_ctx* ctx, const void* x, int x_is_f32, int64_t len, double* out, void* stream_) {
    HYPAD_REQUIRE(ctx && ctx->fin_state && out && (x || len == 0) && len >= 0, "hypad_zscore_clip_apply: bad argument");
    if (len == 0) return HYPAD_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    if (x_is_f32) zscore_clip_kernel<float><<<ew_grid(len), 256, 0, stream>>>((const float*)x, len, fin_scalars(ctx), out);
    else zscore_clip_kernel<double><<<ew_grid(len), 256, 0, stream>>>((const double*)x, len, fin_scalars(ctx), out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_zscore_clip(hypad_ctx* ctx, const void* x, int x_is_f32, int64_t len, double* out, void* stream) {
    HYPAD_REQUIRE(ctx && x && out, "hypad_zscore_clip: NULL argument");
    HYPAD_REQUIRE(len >= 1, "hypad_zscore_clip: len < 1");
    int rc = ensure_fin_state(ctx);
    if (rc != HYPAD_OK) return rc;
    double* rec = fin_local_record(ctx);
    if ((rc = hypad_stats_moments_partial(ctx, x, x_is_f32, len, 0, rec, stream)) != HYPAD_OK) return rc;
    if ((rc = hypad_stats_moments_final(ctx, rec, 1, len, 0, 0, stream)) != HYPAD_OK) return rc;
    return hypad_zscore_clip_apply(ctx, x, x_is_f32, len, out, stream);
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

static int threshold_windows_impl(hypad_ctx* ctx, const double* errors, int64_t len, int64_t window_size, int64_t step,
                                  int64_t first_window, int64_t n_analysis, int ddof, int anomaly_padding, double* stats,
                                  double* runs, int32_t* n_runs, int max_runs, cudaStream_t stream, bool exhaustive) {
    HYPAD_REQUIRE(ctx && errors && stats && runs && n_runs, "hypad_threshold_windows: NULL argument");
    HYPAD_REQUIRE(len >= 1 && window_size >= 1 && step >= 1 && n_analysis >= 1 && max_runs >= 1, "hypad_threshold_windows: bad shape");
    HYPAD_REQUIRE(n_analysis <= 65535, "hypad_threshold_windows: more than 65535 analysis windows");
    HYPAD_REQUIRE(first_window >= 0 && (first_window + n_analysis - 1) * step < len, "hypad_threshold_windows: last window starts beyond the data");
    HYPAD_REQUIRE(anomaly_padding >= 0 && anomaly_padding <= TW_MAXPAD, "hypad_threshold_windows: padding %d outside 0..%d",
                  anomaly_padding, TW_MAXPAD);
    const int stats_f32 = (ddof & HYPAD_STATS_F32) ? 1 : 0;
    ddof &= ~HYPAD_STATS_F32;
    HYPAD_REQUIRE(ddof == 0 || ddof == 1, "hypad_threshold_windows: ddof must be 0 or 1 (optionally | HYPAD_STATS_F32)");
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t wlen = window_size < len ? window_size : len;
    TwArgs a;
    a.errors = errors; a.len = len; a.window_size = window_size; a.step = step;
    a.lim_lo = 0; a.lim_hi = len; a.own_b0 = 0; a.own_b1 = ceil_div(len, TW_TILE); a.cblk = nullptr; a.edge = nullptr; a.lead = nullptr;
    a.k0 = first_window;
    a.n_analysis = (int)n_analysis; a.ddof = ddof; a.pad = anomaly_padding; a.max_runs = max_runs;
    a.stats_f32 = stats_f32;
    a.n_slices = (int)ceil_div(wlen, TW_CHUNK);
    a.stats = stats; a.runs = runs; a.n_runs = n_runs;
    const size_t na = (size_t)n_analysis, mr = (size_t)max_runs;
    const size_t nb = (size_t)ceil_div(len, TW_TILE), wb = (size_t)ceil_div(wlen, TW_TILE) + 2;  // blocks a window can overlap
    const size_t o_partial = 0, o_mean = o_partial + align256(na * a.n_slices * 8), o_cnt = o_mean + align256(na * 8);
    const size_t o_below = o_cnt + align256(na * 2 * 4), o_starts = o_below + align256(na * 8);
    const size_t o_ends = o_starts + align256(na * mr * 8), o_rmax = o_ends + align256(na * mr * 8);
    const size_t o_b1 = o_rmax + align256(na * mr * 8), o_b2 = o_b1 + align256(nb * 8), o_bm = o_b2 + align256(nb * 8);
    const size_t o_wc = o_bm + align256(nb * 8), o_work = o_wc + 256, o_end = o_work + align256(na * wb * 16);
    int rc = ensure_workspace(ctx, o_end);
    if (rc != HYPAD_OK) return rc;
    char* ws = (char*)ctx->workspace;
    a.partial = (double*)(ws + o_partial); a.mean = (double*)(ws + o_mean); a.cnt = (int*)(ws + o_cnt);
    a.below = (unsigned long long*)(ws + o_below); a.starts = (long long*)(ws + o_starts);
    a.ends = (long long*)(ws + o_ends); a.rmax = (unsigned long long*)(ws + o_rmax);
    a.bsum1 = (double*)(ws + o_b1); a.bsum2 = (double*)(ws + o_b2); a.bmax = (unsigned long long*)(ws + o_bm);
    a.work_cnt = (int*)(ws + o_wc); a.work = (longlong2*)(ws + o_work);
    if (exhaustive) {
        const dim3 gsum((unsigned)a.n_slices, (unsigned)n_analysis), gtile((unsigned)ceil_div(wlen, TW_TILE), (unsigned)n_analysis);
        tw_partial_kernel<<<gsum, 256, 0, stream>>>(a, 0);
        HYPAD_LAUNCH_CHECK();
        tw_final_kernel<<<(unsigned)n_analysis, 32, 0, stream>>>(a, 0);
        HYPAD_LAUNCH_CHECK();
        tw_partial_kernel<<<gsum, 256, 0, stream>>>(a, 1);
        HYPAD_LAUNCH_CHECK();
        tw_final_kernel<<<(unsigned)n_analysis, 32, 0, stream>>>(a, 1);
        HYPAD_LAUNCH_CHECK();
        tw_events_kernel<<<gtile, TW_TILE, 0, stream>>>(a);
        HYPAD_LAUNCH_CHECK();
        tw_runmax_kernel<<<gtile, TW_TILE, 0, stream>>>(a);
        HYPAD_LAUNCH_CHECK();
    } else {
        int sms = kNumSMs;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        if ((int64_t)nb <= 2 * (int64_t)sms) {  // short array: one cooperative launch for all five phases
            int nbi = (int)nb;
            int64_t want = (int64_t)nb > n_analysis ? (int64_t)nb : n_analysis;
            const unsigned grid = (unsigned)(want < sms ? (want < 8 ? 8 : want) : sms);
            void* args[] = {(void*)&a, (void*)&nbi};
            HYPAD_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)tw_fused_kernel, dim3(grid), dim3(TW_TILE), args, 0, stream));
            count_launch();
            return HYPAD_OK;
        }
        HYPAD_CUDA_TRY(cudaMemsetAsync(a.work_cnt, 0, sizeof(int), stream));
        tw_blocks_kernel<<<(unsigned)nb, RB, 0, stream>>>(a);
        HYPAD_LAUNCH_CHECK();
        tw_window_kernel<<<(unsigned)n_analysis, RB, 0, stream>>>(a);
        HYPAD_LAUNCH_CHECK();
        tw_events_work_kernel<<<(unsigned)(2 * sms), TW_TILE, 0, stream>>>(a);
        HYPAD_LAUNCH_CHECK();
        tw_runmax_work_kernel<<<(unsigned)(2 * sms), TW_TILE, 0, stream>>>(a);
        HYPAD_LAUNCH_CHECK();
    }
    tw_emit_kernel<<<(unsigned)n_analysis, 256, 0, stream>>>(a);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_threshold_windows(hypad_ctx* ctx, const double* errors, int64_t len, int64_t window_size, int64_t step,
                            int64_t n_analysis, int ddof, int anomaly_padding, double* stats, double* runs,
                            int32_t* n_runs, int max_runs, void* stream_) {
    return threshold_windows_impl(ctx, errors, len, window_size, step, 0, n_analysis, ddof, anomaly_padding, stats, runs, n_runs,
                                  max_runs, (cudaStream_t)stream_, false);
}

int hypad_threshold_windows_range(hypad_ctx* ctx, const double* errors, int64_t len, int64_t window_size, int64_t step,
                                  int64_t first_window, int64_t n_analysis, int ddof, int anomaly_padding, double* stats,
                                  double* runs, int32_t* n_runs, int max_runs, void* stream_) {
    return threshold_windows_impl(ctx, errors, len, window_size, step, first_window, n_analysis, ddof, anomaly_padding, stats, runs,
                                  n_runs, max_runs, (cudaStream_t)stream_, false);
}

int hypad_threshold_windows_exhaustive(hypad_ctx* ctx, const double* errors, int64_t len, int64_t window_size, int64_t step,
                                       int64_t n_analysis, int ddof, int anomaly_padding, double* stats, double* runs,
                                       int32_t* n_runs, int max_runs, void* stream_) {
    return threshold_windows_impl(ctx, errors, len, window_size, step, 0, n_analysis, ddof, anomaly_padding, stats, runs, n_runs,
                                  max_runs, (cudaStream_t)stream_, true);
}

size_t hypad_tw_shard_record_doubles(int64_t blocks_per_rank, int64_t n_analysis, int anomaly_padding) {
    return (size_t)(4 * blocks_per_rank + n_analysis * 2 * TW_TILE + 2 * (anomaly_padding + 1));
}

int hypad_tw_shard_pack(hypad_ctx* ctx, const double* local, int64_t first, int64_t count, int64_t n_total, int64_t window_size,
                        int64_t step, int64_t n_analysis, int anomaly_padding, int64_t blocks_per_rank, double* record, void* stream_) {
    HYPAD_REQUIRE(ctx && record && (local || count == 0), "hypad_tw_shard_pack: NULL argument");
    HYPAD_REQUIRE(first >= 0 && count >= 0 && first + count <= n_total && first % TW_TILE == 0, "hypad_tw_shard_pack: the range must start at "
                  "a multiple of %d positions", TW_TILE);
    HYPAD_REQUIRE(first + count == n_total || count % TW_TILE == 0, "hypad_tw_shard_pack: only the last range may end inside a block");
    HYPAD_REQUIRE(window_size >= 2 * TW_TILE && step >= 1 && n_analysis >= 1 && n_analysis <= 65535, "hypad_tw_shard_pack: analysis windows of "
                  "at least %d positions", 2 * TW_TILE);
    HYPAD_REQUIRE(anomaly_padding >= 0 && anomaly_padding <= TW_MAXPAD && ceil_div(count, TW_TILE) <= blocks_per_rank,
                  "hypad_tw_shard_pack: padding or block count out of range");
    cudaStream_t stream = (cudaStream_t)stream_;
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t rec = hypad_tw_shard_record_doubles(blocks_per_rank, n_analysis, anomaly_padding);
    HYPAD_CUDA_TRY(cudaMemsetAsync(record, 0, rec * 8, stream));
    if (count == 0) return HYPAD_OK;
    TwArgs a;
    memset(&a, 0, sizeof(a));
    a.errors = local; a.len = count;
    a.cblk = record; a.bsum1 = record + blocks_per_rank; a.bsum2 = record + 2 * blocks_per_rank;
    a.bmax = (unsigned long long*)(record + 3 * blocks_per_rank);
    tw_blocks_kernel<<<(unsigned)ceil_div(count, TW_TILE), RB, 0, stream>>>(a);
    HYPAD_LAUNCH_CHECK();
    const int hp = anomaly_padding + 1;
    double* edge = record + 4 * blocks_per_rank;
    tw_shard_edges_kernel<<<dim3((unsigned)n_analysis + 1, 2), RB, 0, stream>>>(local, first, count, n_total, window_size, step, hp, edge,
                                                                              edge + n_analysis * 2 * TW_TILE);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

int hypad_tw_shard_runs(hypad_ctx* ctx, const double* records, int world, int rank, const int64_t* block_start, int64_t blocks_per_rank,
                        const double* ext, int64_t ext0, int64_t ext_len, int64_t n_total, int64_t window_size, int64_t step,
                        int64_t n_analysis, int ddof, int anomaly_padding, int max_runs, double* out, void* stream_) {
    HYPAD_REQUIRE(ctx && records && block_start && ext && out, "hypad_tw_shard_runs: NULL argument");
    HYPAD_REQUIRE(world >= 1 && world <= TW_MAXWORLD && rank >= 0 && rank < world, "hypad_tw_shard_runs: world %d / rank %d", world, rank);
    HYPAD_REQUIRE(window_size >= 2 * TW_TILE && step >= 1 && n_analysis >= 1 && n_analysis <= 65535 && max_runs >= 1,
                  "hypad_tw_shard_runs: bad shape");
    HYPAD_REQUIRE(anomaly_padding >= 0 && anomaly_padding <= TW_MAXPAD, "hypad_tw_shard_runs: padding %d outside 0..%d", anomaly_padding, TW_MAXPAD);
    const int stats_f32 = (ddof & HYPAD_STATS_F32) ? 1 : 0;
    ddof &= ~HYPAD_STATS_F32;
    HYPAD_REQUIRE(ddof == 0 || ddof == 1, "hypad_tw_shard_runs: ddof must be 0 or 1 (optionally | HYPAD_STATS_F32)");
    cudaStream_t stream = (cudaStream_t)stream_;
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    TwShardMap map;
    map.world = world; map.rank = rank;
    for (int r = 0; r <= world; ++r) map.blk_start[r] = block_start[r];
    const int64_t nbt = ceil_div(n_total, TW_TILE);
    HYPAD_REQUIRE(map.blk_start[0] == 0 && map.blk_start[world] == nbt, "hypad_tw_shard_runs: the block ranges do not cover the array");
    const int64_t own0 = map.blk_start[rank] * TW_TILE, own1 = map.blk_start[rank + 1] * TW_TILE < n_total ? map.blk_start[rank + 1] * TW_TILE : n_total;
    const int64_t hp = anomaly_padding + 1;
    HYPAD_REQUIRE(ext0 <= (own0 - hp > 0 ? own0 - hp : 0) && ext0 + ext_len >= (own1 + hp < n_total ? own1 + hp : n_total),
                  "hypad_tw_shard_runs: the slice [%lld, %lld) lacks the dilation halo of the own positions [%lld, %lld)", (long long)ext0,
                  (long long)(ext0 + ext_len), (long long)own0, (long long)own1);
    const size_t na = (size_t)n_analysis, mr = (size_t)max_runs, nb = (size_t)nbt;
    const size_t wb = (size_t)(map.blk_start[rank + 1] - map.blk_start[rank]) + 2;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += align256(bytes); return at; };
    const size_t o_stats = take(na * 4 * 8), o_cnt = take(na * 2 * 4), o_below = take(na * 8), o_lead = take(na * 8);
    const size_t o_starts = take(na * mr * 8), o_ends = take(na * mr * 8), o_rmax = take(na * mr * 8);
    const size_t o_c = take(nb * 8), o_b1 = take(nb * 8), o_b2 = take(nb * 8), o_bm = take(nb * 8), o_edge = take(na * 2 * TW_TILE * 8);
    const size_t o_wc = take(256), o_work = take(na * wb * 16);
    int rc = ensure_workspace(ctx, o);
    if (rc != HYPAD_OK) return rc;
    char* ws = (char*)ctx->workspace;
    TwArgs a;
    memset(&a, 0, sizeof(a));
    a.errors = ext - ext0;  // virtual base: only [lim_lo, lim_hi) is ever read
    a.len = n_total; a.window_size = window_size; a.step = step; a.k0 = 0;
    a.n_analysis = (int)n_analysis; a.ddof = ddof; a.pad = anomaly_padding; a.max_runs = max_runs; a.stats_f32 = stats_f32;
    a.lim_lo = ext0; a.lim_hi = ext0 + ext_len; a.own_b0 = map.blk_start[rank]; a.own_b1 = map.blk_start[rank + 1];
    a.stats = (double*)(ws + o_stats); a.cnt = (int*)(ws + o_cnt); a.below = (unsigned long long*)(ws + o_below);
    a.lead = (unsigned long long*)(ws + o_lead); a.starts = (long long*)(ws + o_starts); a.ends = (long long*)(ws + o_ends);
    a.rmax = (unsigned long long*)(ws + o_rmax); a.cblk = (double*)(ws + o_c); a.bsum1 = (double*)(ws + o_b1); a.bsum2 = (double*)(ws + o_b2);
    a.bmax = (unsigned long long*)(ws + o_bm); a.edge = (double*)(ws + o_edge); a.work_cnt = (int*)(ws + o_wc); a.work = (longlong2*)(ws + o_work);
    const size_t rec_len = hypad_tw_shard_record_doubles(blocks_per_rank, n_analysis, anomaly_padding);
    int sms = kNumSMs;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    HYPAD_CUDA_TRY(cudaMemsetAsync(a.work_cnt, 0, sizeof(int), stream));
    tw_shard_assemble_kernel<<<(unsigned)(2 * sms), 256, 0, stream>>>(records, rec_len, map, (int)blocks_per_rank, nbt, (int)n_analysis, window_size,
                                                                       step, n_total, (double*)a.cblk, a.bsum1, a.bsum2, a.bmax, (double*)a.edge);
    HYPAD_LAUNCH_CHECK();
    tw_shard_window_kernel<<<(unsigned)n_analysis, RB, 0, stream>>>(a);
    HYPAD_LAUNCH_CHECK();
    tw_events_work_kernel<<<(unsigned)(2 * sms), TW_TILE, 0, stream>>>(a);
    HYPAD_LAUNCH_CHECK();
    tw_runmax_work_kernel<<<(unsigned)(2 * sms), TW_TILE, 0, stream>>>(a);
    HYPAD_LAUNCH_CHECK();
    tw_shard_emit_kernel<<<(unsigned)n_analysis, 256, 0, stream>>>(a, out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

}  // extern "C"
