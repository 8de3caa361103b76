// Host tail of find_anomalies on the few runs per analysis window the device kernels extracted: prune (_prune_anomalies,
// utils/anomaly_detection_utils.py:1203-1237), score (_compute_scores :1240-1269) and merge (_merge_sequences :1272-1313).
// Plain C++ on the host -- the reference does this part on the host too, on a handful of rows per window; synthetic noise
// (BASELINE config 4 at 8M rows) produces tens of thousands of runs, where an interpreter loop costs more than the kernels.
// Arithmetic follows numpy / pandas to the bit: descending stable sort with NaN last, float64 (or float32) scores,
// np.average as (v * w).sum() / w.sum() with numpy's pairwise summation order.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace hypad {

// numpy's pairwise_sum (umath loops): plain loop below 8, eight interleaved accumulators up to 128, halves above
static double np_pairwise_sum_host(const double* a, int64_t n) {
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int64_t i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum_host(a, n2) + np_pairwise_sum_host(a + n2, n - n2);
}

struct Row {
    double m, s, e;
};
struct Seq {
    double s, e, score;
};

}  // namespace hypad

using namespace hypad;

// prune + score + merge of one array's windows; appends (start, end, score) triples to `merged`
static int intervals_core(const double* stats, const double* runs, const int32_t* n_runs, int64_t count, int64_t max_runs, int64_t step,
                          double min_percent, int f32, std::vector<double>& merged) {
    std::vector<Seq> seqs;
    std::vector<Row> rows;
    for (int64_t k = 0; k < count; ++k) {
        const double mean = stats[k * 4 + 0], sd = stats[k * 4 + 1], thr = stats[k * 4 + 2], max_below = stats[k * 4 + 3];
        const int64_t nr = n_runs[k];
        HYPAD_REQUIRE(nr >= 0 && nr <= max_runs, "hypad_intervals_from_runs: window %lld has %lld runs, room for %lld", (long long)k,
                      (long long)nr, (long long)max_runs);
        rows.clear();
        rows.push_back({max_below, -1.0, -1.0});
        for (int64_t r = 0; r < nr; ++r) {
            const double* p = runs + ((size_t)k * max_runs + r) * 3;
            rows.push_back({p[2], p[0], p[1]});
        }
        // descending by max error, stable, NaN last (pandas sort_values(ascending=False), :1224)
        std::stable_sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) {
            const bool an = a.m != a.m, bn = b.m != b.m;
            if (an || bn) return !an && bn;
            return a.m > b.m;
        });
        // the last position whose drop to the next maximum is not "too small" (increase < min_percent is False)
        int64_t last = -1;
        for (size_t i = 0; i + 1 < rows.size(); ++i)
            if (!((rows[i].m - rows[i + 1].m) / rows[i].m < min_percent)) last = (int64_t)i;
        const double denom = f32 ? (double)(float)(mean + sd) : mean + sd;
        const double shift = (double)(k * step);
        for (int64_t i = 0; i <= last; ++i) {
            const double d = rows[i].m - thr;
            const double score = f32 ? (double)(float)((double)(float)d / denom) : d / denom;
            seqs.push_back({rows[i].s + shift, rows[i].e + shift, score});
        }
    }
    auto emit = [&](double s, double e, double sc) {
        merged.push_back(s);
        merged.push_back(e);
        merged.push_back(sc);
    };
    if (!seqs.empty()) {
        std::stable_sort(seqs.begin(), seqs.end(), [](const Seq& a, const Seq& b) { return a.s < b.s; });
        std::vector<double> score, weights, prod;
        Seq cur = seqs[0];
        score.assign(1, cur.score);
        weights.assign(1, cur.e - cur.s);
        bool grouped = false;
        auto close = [&]() -> int {
            if (grouped) {  // np.average(score, weights=weights)
                const double scl = np_pairwise_sum_host(weights.data(), (int64_t)weights.size());
                if (scl == 0.0) {
                    set_error("Weights sum to zero, can't be normalized");
                    return HYPAD_EZERODIV;
                }
                prod.resize(score.size());
                for (size_t i = 0; i < score.size(); ++i) prod[i] = score[i] * weights[i];
                cur.score = np_pairwise_sum_host(prod.data(), (int64_t)prod.size()) / scl;
            }
            emit(cur.s, cur.e, cur.score);
            return HYPAD_OK;
        };
        for (size_t i = 1; i < seqs.size(); ++i) {
            const Seq& q = seqs[i];
            if (q.s <= cur.e + 1) {
                score.push_back(q.score);
                weights.push_back(q.e - q.s);
                cur.e = q.e > cur.e ? q.e : cur.e;
                grouped = true;
            } else {
                const int rc = close();
                if (rc != HYPAD_OK) return rc;
                cur = q;
                score.assign(1, q.score);
                weights.assign(1, q.e - q.s);
                grouped = false;
            }
        }
        const int rc = close();
        if (rc != HYPAD_OK) return rc;
    }
    return HYPAD_OK;
}

extern "C" int hypad_intervals_from_runs(const double* stats, const double* runs, const int32_t* n_runs, int64_t count,
                                         int64_t max_runs, int64_t step, double min_percent, int f32, double* out, int64_t cap,
                                         int64_t* n_out) {
    HYPAD_REQUIRE(stats && runs && n_runs && n_out && count >= 0 && max_runs >= 1 && (out || cap == 0), "hypad_intervals_from_runs: bad argument");
    std::vector<double> merged;
    const int rc = intervals_core(stats, runs, n_runs, count, max_runs, step, min_percent, f32, merged);
    if (rc != HYPAD_OK) return rc;
    const int64_t n = (int64_t)merged.size() / 3;
    for (int64_t i = 0; i < n && i < cap; ++i)
        for (int j = 0; j < 3; ++j) out[i * 3 + j] = merged[(size_t)i * 3 + j];
    *n_out = n;
    return HYPAD_OK;
}

// The host tails of a whole sweep in one call: item i's packed thresholding result (stats | runs | n_runs, the layout of
// hypad_score_signal_hyperbolic's `tw`) starts at host_buf + offsets[i] doubles and holds counts[i] analysis windows of step
// steps[i].  out receives the items' (start, end, score) triples back to back, n_out[i] their number per item -- or -1 for an item
// one of whose windows holds more runs than max_runs (the caller redoes that signal with more room) and -2 for an item whose
// merge hits numpy's "Weights sum to zero" (the caller raises for it).  *total = triples written; more than cap: call again.
extern "C" int hypad_sweep_intervals(const double* host_buf, int64_t n_items, const int64_t* offsets, const int64_t* counts,
                                     const int64_t* steps, int max_runs, double min_percent, int f32, double* out, int64_t cap,
                                     int64_t* n_out, int64_t* total) {
    HYPAD_REQUIRE(host_buf && offsets && counts && steps && n_out && total && n_items >= 0 && max_runs >= 1 && (out || cap == 0),
                  "hypad_sweep_intervals: bad argument");
    std::vector<double> merged;
    int64_t written = 0;
    for (int64_t i = 0; i < n_items; ++i) {
        const double* stats = host_buf + offsets[i];
        const double* runs = stats + counts[i] * 4;
        const int32_t* n_runs = reinterpret_cast<const int32_t*>(runs + counts[i] * (int64_t)max_runs * 3);
        bool over = false;
        for (int64_t k = 0; k < counts[i]; ++k) over |= n_runs[k] > max_runs;
        if (over) {
            n_out[i] = -1;
            continue;
        }
        merged.clear();
        const int rc = intervals_core(stats, runs, n_runs, counts[i], max_runs, steps[i], min_percent, f32, merged);
        if (rc == HYPAD_EZERODIV) {
            n_out[i] = -2;
            continue;
        }
        if (rc != HYPAD_OK) return rc;
        const int64_t n = (int64_t)merged.size() / 3;
        n_out[i] = n;
        for (int64_t j = 0; j < n * 3; ++j)
            if (written * 3 + j < cap * 3) out[written * 3 + j] = merged[(size_t)j];
        written += n;
    }
    *total = written;
    return HYPAD_OK;
}

// Joins the run fragments of a sharded find_anomalies (hypad_tw_shard_runs records of every rank, rank order, host memory) into
// the per-window statistics and runs hypad_intervals_from_runs takes.  Ranks own increasing position ranges, so concatenating
// their sorted starts / ends gives the global order; the r-th start pairs with the r-th end; a rank's `lead` (the maximum of its
// above-threshold values in front of its first start) belongs to the run that was open when its range began.
// cap: runs per window the output has room for; *max_needed = the largest run count of a window (call again when > cap),
// *overflow = 1 when a rank's own fragment list did not fit max_runs (redo hypad_tw_shard_runs with more room).
extern "C" int hypad_tw_shard_merge(const double* records, int world, int64_t n_analysis, int max_runs, double* stats, double* runs,
                                    int32_t* n_runs, int64_t cap, int64_t* max_needed, int* overflow) {
    HYPAD_REQUIRE(records && stats && runs && n_runs && max_needed && overflow && world >= 1 && n_analysis >= 1 && max_runs >= 1 && cap >= 1,
                  "hypad_tw_shard_merge: bad argument");
    auto unkey = [](unsigned long long k) {
        unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
        double d;
        memcpy(&d, &b, 8);
        return d;
    };
    const size_t per = 8 + 3 * (size_t)max_runs;
    *max_needed = 0;
    *overflow = 0;
    std::vector<double> st, mx, en;
    for (int64_t k = 0; k < n_analysis; ++k) {
        st.clear();
        mx.clear();
        en.clear();
        unsigned long long below = 0ull;
        for (int r = 0; r < world; ++r) {
            const double* o = records + ((size_t)r * n_analysis + k) * per;
            const int64_t ns = (int64_t)o[0], ne = (int64_t)o[1];
            if (ns > max_runs || ne > max_runs) {
                *overflow = 1;
                const int64_t m = ns > ne ? ns : ne;
                if (m > *max_needed) *max_needed = m;
                continue;
            }
            unsigned long long lead, bk;
            memcpy(&lead, o + 2, 8);
            memcpy(&bk, o + 3, 8);
            if (bk > below) below = bk;
            if (lead && !mx.empty()) {
                const double v = unkey(lead);
                if (v > mx.back()) mx.back() = v;
            }
            for (int64_t j = 0; j < ns; ++j) {
                st.push_back(o[8 + j]);
                mx.push_back(o[8 + max_runs + j]);
            }
            for (int64_t j = 0; j < ne; ++j) en.push_back(o[8 + 2 * (size_t)max_runs + j]);
        }
        const double* o0 = records + (size_t)k * per;  // the statistics are the same in every rank's record
        stats[k * 4 + 0] = o0[4];
        stats[k * 4 + 1] = o0[5];
        stats[k * 4 + 2] = o0[6];
        stats[k * 4 + 3] = below ? unkey(below) : 0.0;  // `above.all()` -> max_below = 0 (:1154-1155)
        if (*overflow) continue;
        HYPAD_REQUIRE(st.size() == en.size(), "hypad_tw_shard_merge: window %lld has %zu run starts but %zu ends", (long long)k, st.size(),
                      en.size());
        const int64_t nr = (int64_t)st.size();
        if (nr > *max_needed) *max_needed = nr;
        n_runs[k] = (int32_t)nr;
        if (nr <= cap)
            for (int64_t j = 0; j < nr; ++j) {
                double* p = runs + ((size_t)k * cap + j) * 3;
                p[0] = st[j];
                p[1] = en[j];
                p[2] = mx[j];
            }
    }
    return HYPAD_OK;
}

