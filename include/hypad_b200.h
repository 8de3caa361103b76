/*
 * hypad_b200 -- C-ABI of the B200-native HypAD windowed anomaly-scoring hot path.
 *
 * The reference (aleflabo/HypAD) is pure Python/PyTorch: it has no FFI/plugin layer, so there is no
 * existing native interface to mirror.  Each entry point below cites the reference
 * Python function (file:line under the reference tree) whose arithmetic it replaces; the Python
 * modules in hypad_b200/{models,hyperspace,utils}/ keep the reference's names and signatures and
 * marshal into these calls with ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++/torch types.
 *   - Every pointer argument is a DEVICE pointer unless its name starts with `h_`.
 *   - The caller owns every buffer it passes.  The library owns only the opaque hypad_ctx
 *     (packed weights + a growable device workspace).
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     except where documented.
 *   - Return value: 0 on success, a negative HYPAD_E* code otherwise; hypad_last_error() returns a
 *     thread-local human-readable message for the last failure.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef HYPAD_B200_H_
#define HYPAD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HYPAD_OK 0
#define HYPAD_EINVAL (-1)   /* bad argument */
#define HYPAD_ECUDA (-2)    /* CUDA runtime error (message in hypad_last_error) */
#define HYPAD_ENOMEM (-3)   /* workspace allocation failed */
#define HYPAD_EZERODIV (-4) /* np.average's "Weights sum to zero" (merged runs of one position each), as the reference raises */
#define HYPAD_ESTATE (-4)   /* weights not packed / context misuse */

#define HYPAD_ABI_VERSION 2
/* OR-ed into the `ddof` argument of hypad_threshold_windows*: mean, std and threshold are rounded to single precision and
 * the threshold is formed in single precision -- what find_anomalies does when handed a float32 torch tensor
 * (combination "rec" / "rec_uncertainty" of the univariate hyperbolic path, utils/anomaly_detection_utils.py:356-360). */
#define HYPAD_STATS_F32 16

typedef struct hypad_ctx hypad_ctx;

/* Raw fp32 parameter tensors of the three modules, row-major as torch stores them
 * (models/tadgan.py:10-106, hyperspace/hyrnn_nets.py:154-185).  weight_hh_* and the forget-gate rows
 * exist in the modules but never enter the arithmetic (sequence length 1, zero state)
 * and are therefore not passed.  Direction index: 0 = forward, 1 = reverse. */
typedef struct hypad_weights {
    int32_t signal_shape; /* S: window length / channel count (100, 123, 51 ...) 1..128 */
    int32_t latent_dim;   /* Encoder/Decoder latent_space_dim: 20 (train.py:413); 1..64 */
    int32_t hyperbolic;   /* Decoder.hyperbolic */
    int32_t critic_dim;   /* CriticX latent_space_dim (models/tadgan.py:71): 20; 1..64 */
    const float* enc_w_ih[2];  /* encoder.lstm.weight_ih_l0{,_reverse}   (200, S)   */
    const float* enc_b_ih[2];  /* encoder.lstm.bias_ih_l0{,_reverse}     (200,)     */
    const float* enc_b_hh[2];  /* encoder.lstm.bias_hh_l0{,_reverse}     (200,)     */
    const float* enc_dense_w;  /* encoder.dense.weight                   (latent, 100) */
    const float* enc_dense_b;  /* encoder.dense.bias                     (latent,)  */
    const float* dec_dense1_w; /* decoder.dense1.weight                  (50, latent) */
    const float* dec_dense1_b; /* decoder.dense1.bias                    (50,)      */
    const float* dec_w_ih[2][2]; /* decoder.lstm.weight_ih_l{0,1}{,_reverse} (256,50) / (256,128) ; [layer][dir] */
    const float* dec_b_ih[2][2]; /* (256,) */
    const float* dec_b_hh[2][2]; /* (256,) */
    const float* dec_dense2_w; /* decoder.dense2.weight                  (S, 128)   */
    const float* dec_dense2_b; /* decoder.dense2.bias                    (S,)       */
    const float* mobius_w;     /* decoder.hyperbolic_linear.weight       (S, S)  ; NULL when !hyperbolic */
    const float* mobius_b;     /* decoder.hyperbolic_linear.bias         (S,) on the ball ; NULL when !hyperbolic */
    const float* critic_w[5];  /* critic_x.dense{1..5}.weight  (20,S) (20,20) (20,20) (20,20) (1,20) */
    const float* critic_b[5];  /* critic_x.dense{1..5}.bias */
} hypad_weights;

/* Which parts of the fused forward run (bit mask for hypad_forward `stages`). */
#define HYPAD_STAGE_ENCODER 1   /* x -> z                      models/tadgan.py:23-27  */
#define HYPAD_STAGE_DECODER 2   /* z -> eucl (-> hyper)        models/tadgan.py:58-67  */
#define HYPAD_STAGE_MOBIUS_X 4  /* x -> hyper_x                anomaly_detection.py:72-74 */
#define HYPAD_STAGE_CRITIC 8    /* x -> critic                 models/tadgan.py:91-106 */
#define HYPAD_STAGE_ALL 15

/* Optional outputs of hypad_forward; any pointer may be NULL (= not stored). */
typedef struct hypad_forward_out {
    float* z;        /* (N, latent)   Encoder.forward                                     */
    float* eucl;     /* (N, S)        Decoder tanh output (= recons_signal when !hyperbolic) */
    float* hyper;    /* (N, S)        Decoder hyperbolic output (recons_signal)           */
    float* hyper_x;  /* (N, S)        decoder.hyperbolic_linear(window) (real_hyper)      */
    float* critic;   /* (N,)          CriticX.forward                                     */
    float* rec;      /* (N,)  fused row-wise Poincare distance  utils/anomaly_detection_utils.py:58-66 */
    float* unorm;    /* (N,)  ||hyper||_2 (the "uncertainty")   utils/anomaly_detection_utils.py:341-343 */
} hypad_forward_out;

int hypad_abi_version(void);
const char* hypad_last_error(void);
/* Number of kernels this library has launched in the calling process so far (all contexts, all streams). */
int64_t hypad_launch_count(void);

/* Context: packed weights + workspace on `device`. */
int hypad_ctx_create(hypad_ctx** out, int device);
int hypad_ctx_destroy(hypad_ctx* ctx);
/* Re-packs the module parameters into the kernel layout (device-to-device, on `stream`). */
int hypad_pack_weights(hypad_ctx* ctx, const hypad_weights* w, void* stream);

/* Fused network forward over `n` windows (anomaly_detection.py:67-113 for all batches at once).
 *   x, x_is_f64 : input samples, fp32 or fp64 (the reference feeds float64 windows and casts with
 *                 .float() inside forward, models/tadgan.py:24,92).
 *   row_stride  : window w starts at x + w*row_stride.  1 = overlapping windows of a signal
 *                 (utils/dataloader.py:139-222 never materialised), S = materialised rows (N,S).
 *   z_in        : when HYPAD_STAGE_ENCODER is not requested and HYPAD_STAGE_DECODER is, the latent
 *                 input (N, latent) fp32 of Decoder.forward; else NULL.
 * Contractions run on the tensor cores on a scaled hi/lo fp16 split of both operands (fp32-class accuracy,
 * csrc/forward_tc.cu).  Operand range of that kernel: |x| < 63, linear-layer / LeakyReLU activations < 255, weights of
 * any magnitude (scaled per layer at pack time).  A call that leaves the range is redone, on the device and without a host
 * round trip, by the FFMA kernel of hypad_forward_ffma, which is queued behind the tensor-core kernel and returns at once
 * unless that kernel raised its range flag: the results are valid either way, hypad_ctx_poll_error counts such calls
 * (hypad_ctx_range_fallbacks).  With hypad_ctx_set_strict_range(ctx, 1) there is no fallback: the violation raises the
 * context's sticky error and hypad_ctx_poll_error reports it.
 */
int hypad_forward(hypad_ctx* ctx, const void* x, int x_is_f64, int64_t n, int64_t row_stride,
                  const float* z_in, int stages, const hypad_forward_out* out, void* stream);
/* Same contract, contractions on the fp32 FFMA pipe instead of the tensor cores: the in-library
 * cross-check of hypad_forward, not the product path. */
int hypad_forward_ffma(hypad_ctx* ctx, const void* x, int x_is_f64, int64_t n, int64_t row_stride,
                       const float* z_in, int stages, const hypad_forward_out* out, void* stream);
/* Synchronises with the device and reports a sticky error raised inside hypad_forward's kernel (a bounded
 * barrier wait that timed out; in strict mode an operand outside the range).  0 = healthy. */
int hypad_ctx_poll_error(hypad_ctx* ctx);
/* strict != 0: hypad_forward does not fall back to the FFMA kernel on a range violation, it reports it (see hypad_forward). */
int hypad_ctx_set_strict_range(hypad_ctx* ctx, int strict);
/* Number of hypad_ctx_poll_error calls that found a hypad_forward call served by the FFMA fallback since the context was made. */
int64_t hypad_ctx_range_fallbacks(hypad_ctx* ctx);
/* Diagnostic: enable/disable per-role cycle counters of hypad_forward's kernel (CTA 0) and read them back into
 * h_out (host, HYPAD_DEBUG_SLOTS values): [0] epilogue total, [1] epilogue waiting for accumulators, [2] MMA warp waiting
 * for operands, [3] MMA warp waiting for weights, [4] producer waiting for free slots, [5] operand (re)load + hand-over,
 * [6] tiles, [8+p] accumulator wait of pass p, [24+p] epilogue work of pass p (thread 0), [40..47] phases of the Mobius
 * row phase (load+norm, reduce, scalars, expmap sums, reduce, mobius_add, reduce+projection, store). */
#define HYPAD_DEBUG_SLOTS 48
int hypad_forward_debug_cycles(hypad_ctx* ctx, int enable, long long* h_out);

/* hyperspace/hyrnn_nets.py:13-35 mobius_linear with hyperbolic_input=False, k=-1:
 * project(mobius_add(expmap0(x W^T), bias)).  x (n,in) W (out,in) bias (out,) or NULL, out (n,out).
 * hyperbolic_bias=0 applies expmap0 to the bias first (hyrnn_nets.py:29-30).  in,out <= 128. */
int hypad_mobius_linear(hypad_ctx* ctx, const float* x, int64_t n, int in_features, int out_features,
                        const float* weight, const float* bias, int hyperbolic_bias, float* out, void* stream);

/* ---- the step in front of the path: utils/dataloader.py:83-137 ------------------------------ */
/* SignalDataset.time_segments_aggregate(method="mean"), utils/dataloader.py:99-137.  ts_sorted (n_rows) float64 timestamps
 * in ascending order with their values; seg_start (n_segments): the segment starts the reference's loop produces
 * (start, start+interval, ... by repeated addition, while <= the last timestamp).  Segment k takes the rows with
 * seg_start[k] <= t <= seg_start[k] + interval - 1 (pandas label slicing is inclusive); out[k] = their NaN-skipping mean in
 * numpy's summation order, NaN for an empty or all-NaN segment. */
int hypad_segments_aggregate(const double* ts_sorted, const double* values, int64_t n_rows, const double* seg_start,
                             double interval, int64_t n_segments, double* out, void* stream);
/* sklearn SimpleImputer() (mean) then MinMaxScaler(feature_range=(lo, hi)) on one column, utils/dataloader.py:86-89:
 * out = (isnan(x) ? mean(valid x) : x) * scale + (lo - min * scale), scale = (hi - lo) / (max - min) (1 for a constant column). */
int hypad_impute_minmax(hypad_ctx* ctx, const double* x, int64_t n, double lo, double hi, double* out, void* stream);

/* scipy.signal.detrend(type="linear") of the YAHOO branch, utils/dataloader.py:36-38 (_detrend_signal), called :66 and from
 * yahoo_preprocess :42: out = x - (slope * t + intercept), least-squares line over t_i = (i + 1) / n.  out may alias x. */
int hypad_detrend_linear(hypad_ctx* ctx, const double* x, int64_t n, double* out, void* stream);

/* ---- hyperspace/poincare_distance.py (caller hyperspace/losses.py:154) ----------------------- */
/* poincare_distance(pred, gt), :5-16: out (n_pred, n_gt) fp32 row-major,
 * acosh(1 + 2 pairwise_distances(pred, gt)_ij / ((1 - square_norm(pred)_i) (1 - square_norm(gt)_j))). */
int hypad_poincare_distance_pairwise(hypad_ctx* ctx, const float* pred, int64_t n_pred, const float* gt, int64_t n_gt, int D,
                                     float* out, void* stream);
/* pairwise_distances(x, y), :28-48: out (n, m) = clamp(|x_i|^2 + |y_j|^2 - 2 <x_i, y_j>, 1e-7, inf); pass y = x for y=None. */
int hypad_pairwise_sqdist(hypad_ctx* ctx, const float* x, int64_t n, const float* y, int64_t m, int D, float* out, void* stream);
/* square_norm(x), :19-25: clamp(torch.norm(x, dim=-1) ** 2, min=1e-5) per row. */
int hypad_square_norm(const float* x, int64_t n, int D, float* out, void* stream);

/* utils/anomaly_detection_utils.py:58-66: acosh(1 + 2|u-v|^2/((1-|u|^2)(1-|v|^2)) + 1e-7), fp32, per row. */
int hypad_poincare_rowdist(const float* recons, const float* truth, int64_t n, int S, float* out, void* stream);
/* np.linalg.norm(x, axis=1) on fp32 rows, utils/anomaly_detection_utils.py:342. */
int hypad_rownorm(const float* x, int64_t n, int S, float* out, void* stream);
/* np.linalg.norm(true_signal - recons_signal, axis=1), utils/anomaly_detection_utils.py:157 (Euclidean multivariate
 * reconstruction error): truth (n,S) float64 or float32, recons (n,S) float32, out (n,) float64. */
int hypad_rowdiff_norm(const void* truth, int truth_is_f64, const float* recons, int64_t n, int S, double* out, void* stream);

/* utils/dataloader.py:139-222 rolling_window_sequences(window_size=S, target_size=1, step_size=1):
 * out[w, j] = X[w + j], w in [0, n_windows).  Output fp32 or fp64. */
int hypad_window_gather(const double* X, int64_t n_windows, int S, void* out, int out_is_f64, void* stream);

/* utils/anomaly_detection_utils.py:372-400 (twin :471-503): overlap aggregation of one critic value per
 * window into one value per timestep by Gaussian-KDE arg-max, float64.  kmax has n_windows+S-1 entries.
 * Sharded use: `t0`,`t_count` select the timestep range [t0, t0+t_count) written to kmax[0..t_count);
 * critic must still hold all n_windows values (or pass the shard's slice through critic_offset):
 * critic[i] is the value of window critic_offset + i, for i in [0, critic_len). */
int hypad_kde_argmax_overlap(const float* critic, int64_t critic_offset, int64_t critic_len, int64_t n_windows,
                             int S, int64_t t0, int64_t t_count, double* kmax, void* stream);
/* Same result computed by evaluating every density in float64 in scipy's accumulation order (no fp32
 * screening); slower, kept as the in-library cross-check of the screened kernel. */
int hypad_kde_argmax_overlap_exhaustive(const float* critic, int64_t critic_offset, int64_t critic_len,
                                        int64_t n_windows, int S, int64_t t0, int64_t t_count, double* kmax,
                                        void* stream);

/* utils/anomaly_detection_utils.py:307-333 _compute_critic_score: quantile band mean, std, |z|+1, centred
 * rolling mean (window = smooth_window, min_periods = smooth_window/2).  len entries in and out. */
int hypad_critic_zscore_smooth(hypad_ctx* ctx, const double* kmax, int64_t len, int64_t smooth_window, double* out,
                               void* stream);

/* pandas Series.rolling(window, center=True, min_periods).mean() as used at :326-331 and :954-961. */
int hypad_rolling_mean_centered(hypad_ctx* ctx, const double* x, int64_t len, int64_t window, int64_t min_periods,
                                double* out, void* stream);

/* scipy.stats.zscore (ddof=0) then clip(min=0)+1: utils/anomaly_detection_utils.py:523-524, :160-161, :177-178.
 * Input fp64 or fp32 (x_is_f32), output fp64. */
int hypad_zscore_clip(hypad_ctx* ctx, const void* x, int x_is_f32, int64_t len, double* out, void* stream);

/* utils/anomaly_detection_utils.py:336-362 combine_scores on device arrays (float64 result of length n).
 * mode: 0 mult, 1 uncertainty, 2 sum, 3 critic, 4 critic_uncertainty, 5 sum_uncertainty, 6 rec, 7 rec_uncertainty,
 *       8 euclidean "sum" of score_anomalies (:558, lambda_rec) .  rec may be fp32 (rec_is_f32) or fp64.  Every mode is
 * evaluated in float64 on the widened inputs, as numpy does with the reference's float64 critic scores -- except mode 7
 * with an fp32 rec, which is the fp32 product of an fp32 torch tensor and the fp32 norms (:360). */
int hypad_combine_scores(int mode, const double* critic_scores, const void* rec, int rec_is_f32, const float* unorm,
                         double lambda_rec, int64_t n, double* out, void* stream);

/* utils/anomaly_detection_utils.py:918-923: per-timestep median over the anti-diagonal of y_hat (n,S) fp32.
 * pred has n+S-1 fp32 entries (np.median of fp32: even count -> fp32 mean of the two middle values). */
int hypad_median_overlap(const float* y_hat, int64_t n, int S, float* pred, void* stream);

/* utils/anomaly_detection_utils.py:908-910: first sample of every window + tail of the last, as float64. */
int hypad_true_from_signal(const void* x, int x_is_f64, int64_t n, int64_t row_stride, int S, double* out, void* stream);

/* utils/anomaly_detection_utils.py:815-863 _dtw_error (pyts.metrics.dtw defaults restated, see oracle/):
 * len entries; y fp64, y_hat fp32 (the medians) or fp64. */
int hypad_dtw_error(const double* y, const void* y_hat, int y_hat_is_f32, int64_t len, int score_window, double* out,
                    void* stream);
/* :761-777 _point_wise_error and :780-812 _area_error. */
int hypad_point_error(const double* y, const void* y_hat, int y_hat_is_f32, int64_t len, double* out, void* stream);
int hypad_area_error(const double* y, const void* y_hat, int y_hat_is_f32, int64_t len, int score_window, double* out,
                     void* stream);

/* Per-analysis-window statistics of find_anomalies (utils/anomaly_detection_utils.py:1363-1472 with
 * fixed_threshold): for window k covering errors[k*step : k*step+window_size] writes
 * stats[k*4 + {0,1,2,3}] = mean, std(ddof), threshold = mean + 4 std, max of the errors not inside any
 * padded above-threshold run (max_below, 0 when every element is inside).  `runs` receives, per window,
 * up to max_runs (start, end, max_error) triples of the dilated above-threshold runs (window-relative
 * indices as doubles) and n_runs[k] their count (may exceed max_runs: caller re-runs with more room). */
int hypad_threshold_windows(hypad_ctx* ctx, const double* errors, int64_t len, int64_t window_size, int64_t step,
                            int64_t n_analysis, int ddof, int anomaly_padding, double* stats, double* runs,
                            int32_t* n_runs, int max_runs, void* stream);
/* hypad_threshold_windows for the analysis windows [first_window, first_window + n_analysis) of the same array (outputs
 * indexed from 0): what one rank computes when the windows are dealt out to several GPUs.  Bitwise the corresponding rows
 * of the full call -- block sums and the shift sample are defined on the whole array. */
int hypad_threshold_windows_range(hypad_ctx* ctx, const double* errors, int64_t len, int64_t window_size, int64_t step,
                                  int64_t first_window, int64_t n_analysis, int ddof, int anomaly_padding, double* stats,
                                  double* runs, int32_t* n_runs, int max_runs, void* stream);
/* Same contract and results (statistics to the last few bits, runs exactly); every tile of every window is visited
 * element-wise with two-pass statistics: the in-library cross-check of hypad_threshold_windows, which reads the array
 * once (per-block sums and maxima) and visits only the blocks near anomalies and window edges. */
int hypad_threshold_windows_exhaustive(hypad_ctx* ctx, const double* errors, int64_t len, int64_t window_size, int64_t step,
                            int64_t n_analysis, int ddof, int anomaly_padding, double* stats, double* runs,
                            int32_t* n_runs, int max_runs, void* stream);

/* One signal through the whole univariate hyperbolic path in one call (anomaly_detection.py:67-155 + univariate_anomaly_detection,
 * utils/anomaly_detection_utils.py:21-94): fused network over the n_windows sliding windows of x (n_windows + S samples) ->
 * KDE arg-max overlap aggregation -> critic z-score + smoothing -> combine_scores(combine_mode) -> the device part of
 * find_anomalies.  Every buffer is the caller's (device memory); tw, when not NULL, receives the packed thresholding result
 * stats (tw_count, 4) | runs (tw_count, max_runs, 3) | n_runs int32 (tw_count) for hypad_intervals_from_runs.  The same kernels
 * as the step-by-step entry points, queued back to back on `stream` without a host round trip in between: short signals are
 * launch-bound, and a chain of calls through a binding costs more than their kernels. */
typedef struct hypad_signal_out {
    float* critic;         /* (n_windows) */
    float* rec;            /* (n_windows) */
    float* unorm;          /* (n_windows) */
    double* kmax;          /* (n_windows + S - 1); may be NULL for combine modes rec / rec_uncertainty */
    double* critic_scores; /* (n_windows + S - 1); likewise */
    double* final;         /* (n_windows) */
    double* tw;            /* packed thresholding result or NULL */
} hypad_signal_out;
int hypad_score_signal_hyperbolic(hypad_ctx* ctx, const void* x, int x_is_f64, int64_t n_windows, int combine_mode,
                                  int64_t tw_window, int64_t tw_step, int64_t tw_count, int ddof_flags, int anomaly_padding,
                                  int max_runs, const hypad_signal_out* out, void* stream);

/* A sweep over many short signals, one (hyperbolic) model each, in one call: item i is scored by hypad_score_signal_hyperbolic on
 * streams[i % n_streams] (the reference runs `python anomaly_detection.py` once per signal; BASELINE config 5).  A context serves
 * one stream at a time: items that share a context must land on the same stream.  Short signals are bound by what the host
 * spends per signal; this loop leaves a dozen launches per signal and nothing else. */
typedef struct hypad_sweep_item {
    hypad_ctx* ctx;      /* the signal's packed model */
    const void* x;       /* device pointer: the scaled signal, n_windows + S samples */
    int64_t n_windows;
    int64_t tw_window, tw_step, tw_count; /* find_anomalies' analysis windows (tw_count = 0: none) */
    hypad_signal_out out;
} hypad_sweep_item;
int hypad_score_signals_hyperbolic(const hypad_sweep_item* items, int64_t n_items, int x_is_f64, int combine_mode, int ddof_flags,
                                   int anomaly_padding, int max_runs, void* const* streams, int n_streams);

/* The per-timestep Euclidean (TadGAN) path of one signal in one call (score_anomalies, utils/anomaly_detection_utils.py:407-576):
 * network -> KDE critic scores -> truth / median prediction -> reconstruction error (rec_error_kind 0 dtw, 1 point, 2 area) ->
 * smoothing -> z-score + clip -> combination (combine_mode 0 mult, 3 critic, 6 rec, 8 sum with lambda_rec) -> find_anomalies'
 * device part.  n_pos = n_windows + S - 1 positions.  Same kernels as the step-by-step entry points. */
typedef struct hypad_signal_eucl_out {
    float* critic;         /* (n_windows) */
    float* eucl;           /* (n_windows, S) reconstruction */
    double* kmax;          /* (n_pos) */
    double* critic_scores; /* (n_pos) */
    double* truth;         /* (n_pos) */
    float* pred;           /* (n_pos) */
    double* errors;        /* (n_pos) smoothed reconstruction error */
    double* rec;           /* (n_pos) z-scored, clipped, + 1 */
    double* final;         /* (n_pos) */
    double* tw;            /* packed thresholding result or NULL */
} hypad_signal_eucl_out;
int hypad_score_signal_euclidean(hypad_ctx* ctx, const void* x, int x_is_f64, int64_t n_windows, int combine_mode, int rec_error_kind,
                                 double lambda_rec, int64_t tw_window, int64_t tw_step, int64_t tw_count, int ddof_flags,
                                 int anomaly_padding, int max_runs, const hypad_signal_eucl_out* out, void* stream);

/* Host tail of find_anomalies on the outputs of hypad_threshold_windows (host pointers, no device work): prune
 * (utils/anomaly_detection_utils.py:1203-1237), score (:1240-1269) and merge (:1272-1313) with numpy's / pandas' arithmetic.
 * out receives up to `cap` (start, end, score) triples in positions of the scored array, *n_out their number (call again with
 * more room when it exceeds cap).  f32: the scores were a float32 tensor (single-precision interval scores). */
int hypad_intervals_from_runs(const double* stats, const double* runs, const int32_t* n_runs, int64_t count, int64_t max_runs,
                              int64_t step, double min_percent, int f32, double* out, int64_t cap, int64_t* n_out);

/* find_anomalies on an array sharded over several GPUs by contiguous ranges that start at multiples of 1024 positions (only the
 * last may end inside a block): nothing of the array's total length is gathered.  hypad_tw_shard_pack reduces this rank's
 * positions [first, first + count) to a record of hypad_tw_shard_record_doubles(...) doubles -- the summaries of its blocks, the
 * elements of the analysis windows' ragged edges that fall into its range, and its first / last padding + 1 values; the caller
 * all-gathers the records (rank order) and passes them, with this rank's positions plus a halo of padding + 1 either side
 * (`ext`, global positions [ext0, ext0 + ext_len)), to hypad_tw_shard_runs, which computes every window's statistics (the
 * arithmetic of hypad_threshold_windows on the same values) and the run fragments of the own positions: per window
 * 8 + 3 max_runs doubles {n_starts, n_ends, lead key, below key, mean, std, threshold, 0 | starts | their maxima | ends}.
 * The gathered fragment records are joined on the host by hypad_tw_shard_merge into the (stats, runs, n_runs) that
 * hypad_intervals_from_runs takes.  block_start[world + 1]: first block of every rank.  All windows (first_window = 0). */
size_t hypad_tw_shard_record_doubles(int64_t blocks_per_rank, int64_t n_analysis, int anomaly_padding);
int hypad_tw_shard_pack(hypad_ctx* ctx, const double* local, int64_t first, int64_t count, int64_t n_total, int64_t window_size,
                        int64_t step, int64_t n_analysis, int anomaly_padding, int64_t blocks_per_rank, double* record, void* stream);
int hypad_tw_shard_runs(hypad_ctx* ctx, const double* records, int world, int rank, const int64_t* block_start, int64_t blocks_per_rank,
                        const double* ext, int64_t ext0, int64_t ext_len, int64_t n_total, int64_t window_size, int64_t step,
                        int64_t n_analysis, int ddof, int anomaly_padding, int max_runs, double* out, void* stream);
int hypad_tw_shard_merge(const double* records, int world, int64_t n_analysis, int max_runs, double* stats, double* runs,
                         int32_t* n_runs, int64_t cap, int64_t* max_needed, int* overflow);

/* The host tails of a whole sweep in one call: item i's packed thresholding result (the `tw` layout of hypad_score_signal_hyperbolic)
 * starts at host_buf + offsets[i] doubles and holds counts[i] analysis windows of step steps[i].  out receives the items'
 * (start, end, score) triples back to back, n_out[i] their number -- -1 for an item one of whose windows holds more runs than
 * max_runs (redo that signal with more room), -2 where numpy's "Weights sum to zero" would be raised.  *total = triples produced
 * (more than cap: call again with more room). */
int hypad_sweep_intervals(const double* host_buf, int64_t n_items, const int64_t* offsets, const int64_t* counts, const int64_t* steps,
                          int max_runs, double min_percent, int f32, double* out, int64_t cap, int64_t* n_out, int64_t* total);

/* ---- staged global statistics: one GPU's slice per call, a small record exchanged between stages -------------------------
 * The reference computes its statistics on whole arrays (np.quantile / mean / std in _compute_critic_score,
 * utils/anomaly_detection_utils.py:307-333; stats.zscore at :177 and :523).  When a signal's windows are sharded over
 * several GPUs each stage below works on the local slice and leaves a record the caller all-gathers (rank order) before
 * the next stage; with one GPU the records are passed straight on (hypad_critic_zscore_smooth / hypad_zscore_clip do that).
 * The state between stages lives in the context: one chain at a time per context, in stream order. */
#define HYPAD_SELECT_HIST_WORDS 8192 /* uint32 words of one rank's histogram record (4 order statistics x 2048 bins) */
#define HYPAD_MOMENTS_RECORD_DOUBLES 8
/* Number of radix-select passes: 3 when the values are fp32-representable (32-bit keys), else 6. */
int hypad_stats_select_passes(int keys_f32);
/* Starts the selection of the four order statistics around the 25 % / 75 % quantiles of n_total values. */
int hypad_stats_select_begin(hypad_ctx* ctx, int64_t n_total, int keys_f32, void* stream);
/* hist[HYPAD_SELECT_HIST_WORDS] = this slice's digit histogram of pass `pass` (zeroed first). */
int hypad_stats_select_hist(hypad_ctx* ctx, const double* x, int64_t len, int pass, uint32_t* hist, void* stream);
/* hists: `world` records back to back.  After the last pass the quantiles are known to the later stages. */
int hypad_stats_select_pick(hypad_ctx* ctx, const uint32_t* hists, int world, int pass, void* stream);
/* record[8] = (hi, lo) pairs of sum x, sum x^2, and -- band != 0 -- sum / count of the x inside [q25, q75]. */
int hypad_stats_moments_partial(hypad_ctx* ctx, const void* x, int x_is_f32, int64_t len, int band, double* record,
                                void* stream);
/* records: `world` records back to back.  band != 0: critic mean (in band) and std (all, ddof 0) for
 * hypad_critic_smooth_shard; band == 0: mean and std (ddof) for hypad_zscore_clip_apply. */
int hypad_stats_moments_final(hypad_ctx* ctx, const double* records, int world, int64_t n_total, int band, int ddof,
                              void* stream);
/* host8[0..7] = q25, q75, mean(all), mean(in band), std(all), z-score mean, z-score std, 0 (synchronises; diagnostics). */
int hypad_stats_read(hypad_ctx* ctx, double* host8, void* stream);
/* _compute_critic_score (:307-333) on one GPU: the chain of the stages above with the records passed straight on.
 * keys_f32: the values are fp32-representable (KDE selections of fp32 critics): three select passes instead of six.
 * hypad_critic_zscore_smooth is this with keys_f32 = 0. */
int hypad_critic_scores(hypad_ctx* ctx, const double* kmax, int64_t len, int64_t smooth_window, int keys_f32, double* out,
                        void* stream);
/* Short signals (n_pos <= hypad_critic_small_max()): hypad_critic_scores and hypad_combine_scores (rec fp32) in ONE launch of one
 * CTA -- a few thousand positions are launch-bound.  Bitwise the results of the two calls. */
int hypad_critic_small_max(void);
int hypad_critic_combine_small(hypad_ctx* ctx, const double* kmax, int64_t n_pos, int64_t smooth_window, int keys_f32, int combine_mode,
                               const float* rec, const float* unorm, int64_t n_windows, double* critic_scores, double* final,
                               void* stream);
/* :322-331 on a slice: kmax_ext holds the global positions [ext0, ext0 + ext_len) of n_total; out[j] = rolling mean
 * (window smooth_window, centred, min_periods window/2) of |x - mean_band| / std + 1 at position p0 + j, j < count.  The slice
 * must hold the smoothing halo: positions p0 - window/2 .. p0 + count - 1 + (window-1)/2, clipped to [0, n_total). */
int hypad_critic_smooth_shard(hypad_ctx* ctx, const double* kmax_ext, int64_t ext_len, int64_t ext0, int64_t n_total, int64_t p0,
                              int64_t count, int64_t smooth_window, double* out, void* stream);
/* hypad_rolling_mean_centered on a slice: x_ext holds the global positions [ext0, ext0 + ext_len) of an n_total-long array;
 * out[j] = the centred rolling mean at position p0 + j, j < count (same halo rule as hypad_critic_smooth_shard).  The
 * reconstruction-error smoothing (:954-961) of a signal sharded over several GPUs. */
int hypad_rolling_mean_shard(hypad_ctx* ctx, const double* x_ext, int64_t ext_len, int64_t ext0, int64_t n_total, int64_t p0,
                             int64_t count, int64_t window, int64_t min_periods, double* out, void* stream);
/* out = clip((x - mean) / std, 0) + 1 with the mean / std of hypad_stats_moments_final(band = 0). */
int hypad_zscore_clip_apply(hypad_ctx* ctx, const void* x, int x_is_f32, int64_t len, double* out, void* stream);

/* Diagnostic (not on the product path): D (128,N) = A (128,K) B(N,K)^T on the tcgen05 tensor cores with the fp32
 * operands split into `pieces` TF32 parts and `terms` partial products accumulated in TMEM: (1,1) plain TF32,
 * (2,3) 3xTF32, (3,6) six-term split.  Used to measure whether a tensor-core contraction can hold score parity. */
int hypad_tc_probe_gemm(const float* A, const float* B, float* D, int K, int N, int pieces, int terms, void* stream);
/* Diagnostic: cycles for `reps` back-to-back tcgen05.mma (M=128, K=8, tf32, width N) on one SM; h_out2[0] = cycles
 * until retired, h_out2[1] = cycles spent issuing.  mode 0 same accumulator, 1 rotating accumulators, 2 alternating operands. */
int hypad_tc_probe_bench(int N, int reps, int mode, long long* h_out2, void* stream);

/* All-gather of a small record through NVLink peer memory in one kernel launch (no host-side collective): CTA p copies `local`
 * (nbytes, a multiple of 8) into peer p's symmetric buffer at data_off + rank * nbytes, raises flag [flag_off + 8 rank] there to
 * `seq` (release, system scope) and waits until peer p's flag in the LOCAL buffer reaches `seq`.  peer_base_dev: device array of
 * the `world` buffer addresses as mapped in THIS process (torch.distributed._symmetric_memory rendezvous).  After the kernel the
 * records of all ranks lie back to back at data_off of the local buffer.  hypad_b200.distributed.PeerExchange numbers the
 * exchanges and rotates the slots.  *error_flag_dev is set when a peer's record does not arrive within ~10 s. */
int hypad_peer_exchange(const void* local, int64_t nbytes, const int64_t* peer_base_dev, int rank, int world, int64_t data_off,
                        int64_t flag_off, uint64_t seq, int* error_flag_dev, void* stream);

/* Diagnostic (not on the product path): one launch of a pipe-rate micro-benchmark filling every SM with `ctas_per_sm` CTAs
 * of 1024 threads, each running 8 independent chains of `iters` instructions.  kind 0 FP32 FFMA, 1 FP64 DFMA, 2 MUFU.EX2
 * (ex2.approx.ftz.f32), 3 MUFU.EX2 packed (ex2.approx.f16x2), 4 SHFL.IDX.  *instr_per_launch = thread-level instructions
 * issued; the caller times the launch (scripts/measure_peaks.py -> profiles/peaks.json, the denominators SURVEY.md 8(d)
 * asks to be measured on the box). */
int hypad_peak_probe(int kind, int iters, int ctas_per_sm, float* sink, long long* instr_per_launch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HYPAD_B200_H_ */
