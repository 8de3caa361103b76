// extern "C" surface of libhypad_b200.so: context, weight packing, fused forward, Mobius linear.
#include <stdarg.h>

#include <atomic>
#include <vector>

#include "common.cuh"

namespace hypad {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int ensure_workspace(hypad_ctx* ctx, size_t bytes) {
    if (ctx->workspace_bytes >= bytes) return HYPAD_OK;
    // Growing frees the old block: wait for work that may still use it.
    HYPAD_CUDA_TRY(cudaDeviceSynchronize());
    if (ctx->workspace) cudaFree(ctx->workspace);
    ctx->workspace = nullptr;
    ctx->workspace_bytes = 0;
    size_t want = bytes + bytes / 4 + (1 << 20);
    if (cudaMalloc(&ctx->workspace, want) != cudaSuccess) {
        cudaGetLastError();
        set_error("workspace allocation of %zu bytes failed", want);
        return HYPAD_ENOMEM;
    }
    ctx->workspace_bytes = want;
    return HYPAD_OK;
}

int launch_forward(const hypad_ctx* ctx, const void* x, int x_is_f64, int64_t n, int64_t row_stride, const float* z_in,
                   int stages, const hypad_forward_out* out, cudaStream_t stream, const int* guard);
int launch_forward_tc(const hypad_ctx* ctx, const void* x, int x_is_f64, int64_t n, int64_t row_stride, const float* z_in,
                      int stages, const hypad_forward_out* out, cudaStream_t stream);
int pack_tc(hypad_ctx* ctx, const hypad_weights* w, cudaStream_t stream);
int launch_mobius(int device, const float* x, int64_t n, int in_f, int out_f, const float* panel, const float* bias,
                  const float* y2, int has_bias, float* out, cudaStream_t stream);

__global__ void pack_panel_kernel(const ColSrc* __restrict__ cols, int ncols, int kpad, float* __restrict__ panel,
                                  float* __restrict__ bias) {
    const int total = kpad * ncols;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int k = e / ncols, c = e - k * ncols;
        const ColSrc s = cols[c];
        panel[e] = (s.w != nullptr && k < s.K) ? s.w[(size_t)s.row * s.K + k] : 0.0f;
        if (k == 0) {
            bias[c] = s.b1 ? s.b1[s.bidx] : 0.0f;
            bias[ncols + c] = s.b2 ? s.b2[s.bidx] : 0.0f;
        }
    }
}

// bias[0..n) -> dst[0..128) zero padded (optionally through expmap0, hyrnn_nets.py:29-30), y2 = sum fl(b^2)
__global__ void pack_mobius_bias_kernel(const float* __restrict__ bias, int n, int apply_expmap0, float* __restrict__ dst,
                                        float* __restrict__ y2) {
    __shared__ double red[128];
    const int t = threadIdx.x;
    float b = (bias != nullptr && t < n) ? bias[t] : 0.0f;
    if (apply_expmap0) {
        red[t] = (double)__fmul_rn(b, b);
        __syncthreads();
        for (int s = 64; s > 0; s >>= 1) {
            if (t < s) red[t] += red[t + s];
            __syncthreads();
        }
        const float nrm = fmaxf(sqrtf((float)red[0]), 1e-15f);
        const float th = (float)tanh((double)fminf(nrm, 15.0f));
        b = __fmul_rn(th, __fdiv_rn(b, nrm));
        __syncthreads();
    }
    dst[t] = b;
    red[t] = (double)__fmul_rn(b, b);
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
        if (t < s) red[t] += red[t + s];
        __syncthreads();
    }
    if (t == 0) *y2 = (float)red[0];
}

static inline int round8(int v) { return (v + 7) / 8 * 8; }

}  // namespace hypad

using namespace hypad;

extern "C" {

int hypad_abi_version(void) { return HYPAD_ABI_VERSION; }

const char* hypad_last_error(void) { return g_err; }

int64_t hypad_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

int hypad_ctx_create(hypad_ctx** out, int device) {
    HYPAD_REQUIRE(out != nullptr, "hypad_ctx_create: out is NULL");
    int count = 0;
    HYPAD_CUDA_TRY(cudaGetDeviceCount(&count));
    HYPAD_REQUIRE(device >= 0 && device < count, "hypad_ctx_create: device %d out of range (%d devices)", device, count);
    HYPAD_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    HYPAD_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("hypad_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
        return HYPAD_ECUDA;
    }
    hypad_ctx* c = new hypad_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    *out = c;
    return HYPAD_OK;
}

int hypad_ctx_destroy(hypad_ctx* ctx) {
    if (!ctx) return HYPAD_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->packed) cudaFree(ctx->packed);
    if (ctx->workspace) cudaFree(ctx->workspace);
    if (ctx->tc_packed) cudaFree(ctx->tc_packed);
    if (ctx->tc_error) cudaFree(ctx->tc_error);
    if (ctx->tc_debug) cudaFree(ctx->tc_debug);
    if (ctx->fin_state) cudaFree(ctx->fin_state);
    delete ctx;
    return HYPAD_OK;
}

int hypad_pack_weights(hypad_ctx* ctx, const hypad_weights* w, void* stream_) {
    HYPAD_REQUIRE(ctx && w, "hypad_pack_weights: NULL argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int S = w->signal_shape, L = w->latent_dim, C = w->critic_dim, hyp = w->hyperbolic != 0;
    HYPAD_REQUIRE(S >= 1 && S <= 128, "signal_shape %d outside 1..128", S);
    HYPAD_REQUIRE(L >= 1 && L <= 64, "latent_dim %d outside 1..64", L);
    HYPAD_REQUIRE(C >= 1 && C <= 64, "critic_dim %d outside 1..64", C);
    HYPAD_REQUIRE(!hyp || (w->mobius_w && w->mobius_b), "hyperbolic decoder needs mobius_w and mobius_b");
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    const int S8 = round8(S);

    NetProgram prog;
    memset(&prog, 0, sizeof(prog));
    prog.S = S; prog.S8 = S8; prog.latent = L; prog.latent_c = C; prog.hyperbolic = hyp;

    std::vector<std::vector<ColSrc>> cols(P_COUNT);
    auto zero_cols = [](int n) { std::vector<ColSrc> v(n); memset(v.data(), 0, sizeof(ColSrc) * n); return v; };
    auto set_pass = [&](int idx, int kpad, int groups, int src, int dst, int dst_row, int epi, int act) {
        PassDesc& p = prog.pass[idx];
        p.kpad = kpad; p.groups = groups; p.src = src; p.dst = dst; p.dst_row = dst_row; p.epi = epi; p.act = act;
        cols[idx] = zero_cols(groups * 64);
    };
    auto linear_cols = [&](int idx, const float* W, const float* b, int rows, int K) {
        for (int c = 0; c < rows; ++c) {
            ColSrc& s = cols[idx][c];
            s.w = W; s.row = c; s.K = K; s.b1 = b; s.bidx = c;
        }
    };
    // One direction-interleaved LSTM pass: unit slot u = 64*half + c covers hidden unit (u % H) of direction (u / H).
    auto lstm_cols = [&](int idx, int half, int H, int n_units, const float* const* Wd, const float* const* bih,
                         const float* const* bhh, int K) {
        const int gate_row[3] = {0, 2 * H, 3 * H};  // torch gate order i, f, g, o: rows of i, g, o
        for (int c = 0; c < 64; ++c) {
            const int u = 64 * half + c;
            if (u >= n_units) continue;
            const int dir = u / H, j = u % H;
            for (int g = 0; g < 3; ++g) {
                ColSrc& s = cols[idx][g * 64 + c];
                s.w = Wd[dir]; s.row = gate_row[g] + j; s.K = K; s.b1 = bih[dir]; s.b2 = bhh[dir]; s.bidx = gate_row[g] + j;
            }
        }
    };

    // CriticX (models/tadgan.py:91-106)
    set_pass(P_C1, S8, 1, BUF_X, BUF_A, 0, EPI_LINEAR, 2);
    linear_cols(P_C1, w->critic_w[0], w->critic_b[0], C, S);
    const int C8 = round8(C);
    set_pass(P_C2, C8, 1, BUF_A, BUF_B, 0, EPI_LINEAR, 2);
    linear_cols(P_C2, w->critic_w[1], w->critic_b[1], C, C);
    set_pass(P_C3, C8, 1, BUF_B, BUF_A, 0, EPI_LINEAR, 2);
    linear_cols(P_C3, w->critic_w[2], w->critic_b[2], C, C);
    set_pass(P_C4, C8, 1, BUF_A, BUF_B, 0, EPI_LINEAR, 2);
    linear_cols(P_C4, w->critic_w[3], w->critic_b[3], C, C);
    // Encoder (models/tadgan.py:15-27): BiLSTM S -> 2 x 50, Linear 100 -> latent
    set_pass(P_ENC0, S8, 3, BUF_X, BUF_A, 0, EPI_LSTM, 0);
    lstm_cols(P_ENC0, 0, 50, 100, w->enc_w_ih, w->enc_b_ih, w->enc_b_hh, S);
    set_pass(P_ENC1, S8, 3, BUF_X, BUF_A, 64, EPI_LSTM, 0);
    lstm_cols(P_ENC1, 1, 50, 100, w->enc_w_ih, w->enc_b_ih, w->enc_b_hh, S);
    set_pass(P_Z, 104, 1, BUF_A, BUF_B, 0, EPI_LINEAR, 0);
    linear_cols(P_Z, w->enc_dense_w, w->enc_dense_b, L, 100);
    // Decoder (models/tadgan.py:34-62): Linear latent -> 50, 2-layer BiLSTM 50 -> 2x64 -> 2x64, Linear 128 -> S, tanh
    set_pass(P_D0, round8(L), 1, BUF_B, BUF_A, 0, EPI_LINEAR, 0);
    linear_cols(P_D0, w->dec_dense1_w, w->dec_dense1_b, 50, L);
    set_pass(P_L0A, 56, 3, BUF_A, BUF_B, 0, EPI_LSTM, 0);
    lstm_cols(P_L0A, 0, 64, 128, w->dec_w_ih[0], w->dec_b_ih[0], w->dec_b_hh[0], 50);
    set_pass(P_L0B, 56, 3, BUF_A, BUF_B, 64, EPI_LSTM, 0);
    lstm_cols(P_L0B, 1, 64, 128, w->dec_w_ih[0], w->dec_b_ih[0], w->dec_b_hh[0], 50);
    set_pass(P_L1A, 128, 3, BUF_B, BUF_A, 0, EPI_LSTM, 0);
    lstm_cols(P_L1A, 0, 64, 128, w->dec_w_ih[1], w->dec_b_ih[1], w->dec_b_hh[1], 128);
    set_pass(P_L1B, 128, 3, BUF_B, BUF_A, 64, EPI_LSTM, 0);
    lstm_cols(P_L1B, 1, 64, 128, w->dec_w_ih[1], w->dec_b_ih[1], w->dec_b_hh[1], 128);
    set_pass(P_DENSE2, 128, 2, BUF_A, BUF_B, 0, EPI_LINEAR, 1);
    linear_cols(P_DENSE2, w->dec_dense2_w, w->dec_dense2_b, S, 128);
    // MobiusLinear matmul (hyrnn_nets.py:26: F.linear(input, weight), no Euclidean bias)
    set_pass(P_MOB_R, S8, 2, BUF_B, BUF_A, 0, EPI_LINEAR, 0);
    if (hyp) linear_cols(P_MOB_R, w->mobius_w, nullptr, S, S);
    set_pass(P_MOB_X, S8, 2, BUF_X, BUF_B, 0, EPI_LINEAR, 0);  // shares the panel of P_MOB_R

    // offsets
    size_t off = 0, table_cols = 0;
    for (int i = 0; i < P_COUNT; ++i) {
        if (i == P_MOB_X) {
            prog.pass[i].w_off = prog.pass[P_MOB_R].w_off;
            prog.pass[i].b_off = prog.pass[P_MOB_R].b_off;
            continue;
        }
        const int ncols = prog.pass[i].groups * 64;
        prog.pass[i].w_off = (int32_t)off;
        off += (size_t)prog.pass[i].kpad * ncols;
        prog.pass[i].b_off = (int32_t)off;
        off += 2 * (size_t)ncols;
        table_cols += ncols;
    }
    prog.mob_bias_off = (int32_t)off; off += 128;
    prog.mob_y2_off = (int32_t)off; off += 64;
    prog.critic5_off = (int32_t)off; off += 128;

    if (ctx->packed_floats < off) {
        HYPAD_CUDA_TRY(cudaDeviceSynchronize());
        if (ctx->packed) cudaFree(ctx->packed);
        ctx->packed = nullptr;
        ctx->packed_floats = 0;
        HYPAD_CUDA_TRY(cudaMalloc(&ctx->packed, off * sizeof(float)));
        ctx->packed_floats = off;
    }
    int rc = ensure_workspace(ctx, table_cols * sizeof(ColSrc));
    if (rc != HYPAD_OK) return rc;

    std::vector<ColSrc> flat;
    flat.reserve(table_cols);
    std::vector<size_t> start(P_COUNT, 0);
    for (int i = 0; i < P_COUNT; ++i) {
        if (i == P_MOB_X) continue;
        start[i] = flat.size();
        flat.insert(flat.end(), cols[i].begin(), cols[i].end());
    }
    // pageable host source: the copy is staged before the call returns, and ordered on `stream`
    HYPAD_CUDA_TRY(cudaMemcpyAsync(ctx->workspace, flat.data(), flat.size() * sizeof(ColSrc), cudaMemcpyHostToDevice, stream));
    HYPAD_CUDA_TRY(cudaMemsetAsync(ctx->packed, 0, ctx->packed_floats * sizeof(float), stream));
    for (int i = 0; i < P_COUNT; ++i) {
        if (i == P_MOB_X) continue;
        const PassDesc& p = prog.pass[i];
        const int ncols = p.groups * 64;
        const int total = p.kpad * ncols;
        pack_panel_kernel<<<(total + 255) / 256, 256, 0, stream>>>((const ColSrc*)ctx->workspace + start[i], ncols, p.kpad,
                                                                  ctx->packed + p.w_off, ctx->packed + p.b_off);
        HYPAD_LAUNCH_CHECK();
    }
    pack_mobius_bias_kernel<<<1, 128, 0, stream>>>(hyp ? w->mobius_b : nullptr, S, 0, ctx->packed + prog.mob_bias_off,
                                                   ctx->packed + prog.mob_y2_off);
    HYPAD_LAUNCH_CHECK();
    HYPAD_CUDA_TRY(cudaMemcpyAsync(ctx->packed + prog.critic5_off, w->critic_w[4], C * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    HYPAD_CUDA_TRY(cudaMemcpyAsync(ctx->packed + prog.critic5_off + C, w->critic_b[4], sizeof(float), cudaMemcpyDeviceToDevice, stream));
    // the column table in the workspace must outlive the pack kernels before the workspace is reused
    HYPAD_CUDA_TRY(cudaStreamSynchronize(stream));
    ctx->prog = prog;
    rc = pack_tc(ctx, w, stream);
    if (rc != HYPAD_OK) return rc;
    ctx->has_weights = true;
    return HYPAD_OK;
}

static int check_forward_args(hypad_ctx* ctx, const void* x, int64_t n, int64_t row_stride, const float* z_in, int stages,
                              const hypad_forward_out* out) {
    HYPAD_REQUIRE(ctx && out, "hypad_forward: NULL argument");
    if (!ctx->has_weights) {
        set_error("hypad_forward: call hypad_pack_weights first");
        return HYPAD_ESTATE;
    }
    HYPAD_REQUIRE(n >= 0, "hypad_forward: n < 0");
    HYPAD_REQUIRE((stages & ~HYPAD_STAGE_ALL) == 0 && stages != 0, "hypad_forward: bad stage mask %d", stages);
    const bool need_x = (stages & (HYPAD_STAGE_ENCODER | HYPAD_STAGE_MOBIUS_X | HYPAD_STAGE_CRITIC)) != 0;
    HYPAD_REQUIRE(!need_x || x != nullptr, "hypad_forward: x is NULL");
    HYPAD_REQUIRE(!need_x || row_stride >= 1, "hypad_forward: row_stride must be >= 1");
    if ((stages & HYPAD_STAGE_DECODER) && !(stages & HYPAD_STAGE_ENCODER))
        HYPAD_REQUIRE(z_in != nullptr, "hypad_forward: decoder without encoder needs z_in");
    HYPAD_REQUIRE(!(stages & HYPAD_STAGE_CRITIC) || out->critic, "hypad_forward: critic stage needs out->critic");
    HYPAD_REQUIRE(!(stages & HYPAD_STAGE_MOBIUS_X) || ctx->prog.hyperbolic, "hypad_forward: MOBIUS_X needs a hyperbolic decoder");
    HYPAD_REQUIRE(!out->rec || ((stages & HYPAD_STAGE_DECODER) && (stages & HYPAD_STAGE_MOBIUS_X)),
                  "hypad_forward: rec needs DECODER|MOBIUS_X");
    return HYPAD_OK;
}

int hypad_forward(hypad_ctx* ctx, const void* x, int x_is_f64, int64_t n, int64_t row_stride, const float* z_in,
                  int stages, const hypad_forward_out* out, void* stream) {
    int rc = check_forward_args(ctx, x, n, row_stride, z_in, stages, out);
    if (rc != HYPAD_OK || n == 0) return rc;
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    rc = launch_forward_tc(ctx, x, x_is_f64, n, row_stride, z_in, stages, out, (cudaStream_t)stream);
    if (rc != HYPAD_OK || ctx->strict_range) return rc;
    // Range fallback, decided on the device: the FFMA kernel is queued behind the tensor-core kernel and returns at once unless
    // that kernel raised the range flag, in which case it recomputes every output of the call (no operand limits there).
    return launch_forward(ctx, x, x_is_f64, n, row_stride, z_in, stages, out, (cudaStream_t)stream, ctx->tc_error);
}

int hypad_score_signal_hyperbolic(hypad_ctx* ctx, const void* x, int x_is_f64, int64_t n_windows, int combine_mode,
                                  int64_t tw_window, int64_t tw_step, int64_t tw_count, int ddof_flags, int anomaly_padding,
                                  int max_runs, const hypad_signal_out* o, void* stream) {
    HYPAD_REQUIRE(ctx && x && o && o->critic && o->rec && o->unorm && o->final, "hypad_score_signal_hyperbolic: NULL argument");
    HYPAD_REQUIRE(ctx->has_weights && ctx->prog.hyperbolic, "hypad_score_signal_hyperbolic: the context holds no hyperbolic model");
    HYPAD_REQUIRE(n_windows >= 1, "hypad_score_signal_hyperbolic: no windows to score (signal shorter than the window?)");
    const int S = ctx->prog.S;
    const int64_t n_pos = n_windows + S - 1;
    hypad_forward_out fo;
    memset(&fo, 0, sizeof(fo));
    fo.critic = o->critic;
    fo.rec = o->rec;
    fo.unorm = o->unorm;
    int rc = hypad_forward(ctx, x, x_is_f64, n_windows, 1, nullptr,
                           HYPAD_STAGE_ENCODER | HYPAD_STAGE_DECODER | HYPAD_STAGE_CRITIC | HYPAD_STAGE_MOBIUS_X, &fo, stream);
    if (rc != HYPAD_OK) return rc;
    const bool need_c = combine_mode != 6 && combine_mode != 7;  // rec, rec_uncertainty: no critic scores (:356-360)
    bool combined = false;
    if (need_c) {
        HYPAD_REQUIRE(o->kmax && o->critic_scores, "hypad_score_signal_hyperbolic: kmax / critic_scores buffers missing");
        if ((rc = hypad_kde_argmax_overlap(o->critic, 0, n_windows, n_windows, S, 0, n_pos, o->kmax, stream)) != HYPAD_OK) return rc;
        // final_critic_scores (:365-404): smoothing window trunc(0.01 n_windows); fp32-valued selections: 32-bit keys
        const int64_t smooth = (int64_t)((double)n_windows * 0.01);
        if (n_pos <= hypad_critic_small_max()) {  // short signal: statistics, smoothing and combination in one launch
            rc = hypad_critic_combine_small(ctx, o->kmax, n_pos, smooth, 1, combine_mode, o->rec, o->unorm, n_windows, o->critic_scores,
                                            o->final, stream);
            combined = true;
        } else {
            rc = hypad_critic_scores(ctx, o->kmax, n_pos, smooth, 1, o->critic_scores, stream);
        }
        if (rc != HYPAD_OK) return rc;
    }
    if (!combined && (rc = hypad_combine_scores(combine_mode, need_c ? o->critic_scores : nullptr, o->rec, 1, o->unorm, 0.5, n_windows,
                                                o->final, stream)) != HYPAD_OK)
        return rc;
    if (o->tw && tw_count > 0) {
        double* stats = o->tw;
        double* runs = stats + tw_count * 4;
        int32_t* n_runs = (int32_t*)(runs + tw_count * (int64_t)max_runs * 3);
        rc = hypad_threshold_windows(ctx, o->final, n_windows, tw_window, tw_step, tw_count, ddof_flags, anomaly_padding, stats, runs,
                                     n_runs, max_runs, stream);
    }
    return rc;
}

int hypad_score_signals_hyperbolic(const hypad_sweep_item* items, int64_t n_items, int x_is_f64, int combine_mode, int ddof_flags,
                                   int anomaly_padding, int max_runs, void* const* streams, int n_streams) {
    HYPAD_REQUIRE((items || n_items == 0) && n_items >= 0 && streams && n_streams >= 1, "hypad_score_signals_hyperbolic: bad argument");
    for (int64_t i = 0; i < n_items; ++i) {
        const hypad_sweep_item& it = items[i];
        const int rc = hypad_score_signal_hyperbolic(it.ctx, it.x, x_is_f64, it.n_windows, combine_mode, it.tw_window, it.tw_step, it.tw_count,
                                                     ddof_flags, anomaly_padding, max_runs, &it.out, streams[i % n_streams]);
        if (rc != HYPAD_OK) return rc;
    }
    return HYPAD_OK;
}

int hypad_score_signal_euclidean(hypad_ctx* ctx, const void* x, int x_is_f64, int64_t n_windows, int combine_mode, int rec_error_kind,
                                 double lambda_rec, int64_t tw_window, int64_t tw_step, int64_t tw_count, int ddof_flags,
                                 int anomaly_padding, int max_runs, const hypad_signal_eucl_out* o, void* stream) {
    HYPAD_REQUIRE(ctx && x && o && o->critic && o->eucl && o->kmax && o->critic_scores && o->truth && o->pred && o->errors && o->rec && o->final,
                  "hypad_score_signal_euclidean: NULL argument");
    HYPAD_REQUIRE(ctx->has_weights && !ctx->prog.hyperbolic, "hypad_score_signal_euclidean: the context holds no Euclidean model");
    HYPAD_REQUIRE(n_windows >= 1, "hypad_score_signal_euclidean: no windows to score (signal shorter than the window?)");
    HYPAD_REQUIRE(combine_mode == 0 || combine_mode == 3 || combine_mode == 6 || combine_mode == 8, "hypad_score_signal_euclidean: combine "
                  "mode %d (mult 0, critic 3, rec 6, sum 8)", combine_mode);
    HYPAD_REQUIRE(rec_error_kind >= 0 && rec_error_kind <= 2, "hypad_score_signal_euclidean: rec_error_kind %d (dtw 0, point 1, area 2)", rec_error_kind);
    const int S = ctx->prog.S;
    const int64_t n_pos = n_windows + S - 1;
    const int64_t smooth = (int64_t)((double)n_windows * 0.01);
    hypad_forward_out fo;
    memset(&fo, 0, sizeof(fo));
    fo.critic = o->critic;
    fo.eucl = o->eucl;
    int rc = hypad_forward(ctx, x, x_is_f64, n_windows, 1, nullptr, HYPAD_STAGE_ENCODER | HYPAD_STAGE_DECODER | HYPAD_STAGE_CRITIC, &fo, stream);
    if (rc != HYPAD_OK) return rc;
    if ((rc = hypad_kde_argmax_overlap(o->critic, 0, n_windows, n_windows, S, 0, n_pos, o->kmax, stream)) != HYPAD_OK) return rc;
    // critic scores (:365-404); the single-CTA kernel for short signals (its combination output is not used here: 0 windows)
    if (n_pos <= hypad_critic_small_max())
        rc = hypad_critic_combine_small(ctx, o->kmax, n_pos, smooth, 1, 3, nullptr, nullptr, 0, o->critic_scores, o->final, stream);
    else
        rc = hypad_critic_scores(ctx, o->kmax, n_pos, smooth, 1, o->critic_scores, stream);
    if (rc != HYPAD_OK) return rc;
    // reconstruction_errors (:866-962): truth unrolled, prediction = median over the overlapping windows, error, smoothing
    if ((rc = hypad_true_from_signal(x, x_is_f64, n_windows, 1, S, o->truth, stream)) != HYPAD_OK) return rc;
    if ((rc = hypad_median_overlap(o->eucl, n_windows, S, o->pred, stream)) != HYPAD_OK) return rc;
    if (rec_error_kind == 0) rc = hypad_dtw_error(o->truth, o->pred, 1, n_pos, 10, o->rec, stream);  // o->rec as scratch for the raw error
    else if (rec_error_kind == 1) rc = hypad_point_error(o->truth, o->pred, 1, n_pos, o->rec, stream);
    else rc = hypad_area_error(o->truth, o->pred, 1, n_pos, 10, o->rec, stream);
    if (rc != HYPAD_OK) return rc;
    if ((rc = hypad_rolling_mean_centered(ctx, o->rec, n_pos, smooth, smooth / 2, o->errors, stream)) != HYPAD_OK) return rc;
    if ((rc = hypad_zscore_clip(ctx, o->errors, 0, n_pos, o->rec, stream)) != HYPAD_OK) return rc;  // :523-524
    if ((rc = hypad_combine_scores(combine_mode, o->critic_scores, o->rec, 0, nullptr, lambda_rec, n_pos, o->final, stream)) != HYPAD_OK) return rc;
    if (o->tw && tw_count > 0) {
        double* stats = o->tw;
        double* runs = stats + tw_count * 4;
        int32_t* n_runs = (int32_t*)(runs + tw_count * (int64_t)max_runs * 3);
        rc = hypad_threshold_windows(ctx, o->final, n_pos, tw_window, tw_step, tw_count, ddof_flags, anomaly_padding, stats, runs, n_runs,
                                     max_runs, stream);
    }
    return rc;
}

int hypad_ctx_set_strict_range(hypad_ctx* ctx, int strict) {
    HYPAD_REQUIRE(ctx != nullptr, "hypad_ctx_set_strict_range: NULL context");
    ctx->strict_range = strict != 0;
    return HYPAD_OK;
}

int64_t hypad_ctx_range_fallbacks(hypad_ctx* ctx) { return ctx ? (int64_t)ctx->range_fallbacks : 0; }

int hypad_forward_ffma(hypad_ctx* ctx, const void* x, int x_is_f64, int64_t n, int64_t row_stride, const float* z_in,
                       int stages, const hypad_forward_out* out, void* stream) {
    int rc = check_forward_args(ctx, x, n, row_stride, z_in, stages, out);
    if (rc != HYPAD_OK || n == 0) return rc;
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    return launch_forward(ctx, x, x_is_f64, n, row_stride, z_in, stages, out, (cudaStream_t)stream, nullptr);
}

int hypad_forward_debug_cycles(hypad_ctx* ctx, int enable, long long* h_out) {
    HYPAD_REQUIRE(ctx != nullptr, "hypad_forward_debug_cycles: NULL context");
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    if (enable && !ctx->tc_debug) {
        HYPAD_CUDA_TRY(cudaMalloc(&ctx->tc_debug, HYPAD_DEBUG_SLOTS * sizeof(long long)));
        HYPAD_CUDA_TRY(cudaMemset(ctx->tc_debug, 0, HYPAD_DEBUG_SLOTS * sizeof(long long)));
    }
    if (h_out && ctx->tc_debug) HYPAD_CUDA_TRY(cudaMemcpy(h_out, ctx->tc_debug, HYPAD_DEBUG_SLOTS * sizeof(long long), cudaMemcpyDeviceToHost));
    if (!enable && ctx->tc_debug) {
        HYPAD_CUDA_TRY(cudaDeviceSynchronize());
        cudaFree(ctx->tc_debug);
        ctx->tc_debug = nullptr;
    }
    return HYPAD_OK;
}

int hypad_ctx_poll_error(hypad_ctx* ctx) {
    HYPAD_REQUIRE(ctx != nullptr, "hypad_ctx_poll_error: NULL context");
    if (!ctx->tc_error) return HYPAD_OK;
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    int flag = 0;
    HYPAD_CUDA_TRY(cudaMemcpy(&flag, ctx->tc_error, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag & 1) {
        set_error("forward_tc_kernel: a barrier wait timed out (pipeline protocol error)");
        return HYPAD_ECUDA;
    }
    if (flag & 2) {  // strict mode (or a call still in flight when polled): the outputs of that call are saturated
        set_error("forward_tc_kernel: an activation left the range of the scaled fp16 operand split (|x| < 63, linear / critic "
                  "activations < 255); results of that call are invalid -- use hypad_forward_ffma for such data");
        HYPAD_CUDA_TRY(cudaMemset(ctx->tc_error, 0, sizeof(int)));
        return HYPAD_EINVAL;
    }
    if (flag & 4) {  // the guarded FFMA kernel redid the call: valid results, slower path -- reported, not an error
        ctx->range_fallbacks += 1;
        HYPAD_CUDA_TRY(cudaMemset(ctx->tc_error, 0, sizeof(int)));
    }
    return HYPAD_OK;
}

int hypad_mobius_linear(hypad_ctx* ctx, const float* x, int64_t n, int in_features, int out_features,
                        const float* weight, const float* bias, int hyperbolic_bias, float* out, void* stream_) {
    HYPAD_REQUIRE(ctx && x && weight && out, "hypad_mobius_linear: NULL argument");
    HYPAD_REQUIRE(in_features >= 1 && in_features <= 128 && out_features >= 1 && out_features <= 128,
                  "hypad_mobius_linear: features (%d -> %d) outside 1..128", in_features, out_features);
    if (n <= 0) return n == 0 ? HYPAD_OK : HYPAD_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    const int in8 = round8(in_features), G = (out_features + 63) / 64, ncols = G * 64;
    // workspace: column table | panel [in8][ncols] | zero bias [2*ncols] | ball bias [128] | y2 [64]
    const size_t tbl = (size_t)ncols * sizeof(ColSrc);
    const size_t fl = (size_t)in8 * ncols + 2 * (size_t)ncols + 128 + 64;
    int rc = ensure_workspace(ctx, tbl + fl * sizeof(float) + 256);
    if (rc != HYPAD_OK) return rc;
    std::vector<ColSrc> cols(ncols);
    memset(cols.data(), 0, tbl);
    for (int c = 0; c < out_features; ++c) {
        cols[c].w = weight; cols[c].row = c; cols[c].K = in_features;
    }
    char* base = (char*)ctx->workspace;
    ColSrc* d_tbl = (ColSrc*)base;
    float* panel = (float*)(base + ((tbl + 255) / 256) * 256);
    float* zb = panel + (size_t)in8 * ncols;
    float* ball = zb + 2 * ncols;
    float* y2 = ball + 128;
    HYPAD_CUDA_TRY(cudaMemcpyAsync(d_tbl, cols.data(), tbl, cudaMemcpyHostToDevice, stream));
    const int total = in8 * ncols;
    pack_panel_kernel<<<(total + 255) / 256, 256, 0, stream>>>(d_tbl, ncols, in8, panel, zb);
    HYPAD_LAUNCH_CHECK();
    pack_mobius_bias_kernel<<<1, 128, 0, stream>>>(bias, out_features, bias != nullptr && !hyperbolic_bias, ball, y2);
    HYPAD_LAUNCH_CHECK();
    rc = launch_mobius(ctx->device, x, n, in_features, out_features, panel, ball, y2, bias != nullptr, out, stream);
    if (rc != HYPAD_OK) return rc;
    // the host-side table is a temporary: make sure the H2D copy has consumed it
    HYPAD_CUDA_TRY(cudaStreamSynchronize(stream));
    return HYPAD_OK;
}

}  // extern "C"
