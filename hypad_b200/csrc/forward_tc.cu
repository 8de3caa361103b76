// Fused TadGAN forward on the 5th-generation tensor cores (tcgen05.mma kind::f16, accumulators in TMEM) -- the
// product path of hypad_forward.  Same arithmetic contract as the FFMA kernel in forward.cu (which stays as the
// in-library cross-check, hypad_forward_ffma): Encoder -> Decoder -> MobiusLinear x2 -> Poincare row distance, CriticX.
//
// Every layer is a dense contraction  D[128 windows x N] = A[128 x K] * W[N x K]^T  with fp32-class accuracy obtained
// from three half-precision products of an error-compensated split: with power-of-two scales 2^sa, 2^sw chosen so that
// both pieces stay in fp16's normal range, A 2^sa = A_hi + A_lo and W 2^sw = W_hi + W_lo (A_hi = fp16(A 2^sa),
// A_lo = fp16(A 2^sa - A_hi): 11 + 11 significant bits plus the sign of the residual), and
// D 2^(sa+sw) = A_lo W_hi + A_hi W_lo + A_hi W_hi accumulated in fp32 in TMEM by tcgen05.mma kind::f16 (fp16 products are
// exact in fp32); the epilogue undoes the scale for free inside the bias FMA.  Same accuracy as the 3xTF32 split measured
// in tests/test_gpu_tensor_probe.py (rms 3.4-4.3e-8 of sum|a||w| against 2.6e-8 for the fp32 FFMA order; plain TF32 or
// fp16 is 1000x worse) at HALF the tensor instructions, shared-memory operand traffic and weight bytes: one kind::f16
// MMA covers K = 16.  Range: |activation| 2^sa and |weight| 2^sw must stay below 65504; sw is chosen per layer from the
// weights at pack time, sa is 11 for activations bounded by 1 (LSTM outputs, tanh), 10 for the window itself (|x| < 63),
// 8 for unbounded linear / LeakyReLU outputs (|a| < 255); leaving the range raises the context's sticky error flag.
//
// One persistent CTA per SM owns tiles of 128 windows (TMEM lane = window), TWO tiles in flight: each has its own 64 KB
// operand buffer and its own 256 TMEM columns, and while the epilogue warps work on one tile's accumulators the tensor pipe
// runs the other tile's next pass -- the tensor work hides behind the activation math, which bounds the kernel.  Warp roles:
//   warps 0-15 epilogue: warp w reads TMEM lanes 32*(w%4).. (its 32 windows) and every 4th 8-column chunk (w/4);
//              gate / activation math in registers, then writes the next pass's A operand (hi and lo pieces) into shared
//              memory in the UMMA K-major core-matrix layout, element (row r, feature k) at ((k/8)*128 + r)*16 B + (k%8)*2 B
//   warp 16    weight producer: cp.async.bulk (TMA engine) of one <=16 KB weight stage (1-8 k-steps of 16 k x <=256
//              columns, hi+lo) per mbarrier slot, a ring of 5 slots running ahead across passes and tiles
//   warp 17    MMA issuer: one elected thread issues the three tcgen05.mma per k-step, tcgen05.commit frees the slot and,
//              per pass and tile, signals the epilogue
// A model is 11 passes (15 when the critic cannot ride along, see TcPass): an LSTM layer is a g|i pass (one N = 2 nu
// contraction) and an o pass; the four CriticX layers ride along with the first four passes as a second small block.
// Activations never leave the SM: a tile's operand buffer (64 KB = 128 features x 128 windows x hi/lo) is overwritten in
// place pass by pass (all MMAs of a pass retire before its epilogue runs).  Weights (0.6 MB) stream from L2; biases, scales
// and row-phase parameters are copied to shared memory once per kernel.
#include <cuda_fp16.h>

#include <vector>

#include "common.cuh"

namespace hypad {

constexpr int TC_M = 128;                 // windows per tile
constexpr int TC_NSPLIT = 4;              // epilogue warps per TMEM lane quarter (column chunks dealt round-robin)
constexpr int TC_EPI_THREADS = 128 * TC_NSPLIT;
constexpr int TC_THREADS = TC_EPI_THREADS + 64;
constexpr int TC_PRODUCER_WARP = TC_EPI_THREADS / 32, TC_MMA_WARP = TC_PRODUCER_WARP + 1;
constexpr int TC_PIECE_BYTES = 32768;     // one piece of the A buffer: 128 features x 128 rows x 2 B
constexpr int TC_ACT_BYTES = 2 * TC_PIECE_BYTES;  // hi + lo piece of one tile
constexpr int TC_TILES = 2;               // tiles in flight per CTA (one in the tensor pipe, one in the epilogue warps)
constexpr int TC_TILE_COLS = 256;         // TMEM columns per tile slot
constexpr int TC_STAGE_BYTES = 16384;     // one weight stage: kstage k-steps of 16 k x n columns x (hi + lo) x 2 B
constexpr int TC_NSLOT = 5;
constexpr int TC_RED_BYTES = TC_NSPLIT * TC_M * 2 * 8;  // row reductions of at most two fp64 values per row
constexpr int TC_BIAS_BYTES = 9216;       // shared copy of the small-parameter buffer: biases, Mobius bias, critic output layer, scales
constexpr int TC_XS_FLOATS = 240;         // per tile slot: the 128 + S - 1 samples a tile of sliding windows covers (S <= 113), as fp32
constexpr int TC_CRITIC_SHIFT = 8;        // sa of the critic's hidden activations (unbounded LeakyReLU outputs)
constexpr int TC_CRITIC_K0 = 104;         // operand feature where the critic chain keeps its hidden state when it rides along

enum TcEpi : int32_t { TE_LSTM_GI = 0, TE_LSTM_O, TE_Z, TE_LINEAR, TE_TANH, TE_MOB_R, TE_MOB_X, TE_CRITIC_HID, TE_CRITIC_OUT };
// An LSTM layer is two passes over the same A operand: GI accumulates the cell-candidate and input gates of all units
// as ONE N = 2 nu contraction (columns [0,nu) = g, [nu,2nu) = i), its epilogue leaves tanh(sigmoid(i) tanh(g)) in the i
// columns; O accumulates the output gate over the g columns and its epilogue writes h.  That keeps a tile within 256
// TMEM columns, so two tiles are in flight per CTA.
enum TcPassIdx : int32_t {
    T_ENC_GI = 0, T_ENC_O, T_Z, T_D0, T_L0_GI, T_L0_O, T_L1_GI, T_L1_O, T_D2, T_MR, T_MX, T_C1, T_C2, T_C3, T_C4, T_COUNT
};

struct TcPass {
    int32_t k16;     // K / 16 (k-steps of the pass)
    int32_t k_lo;    // first k-step of the operand buffer the pass reads (0 except for stand-alone critic layers)
    int32_t out_k0;  // operand feature where the epilogue stores its output row (0 except for critic layers)
    int32_t n;       // accumulator columns, multiple of 16, <= 256
    int32_t n_live;  // columns (units for an LSTM pass) the epilogue processes: real outputs rounded up to 8; the rest are
                     // zero-weight padding whose operand features keep whatever finite value they had
    int32_t d_col;   // TMEM column within the tile slot's 256
    int32_t w_off;   // byte offset of the packed weights: [k16]{hi,lo}[2 chunks][n][8] fp16
    int32_t b_off;   // float offset of the biases b[n]
    int32_t epi;     // TcEpi
    int32_t needs_x; // the A operand is the window itself
    int32_t in_shift;   // sa of this pass's A operand
    int32_t kstage;     // k-steps per weight stage (<= 16 KB)
    float out_scale;    // 2^sa of the pass that consumes this pass's output
    // Rider: one CriticX layer contracted in the same pass from the same operand buffer (its hidden state lives in features
    // TC_CRITIC_K0.. of the buffer, beyond everything the main chain uses before the decoder LSTMs).  n2 == 0: none.
    int32_t n2, n2_live;  // accumulator columns / live columns
    int32_t d_col2;       // TMEM column within the tile slot
    int32_t k2_lo, k2_n;  // first k-step and number of k-steps of the rider's weights
    int32_t w_off2, b_off2, kstage2;
    int32_t epi2;         // TE_CRITIC_HID or TE_CRITIC_OUT
    int32_t in_shift2;    // sa of the features the rider reads
};

struct TcProgram {
    TcPass pass[T_COUNT];
    int32_t S, S16, latent, latent_c, hyperbolic;
    int32_t mob_bias_off, mob_y2_off, critic5_off;  // float offsets into the small-parameter buffer
    int32_t bias_floats;  // floats at the start of the small-parameter buffer that the kernel copies to shared memory
    int32_t post_off;   // per pass: {2^sw, 2^-(sa+sw)} of the pass and of its rider, written by tc_wscale_kernel
    int32_t can_ride;   // the program holds riders (S and the encoder's hidden state stay below TC_CRITIC_K0)
};

struct TcParams {
    const void* x;
    const float* z_in;
    const unsigned char* wpacked;  // TF32-split weight stages
    const float* small;            // biases and row-phase parameters
    int64_t n, row_stride;
    int32_t x_is_f64, stages;
    uint32_t pass_mask;
    int32_t want_rowstats;
    int32_t ride;    // the CriticX layers ride along with the encoder passes (riders enabled, T_C1..T_C4 masked out)
    hypad_forward_out out;
    int* error_flag;
    long long* debug;  // optional cycle counters (block 0 only), see hypad_forward_debug_cycles
    TcProgram prog;
};

// ------------------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// two fp32 -> packed fp16x2 (round to nearest even, saturating to +-65504 instead of inf so that a range violation
// can never put a NaN into a contraction); e0 goes to the low half
__device__ __forceinline__ uint32_t pack_h2(float e0, float e1) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
    return r;
}
// the error-compensated split of two (already scaled) values: hi = fp16(t), lo = fp16(t - hi)
__device__ __forceinline__ void split_h2(float t0, float t1, uint32_t& hi, uint32_t& lo) {
    hi = pack_h2(t0, t1);
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    lo = pack_h2(t0 - hf.x, t1 - hf.y);
}

// y / d given r ~ 1/d: one residual correction makes the quotient correctly rounded (barring the usual ties),
// 3 instructions instead of the IEEE division sequence
__device__ __forceinline__ float div_refined(float y, float d, float r) {
    const float q = __fmul_rn(y, r);
    return fmaf(fmaf(-q, d, y), r, q);
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}
// instruction descriptor, kind::f16: fp32 accumulate (bit 4), A and B fp16 (formats 0), K-major both, N >> 3, M = 128
__device__ __forceinline__ uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol error must never hang the GPU.  Returns false on timeout (and raises the error flag).
// HINT_NS > 0 lets the hardware park the thread for up to that long per probe: the producer and MMA warps share their
// scheduler with epilogue warps, and a tight probe loop costs those ~10 % of their issue slots.
template <int HINT_NS = 0>
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* error_flag) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22); ++spin) {
        if (HINT_NS > 0) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(bar), "r"(parity), "r"((uint32_t)HINT_NS)
                : "memory");
        } else {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(bar), "r"(parity)
                : "memory");
        }
        if (done) return true;
    }
    atomicExch(error_flag, 1);
    return false;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// 8 consecutive biases (from the shared-memory copy: every lane reads the same address, one broadcast wavefront each)
__device__ __forceinline__ void ldg8(const float* __restrict__ p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// features k0..k0+7 (k0 a multiple of 8) of row r, times the consumer's scale -> hi / lo fp16 pieces of the A operand
// buffer: one 16-byte core-matrix row each
__device__ __forceinline__ void store_act8(unsigned char* act, int r, int k0, const float (&v)[8], float sc) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_h2(__fmul_rn(v[2 * i], sc), __fmul_rn(v[2 * i + 1], sc), hi[i], lo[i]);
    const int off = ((k0 >> 3) * TC_M + r) * 16;
    *reinterpret_cast<uint4*>(act + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(act + TC_PIECE_BYTES + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
// unbounded activations: flag a value the scaled fp16 split cannot hold (the store itself saturates)
__device__ __forceinline__ void check_range8(const float (&v)[8], float sc, int* error_flag) {
    float m = fabsf(v[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, fabsf(v[i]));
    if (!(m * sc < 65000.0f)) atomicExch(error_flag, 2);
}

// 1 / (1 + exp(-x)).  exp(-x) = 2^(-x log2 e) straight through MUFU.EX2 (relative error 2^-22, like expf's own core):
// without expf's two-step argument reduction the exponent carries a rounding error of |x| 2^-24, which moves the result by
// sigma (1 - sigma) |x| 6e-8 <= 1.3e-8 in absolute terms whatever x is -- below half an ulp of the gate values that matter.
// The reciprocal is MUFU.RCP as it comes (<= 1 ulp of sigma, i.e. <= 6e-8 absolute, next to the 6e-8 of the exponential): the
// Newton step the first version added (HYPAD_SIGMOID_NEWTON) changed no parity count of any golden case and cost 2 of 6
// instructions, 912 times per window.
__device__ __forceinline__ float sigmoid_tc(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(x, -1.4426950408889634f)));
    const float d = __fadd_rn(1.0f, e);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
#ifdef HYPAD_SIGMOID_NEWTON
    return fmaf(r, fmaf(-d, r, 1.0f), r);
#else
    return r;
#endif
}

// tanh(x) for |x| <= 1 (the LSTM cell value c = sigmoid(i) tanh(g)): x + x s C(s) / B(s), s = x^2, a (1,2) rational fit of
// (tanh(x)/x - 1)/s with relative error 1.6e-9 on [0,1].  The correction is at most 0.24 |x|, so the rounding of its 8
// operations (one MUFU.RCP, unrefined) weighs a quarter: measured against fp64 over 2.5M points max 1.5 ulp, mean 0.26 ulp --
// tighter than tanhf (2 ulp) at half its instructions and one MUFU instead of two.
__device__ __forceinline__ float tanh_unit(float x) {
    const float s = __fmul_rn(x, x);
    const float c = fmaf(-0.014719938859343529f, s, -0.3333333432674408f);
    const float b = fmaf(fmaf(0.01575944945216179f, s, 0.44415977597236633f), s, 1.0f);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return fmaf(x, __fmul_rn(__fmul_rn(s, c), r), x);
}

__device__ __forceinline__ float sumsq4(float a, float b, float c, float d) { return fmaf(d, d, fmaf(c, c, fmaf(b, b, __fmul_rn(a, a)))); }
__device__ __forceinline__ float dot4(float a, float b, float c, float d, float w, float x, float y, float z) {
    return fmaf(d, z, fmaf(c, y, fmaf(b, x, __fmul_rn(a, w))));
}

// the 4 column splits of a row live in the 4 warps of one TMEM lane quarter: a 128-thread named barrier per quarter
__device__ __forceinline__ void quarter_bar(int quarter) { asm volatile("bar.sync %0, 128;" ::"r"(2 + quarter) : "memory"); }
// the same barrier, returning the OR of `pred` over the quarter
__device__ __forceinline__ bool quarter_bar_or(int quarter, bool pred) {
    uint32_t out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %2, 0;\n\tbar.red.or.pred p, %1, 128, q;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(out)
        : "r"(2 + quarter), "r"((uint32_t)pred)
        : "memory");
    return out != 0;
}

// sum of the column-split partials of a row (fp64, fixed order); every split gets the same value
template <int NV>
__device__ __forceinline__ void row_allreduce_tc(double (&v)[NV], double* red, int r, int split) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[(i * TC_NSPLIT + split) * TC_M + r] = v[i];
    quarter_bar(r >> 5);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double t = red[(i * TC_NSPLIT) * TC_M + r];
#pragma unroll
        for (int q = 1; q < TC_NSPLIT; ++q) t += red[(i * TC_NSPLIT + q) * TC_M + r];
        v[i] = t;
    }
    quarter_bar(r >> 5);
}

// the window tile (or a caller-provided latent) -> A operand buffer; threads: row = t & 127, chunk lane = t >> 7
template <typename T>
__device__ __forceinline__ void load_rows_to_act(unsigned char* act, const T* __restrict__ x, int64_t w0, int64_t n, int64_t stride,
                                                 int width, int width16, int t, float sc, int* error_flag) {
    const int r = t & (TC_M - 1);
    const bool live = w0 + r < n;
    const T* row = x + (w0 + r) * stride;
    // a thread owns at most four 8-feature chunks (width16 <= 128): all of its loads are issued before the first conversion,
    // so the tile costs one trip to L2 / HBM instead of one per chunk
    float v[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = (t >> 7) + j * TC_NSPLIT;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = 8 * c + e;
            v[j][e] = (live && k < width) ? (float)row[k] : 0.0f;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = (t >> 7) + j * TC_NSPLIT;
        if (c < width16 / 8) {
            check_range8(v[j], sc, error_flag);
            store_act8(act, r, 8 * c, v[j], sc);
        }
    }
}

// Sliding windows (row stride 1): a tile's 128 windows cover 128 + S - 1 consecutive samples, each read by up to S windows and
// by two passes (the encoder's first layer and the Mobius layer of the window itself).  They are fetched from global memory
// ONCE per tile into shared memory as fp32 (one load per thread, one L2 latency) and the operand rows are cut from there:
// lane r of a warp reads xs[r + k], consecutive words, no bank conflict.
template <typename T>
__device__ __forceinline__ void stage_samples(float* xs, const T* __restrict__ x, int64_t w0, int64_t n_samples, int count, int t) {
    for (int i = t; i < count; i += TC_EPI_THREADS) xs[i] = w0 + i < n_samples ? (float)x[w0 + i] : 0.0f;
}
__device__ __forceinline__ void staged_rows_to_act(unsigned char* act, const float* xs, bool live, int width, int width16, int t, float sc,
                                                   int* error_flag) {
    const int r = t & (TC_M - 1);
    for (int c = t >> 7; c < width16 / 8; c += TC_NSPLIT) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = 8 * c + e;
            v[e] = (live && k < width) ? xs[r + k] : 0.0f;
        }
        check_range8(v, sc, error_flag);
        store_act8(act, r, 8 * c, v, sc);
    }
}

constexpr int TC_ROWCH = 4;  // 8-column chunks a thread owns of a <= 128 column row phase (chunk j = columns 8 (split + 4 j))

// y = x W^T (accumulators of this thread's chunks, TMEM columns col0 + ..) -> q = project(mobius_add(expmap0(y), bias)),
// the reference's MobiusLinear row phase (see forward.cu row_mobius): returned in registers, optionally written back to
// TMEM (to_tmem) and to global memory (gout).  post = 2^-(sa+sw) undoes the operand scales (exact).  Returns the row's
// squared norm that the Poincare distance needs, identical in every split.  Row sums: groups of four terms in fp32 FFMA
// chains, the groups accumulated in fp64 and rounded to fp32 once per sum.
template <bool DBG>
__device__ __forceinline__ float row_mobius_tc(uint32_t trow, int col0, int ncols, const float* __restrict__ bias,
                                               float y2, double* red, int r, int split, float* gout, bool to_tmem, int S, float post,
                                               float (&q)[TC_ROWCH][8], long long* dbgp) {
    const int cbeg = 8 * split, cstep = 8 * TC_NSPLIT;
    long long tprev = DBG ? clock64() : 0;
    auto mark = [&](int slot) {  // debug instantiation: cycles per phase of thread 0
        if (DBG && dbgp) {
            const long long now = clock64();
            dbgp[slot] += now - tprev;
            tprev = now;
        }
    };
    // TMEM reads run at 64 B/clk per SM and all 16 warps want their row at once: the loads are software-pipelined against
    // the sums below (tcgen05.wait::ld waits for every outstanding load, so at most one is in flight behind the math)
    tmem_ld8(trow + col0 + cbeg, q[0]);  // cbeg < 32 <= ncols
    mark(0);
    // One pass over y = x W^T gives both row sums the expmap0 / mobius_add scalars need: with p = tanh(|y|) y / |y|,
    // |p|^2 = (tanh|y| / |y|)^2 sum y^2 and <p, b> = (tanh|y| / |y|) sum y b.  (The reference sums the rounded p_i; the two
    // differ by the rounding noise of the p_i, ~1e-8 relative on quantities that enter 1 - |p|^2 and 1 + 2<p,b> with
    // magnitudes of a few 1e-2 -- far below the fp32 rounding of those expressions.)
    // Row sums: products and groups of four in fp32 (FFMA chains), the groups in fp64 -- the plain fp64 pipe of this GPU
    // is slow enough that widening every term costs a third of the row phase, and a group's three fp32 roundings average
    // out over the 25 groups to ~3e-8 relative, half an ulp of the fp32 value the sum is rounded to.
    double s1[2] = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < TC_ROWCH; ++j)
        if (cbeg + j * cstep < ncols) {
            float bv[8];
            ldg8(bias + cbeg + j * cstep, bv);
            tmem_ld_wait();
            if (j + 1 < TC_ROWCH && cbeg + (j + 1) * cstep < ncols) tmem_ld8(trow + col0 + cbeg + (j + 1) * cstep, q[j + 1]);
#pragma unroll
            for (int i = 0; i < 8; ++i) q[j][i] *= post;
#pragma unroll
            for (int h = 0; h < 8; h += 4) {
                s1[0] += (double)sumsq4(q[j][h], q[j][h + 1], q[j][h + 2], q[j][h + 3]);
                s1[1] += (double)dot4(q[j][h], q[j][h + 1], q[j][h + 2], q[j][h + 3], bv[h], bv[h + 1], bv[h + 2], bv[h + 3]);
            }
        }
    mark(1);
    row_allreduce_tc<2>(s1, red, r, split);
    mark(2);
    const float nrm = fmaxf(sqrtf((float)s1[0]), 1e-15f);
    const float rnrm = __frcp_rn(nrm);
    const double thd = tanh((double)fminf(nrm, 15.0f));
    const float th = (float)thd;
    const double g = (double)th / (double)nrm;
    const float x2 = (float)(g * g * s1[0]), xy = (float)(g * s1[1]);
    const float one_2xy = __fadd_rn(1.0f, __fmul_rn(2.0f, xy));
    const float ca = __fadd_rn(one_2xy, y2);
    const float cb = __fsub_rn(1.0f, x2);
    const float den = fmaxf(__fadd_rn(one_2xy, __fmul_rn(x2, y2)), 1e-15f);
    const float rden = __frcp_rn(den);
    mark(3);
    double s3[1] = {0.0};
#pragma unroll
    for (int j = 0; j < TC_ROWCH; ++j)
        if (cbeg + j * cstep < ncols) {
            float bv[8];
            ldg8(bias + cbeg + j * cstep, bv);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float pi = __fmul_rn(th, div_refined(q[j][i], nrm, rnrm));
                q[j][i] = div_refined(__fadd_rn(__fmul_rn(ca, pi), __fmul_rn(cb, bv[i])), den, rden);
            }
#pragma unroll
            for (int h = 0; h < 8; h += 4) s3[0] += (double)sumsq4(q[j][h], q[j][h + 1], q[j][h + 2], q[j][h + 3]);
        }
    mark(4);
    // last reduction by hand: its trailing barrier also ORs the (rare) "this row must be projected back into the ball"
    // predicate over the quarter, so the projection's extra reduction is collective without costing the common case a barrier
    red[split * TC_M + r] = s3[0];
    quarter_bar(r >> 5);
    double t3 = red[r];
#pragma unroll
    for (int k = 1; k < TC_NSPLIT; ++k) t3 += red[k * TC_M + r];
    float sq = (float)t3;
    const float norm = fmaxf(sqrtf(sq), 1e-15f);
    const float maxnorm = 0.996f;
    const bool proj = norm > maxnorm;
    if (quarter_bar_or(r >> 5, proj)) {
        double s4[1] = {0.0};
        if (proj) {
#pragma unroll
            for (int j = 0; j < TC_ROWCH; ++j)
                if (cbeg + j * cstep < ncols) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        q[j][i] = __fmul_rn(__fdiv_rn(q[j][i], norm), maxnorm);
                    }
#pragma unroll
                    for (int h = 0; h < 8; h += 4) s4[0] += (double)sumsq4(q[j][h], q[j][h + 1], q[j][h + 2], q[j][h + 3]);
                }
        }
        row_allreduce_tc<1>(s4, red, r, split);
        if (proj) sq = (float)s4[0];
    }
    mark(5);
    if (to_tmem) {
#pragma unroll
        for (int j = 0; j < TC_ROWCH; ++j)
            if (cbeg + j * cstep < ncols) tmem_st8(trow + col0 + cbeg + j * cstep, q[j]);
        tmem_st_wait();
    }
    if (gout) {
#pragma unroll
        for (int j = 0; j < TC_ROWCH; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = cbeg + j * cstep + i;
                if (c < S) gout[c] = q[j][i];
            }
    }
    mark(6);
    return sq;
}

// CriticX dense4 + LeakyReLU, then Linear(latent_c -> 1): column split s forms the FFMA chain over columns 8s..8s+7
// (ascending), the partial sums are joined through shared memory in ascending order (quarter-scoped barriers) and split 0
// adds the bias and writes the critic value.  Every epilogue thread of the quarter must call this.
__device__ __forceinline__ void critic_out_tc(uint32_t tcol, float post, const float* __restrict__ bias, const float* __restrict__ w5,
                                              int latent_c, float* out, double* red, int r, int split) {
    float* part = reinterpret_cast<float*>(red);
    const int c = 8 * split;
    float fdot = 0.0f;
    if (c < latent_c) {  // warp-uniform
        float v[8];
        tmem_ld8(tcol + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float tv = fmaf(v[i], post, bias[c + i]);
            tv = tv > 0.0f ? tv : __fmul_rn(tv, 0.2f);
            if (c + i < latent_c) fdot = fmaf(tv, w5[c + i], fdot);
        }
    }
    part[split * TC_M + r] = fdot;
    quarter_bar(r >> 5);
    if (split == 0 && out) {
        float tot = part[r];
#pragma unroll
        for (int q = 1; q < TC_NSPLIT; ++q) tot = __fadd_rn(tot, part[q * TC_M + r]);
        *out = __fadd_rn(tot, w5[latent_c]);
    }
    quarter_bar(r >> 5);
}

// DBG: cycle counters for scripts/tc_cycles.py (hypad_forward_debug_cycles); the product instantiation carries none.
template <bool DBG>
__global__ void __launch_bounds__(TC_THREADS, 1) forward_tc_kernel(const __grid_constant__ TcParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* act_base = smem;                                    // TC_TILES x (hi + lo piece)
    unsigned char* ring = smem + TC_TILES * TC_ACT_BYTES;              // TC_NSLOT weight stages
    double* red = reinterpret_cast<double*>(ring + TC_NSLOT * TC_STAGE_BYTES);
    float* sbias = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(red) + TC_RED_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(sbias) + TC_BIAS_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_NSLOT + 2 * TC_TILES);
    float* xs_base = reinterpret_cast<float*>(tmem_slot + 4);         // TC_TILES x TC_XS_FLOATS staged samples
    const uint32_t bar_full = s_u32(bars), bar_empty = s_u32(bars + TC_NSLOT);
    const uint32_t bar_acc = s_u32(bars + 2 * TC_NSLOT), bar_a = s_u32(bars + 2 * TC_NSLOT + TC_TILES);  // one per tile slot

    const TcProgram& prog = P.prog;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = prog.S, S16 = prog.S16;
    const int64_t ntiles = (P.n + TC_M - 1) / TC_M;
    const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t npairs = (my_tiles + TC_TILES - 1) / TC_TILES;
    // Work order of every role: for each pair of tiles, for each pass, tile slot 0 then tile slot 1.  The MMA of one slot
    // runs while the epilogue warps work on the other slot's accumulators.

    if (warp == TC_PRODUCER_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        for (int s = 0; s < TC_NSLOT; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int s = 0; s < TC_TILES; ++s) {
            mbar_init(bar_acc + 8 * s, 1);
            mbar_init(bar_a + 8 * s, TC_EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::);
    }
    // biases, scales and row-phase parameters: a load from global memory in front of a chunk's math exposes the L2 latency
    for (int i = tid; i < prog.bias_floats; i += TC_THREADS) sbias[i] = P.small[i];
    // operand buffers start as zeros: padding features are never written, and must never be NaN bit patterns
    for (int i = tid; i < TC_TILES * TC_ACT_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4*>(act_base)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == TC_PRODUCER_WARP) {
        // ===== weight producer (one lane) ==================================================================
        if (lane == 0) {
            uint32_t slot = 0, par = 1;  // waiting on parity 1 of a fresh barrier returns at once: the first lap is free
            bool ok = true;
            long long dbg_prod = 0;
            for (int64_t pr = 0; pr < npairs && ok; ++pr)
                for (int p = 0; p < T_COUNT && ok; ++p) {
                    if (!((P.pass_mask >> p) & 1u)) continue;
                    const TcPass& ps = prog.pass[p];
                    for (int sl = 0; sl < TC_TILES && ok; ++sl) {
                        if (pr * TC_TILES + sl >= my_tiles) continue;
                        const int nblk = (P.ride && ps.n2) ? 2 : 1;
                        for (int blk = 0; blk < nblk && ok; ++blk) {
                            const unsigned char* src = P.wpacked + (blk ? ps.w_off2 : ps.w_off);
                            const int k16 = blk ? ps.k2_n : ps.k16, kstage = blk ? ps.kstage2 : ps.kstage, n = blk ? ps.n2 : ps.n;
                            for (int kp = 0; kp < k16; kp += kstage) {
                                const int kk = k16 - kp < kstage ? k16 - kp : kstage;
                                const uint32_t bytes = (uint32_t)(kk * n) * 64u;
                                const long long c0 = DBG ? clock64() : 0;
                                ok = mbar_wait<2000>(bar_empty + 8 * slot, par, P.error_flag);
                                if (DBG) dbg_prod += clock64() - c0;
                                if (!ok) break;
                                mbar_expect_tx(bar_full + 8 * slot, bytes);
                                bulk_g2s(s_u32(ring + slot * TC_STAGE_BYTES), src, bytes, bar_full + 8 * slot);
                                src += bytes;
                                if (++slot == TC_NSLOT) slot = 0, par ^= 1u;
                            }
                        }
                    }
                }
            if (DBG && blockIdx.x == 0) P.debug[4] = dbg_prod;
        }
    } else if (warp == TC_MMA_WARP) {
        // ===== MMA issuer: the whole warp runs the (uniform) control flow, one elected lane issues ============
        // The issuing thread is a scalar instruction stream next to the tensor pipe: whatever it executes between two
        // tcgen05.mma is hidden only while earlier MMAs are still queued.  So descriptors are additive (one IADD per
        // operand), and the full barrier of the NEXT stage is peeked right after the first k-step of the current one.
        uint32_t slot = 0, par = 0;
        uint32_t a_par = 0;  // bit sl = parity of slot sl's next operand hand-over
        bool ok = true, have = false;
        const bool lead = elect_one();  // the one thread that issues every tcgen05.mma / commit of this CTA
        long long dbg_a = 0, dbg_full = 0;
        constexpr bool dbg = DBG;
        const uint32_t ring0 = (s_u32(ring) & 0x3FFFFu) >> 4;
        auto desc64 = [](uint32_t lo32) { return ((uint64_t)0x4008u << 32) | lo32; };  // SBO 128 B, descriptor version 1
        int64_t stages_left = 0;
        for (int p = 0; p < T_COUNT; ++p)
            if ((P.pass_mask >> p) & 1u) {
                stages_left += (prog.pass[p].k16 + prog.pass[p].kstage - 1) / prog.pass[p].kstage;
                if (P.ride && prog.pass[p].n2) stages_left += (prog.pass[p].k2_n + prog.pass[p].kstage2 - 1) / prog.pass[p].kstage2;
            }
        stages_left *= my_tiles;
        for (int64_t pr = 0; pr < npairs && ok; ++pr)
            for (int p = 0; p < T_COUNT && ok; ++p) {
                if (!((P.pass_mask >> p) & 1u)) continue;
                const TcPass& ps = prog.pass[p];
                const int nblk = (P.ride && ps.n2) ? 2 : 1;
#pragma unroll
                for (int sl = 0; sl < TC_TILES; ++sl) {
                    if (pr * TC_TILES + sl >= my_tiles || !ok) continue;
                    const uint32_t a_hi0 = ((s_u32(act_base + sl * TC_ACT_BYTES) & 0x3FFFFu) >> 4) | (128u << 16);  // LBO 2048 B
                    const uint32_t a_lo0 = a_hi0 + (TC_PIECE_BYTES >> 4);
                    long long c0 = dbg ? clock64() : 0;
                    ok = mbar_wait<500>(bar_a + 8 * sl, (a_par >> sl) & 1u, P.error_flag);  // A operand written, TMEM of this slot drained
                    a_par ^= 1u << sl;
                    if (dbg) dbg_a += clock64() - c0;
                    if (!ok) break;
                    tc_fence_after();
                    for (int blk = 0; blk < nblk && ok; ++blk) {
                    // block 0: the pass itself; block 1: its rider (own columns, own k range of the same operand buffer)
                    const uint32_t n = (uint32_t)(blk ? ps.n2 : ps.n);
                    const uint32_t idesc = idesc_f16((int)n);
                    const uint32_t w_lbo = n << 16;  // LBO = 16 n bytes: the two 8-k chunks of a k-step
                    const int k16 = blk ? ps.k2_n : ps.k16, kstage = blk ? ps.kstage2 : ps.kstage, k_lo = blk ? ps.k2_lo : ps.k_lo;
                    const uint32_t d = tmem + (uint32_t)(sl * TC_TILE_COLS + (blk ? ps.d_col2 : ps.d_col));
                    for (int kp = 0; kp < k16; kp += kstage) {
                        const int kk = k16 - kp < kstage ? k16 - kp : kstage;
                        if (!have) {
                            if (dbg) c0 = clock64();
                            ok = mbar_wait(bar_full + 8 * slot, par, P.error_flag);
                            if (dbg) dbg_full += clock64() - c0;
                            if (!ok) break;
                        }
                        have = false;
                        --stages_left;
                        const uint32_t w0 = (ring0 + slot * (TC_STAGE_BYTES >> 4)) | w_lbo;
                        const uint32_t nslot = slot + 1 == TC_NSLOT ? 0 : slot + 1, npar = slot + 1 == TC_NSLOT ? par ^ 1u : par;
                        if (lead) {
                            const uint32_t ka = (uint32_t)(k_lo + kp) * 256u;  // 4096 B per k-step (16 features) of A
                            mma_f16(d, desc64(a_lo0 + ka), desc64(w0), idesc, kp > 0);  // small terms first
                            mma_f16(d, desc64(a_hi0 + ka), desc64(w0 + 2 * n), idesc, 1);
                            mma_f16(d, desc64(a_hi0 + ka), desc64(w0), idesc, 1);
                        }
                        if (stages_left > 0) {
                            // non-blocking peek, in the shadow of the MMAs just queued
                            uint32_t done;
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                                : "=r"(done)
                                : "r"(bar_full + 8 * nslot), "r"(npar)
                                : "memory");
                            have = done != 0;
                        }
                        if (lead) {
                            for (int j = 1; j < kk; ++j) {
                                const uint32_t ka = (uint32_t)(k_lo + kp + j) * 256u, wj = w0 + (uint32_t)j * 4u * n;
                                mma_f16(d, desc64(a_lo0 + ka), desc64(wj), idesc, 1);
                                mma_f16(d, desc64(a_hi0 + ka), desc64(wj + 2 * n), idesc, 1);
                                mma_f16(d, desc64(a_hi0 + ka), desc64(wj), idesc, 1);
                            }
                            mma_commit(bar_empty + 8 * slot);  // slot is free once these MMAs have read it
                        }
                        slot = nslot;
                        par = npar;
                    }
                    }
                    if (lead) mma_commit(bar_acc + 8 * sl);  // accumulators of this pass and slot complete
                }
            }
        if (dbg && blockIdx.x == 0 && lane == 0) {
            P.debug[2] = dbg_a;
            P.debug[3] = dbg_full;
        }
    } else {
        // ===== epilogue warps ==============================================================================
        const int quarter = warp & 3, split = warp >> 2;
        const int r = quarter * 32 + lane;                         // TMEM lane = window within the tile
        const float* __restrict__ small = P.small;
        uint32_t acc_par = 0;   // bit sl = parity of slot sl's next accumulator hand-over
        uint32_t has_x = 0;     // bit sl = slot sl's A operand buffer currently holds the window tile
        uint32_t staged = 0;    // bit sl = slot sl's samples of the current tile are in shared memory
        const bool stage_x = P.row_stride == 1 && TC_M + S - 1 <= TC_XS_FLOATS;
        float sq_mr0 = 0.0f, sq_mr1 = 0.0f;  // squared norm of the reconstruction's hyperbolic point, per slot
        bool ok = true;
        long long dbg_wait = 0, dbg_xload = 0;
        constexpr bool dbg = DBG;
        const long long dbg_t0 = dbg ? clock64() : 0;
        int p_first = 0;
        while (p_first < T_COUNT && !((P.pass_mask >> p_first) & 1u)) ++p_first;

        // makes the A operand of (local tile t, pass p) available in slot sl, then hands the slot to the MMA warp
        auto hand_over = [&](int sl, int64_t t, int p) {
            const TcPass& ps = prog.pass[p];
            unsigned char* act = act_base + sl * TC_ACT_BYTES;
            const int64_t w0 = (blockIdx.x + t * gridDim.x) * TC_M;
            const long long cx0 = dbg ? clock64() : 0;
            const float sc = __int_as_float((127 + ps.in_shift) << 23);
            if (ps.needs_x && !((has_x >> sl) & 1u)) {
                epi_bar();  // every column split of every row is done writing the previous layer's output
                if (stage_x) {
                    float* xs = xs_base + sl * TC_XS_FLOATS;
                    if (!((staged >> sl) & 1u)) {
                        // window w reads samples [w, w + S): the n windows of this call cover n + S - 1 samples
                        if (P.x_is_f64) stage_samples<double>(xs, (const double*)P.x, w0, P.n + S - 1, TC_M + S - 1, tid);
                        else stage_samples<float>(xs, (const float*)P.x, w0, P.n + S - 1, TC_M + S - 1, tid);
                        staged |= 1u << sl;
                        epi_bar();
                    }
                    staged_rows_to_act(act, xs, w0 + (tid & (TC_M - 1)) < P.n, S, S16, tid, sc, P.error_flag);
                } else if (P.x_is_f64) load_rows_to_act<double>(act, (const double*)P.x, w0, P.n, P.row_stride, S, S16, tid, sc, P.error_flag);
                else load_rows_to_act<float>(act, (const float*)P.x, w0, P.n, P.row_stride, S, S16, tid, sc, P.error_flag);
                has_x |= 1u << sl;
            } else if (p == T_D0 && !(P.stages & HYPAD_STAGE_ENCODER)) {
                epi_bar();
                load_rows_to_act<float>(act, P.z_in, w0, P.n, prog.latent, prog.latent, 16 * ps.k16, tid, sc, P.error_flag);
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(bar_a + 8 * sl);
            if (dbg) dbg_xload += clock64() - cx0;
        };
        for (int sl = 0; sl < TC_TILES; ++sl)
            if (sl < my_tiles && p_first < T_COUNT) hand_over(sl, sl, p_first);

        for (int64_t pr = 0; pr < npairs && ok; ++pr)
            for (int p = 0; p < T_COUNT && ok; ++p) {
                if (!((P.pass_mask >> p) & 1u)) continue;
                const TcPass& ps = prog.pass[p];
                const float* __restrict__ b1 = sbias + ps.b_off;
                const float post = sbias[prog.post_off + 4 * p + 1];  // 2^-(sa+sw): exact, folded into the bias FMA
                const int cbeg = 8 * split, cstep = 8 * TC_NSPLIT;
#pragma unroll 1
                for (int sl = 0; sl < TC_TILES; ++sl) {
                    const int64_t t = pr * TC_TILES + sl;
                    if (t >= my_tiles) continue;
                    unsigned char* act = act_base + sl * TC_ACT_BYTES;
                    const int64_t w0 = (blockIdx.x + t * gridDim.x) * TC_M;
                    const bool live = w0 + r < P.n;
                    const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(sl * TC_TILE_COLS);
                    // ---- wait for the accumulators --------------------------------------------------------------
                    const long long cw0 = dbg ? clock64() : 0;
                    ok = mbar_wait(bar_acc + 8 * sl, (acc_par >> sl) & 1u, P.error_flag);
                    acc_par ^= 1u << sl;
                    const long long ce0 = dbg ? clock64() : 0;
                    if (dbg) dbg_wait += ce0 - cw0;
                    if (dbg && blockIdx.x == 0 && tid == 0) P.debug[8 + p] += ce0 - cw0;
                    if (!ok) break;
                    tc_fence_after();
                    if (ps.epi == TE_LSTM_GI) {
                        const int nu = ps.n >> 1;
                        for (int c = cbeg; c < ps.n_live; c += cstep) {
                            float gg[8], gi[8], bg[8], bi[8], tc[8];
                            tmem_ld8(trow + ps.d_col + c, gg);
                            tmem_ld8(trow + ps.d_col + nu + c, gi);
                            ldg8(b1 + c, bg);
                            ldg8(b1 + nu + c, bi);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float vg = fmaf(gg[i], post, bg[i]);
                                const float vi = fmaf(gi[i], post, bi[i]);
                                tc[i] = tanh_unit(__fmul_rn(sigmoid_tc(vi), tanhf(vg)));
                            }
                            tmem_st8(trow + ps.d_col + nu + c, tc);
                        }
                        tmem_st_wait();
                    } else if (ps.epi == TE_LSTM_O) {
                        for (int c = cbeg; c < ps.n_live; c += cstep) {
                            float go[8], tc[8], bo[8], h[8];
                            tmem_ld8(trow + ps.d_col + c, go);
                            tmem_ld8(trow + ps.d_col + ps.n + c, tc);
                            ldg8(b1 + c, bo);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 8; ++i) h[i] = __fmul_rn(sigmoid_tc(fmaf(go[i], post, bo[i])), tc[i]);
                            store_act8(act, r, c, h, ps.out_scale);
                        }
                        has_x &= ~(1u << sl);
                    } else if (ps.epi == TE_Z || ps.epi == TE_LINEAR || ps.epi == TE_TANH || ps.epi == TE_CRITIC_HID || ps.epi == TE_CRITIC_OUT) {
                        float* gout = nullptr;
                        int gw = 0;
                        if (ps.epi == TE_Z && P.out.z && live) gout = P.out.z + (w0 + r) * prog.latent, gw = prog.latent;
                        if (ps.epi == TE_TANH && P.out.eucl && live) gout = P.out.eucl + (w0 + r) * (int64_t)S, gw = S;
                        if (ps.epi == TE_CRITIC_OUT) {
                            critic_out_tc(trow + ps.d_col, post, b1, sbias + prog.critic5_off, prog.latent_c, live ? P.out.critic + w0 + r : nullptr, red, r, split);
                        } else {
                            for (int c = cbeg; c < ps.n_live; c += cstep) {
                                float v[8], bv[8];
                                tmem_ld8(trow + ps.d_col + c, v);
                                ldg8(b1 + c, bv);
                                tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    float tv = fmaf(v[i], post, bv[i]);
                                    if (ps.epi == TE_TANH) tv = tanhf(tv);
                                    else if (ps.epi == TE_CRITIC_HID) tv = tv > 0.0f ? tv : __fmul_rn(tv, 0.2f);
                                    v[i] = tv;
                                }
                                if (ps.epi != TE_TANH) check_range8(v, ps.out_scale, P.error_flag);
                                store_act8(act, r, ps.out_k0 + c, v, ps.out_scale);
                                if (gout) {
#pragma unroll
                                    for (int i = 0; i < 8; ++i)
                                        if (c + i < gw) gout[c + i] = v[i];
                                }
                            }
                        }
                        has_x &= ~(1u << sl);
                    } else if (ps.epi == TE_MOB_R || ps.epi == TE_MOB_X) {
                        const bool is_x = ps.epi == TE_MOB_X;
                        float* gbase = is_x ? P.out.hyper_x : P.out.hyper;
                        float* gout = (gbase && live) ? gbase + (w0 + r) * (int64_t)S : nullptr;
                        float q[TC_ROWCH][8];
                        // the reconstruction's point stays in TMEM until the window's point has been computed
                        const float sq = row_mobius_tc<DBG>(trow, ps.d_col, ps.n_live, sbias + prog.mob_bias_off, sbias[prog.mob_y2_off], red, r,
                                                            split, gout, !is_x, S, post, q, (DBG && blockIdx.x == 0 && tid == 0) ? P.debug + 40 : nullptr);
                        if (!is_x) {
                            if (sl == 0) sq_mr0 = sq;
                            else sq_mr1 = sq;
                        }
                        // ---- row statistics once both hyperbolic points exist ---------------------------------------
                        const bool with_x = (P.pass_mask >> T_MX) & 1u;
                        if (P.want_rowstats && (is_x || !with_x)) {
                            const float sqvnorm = is_x ? (sl == 0 ? sq_mr0 : sq_mr1) : sq;
                            if (is_x && P.out.rec != nullptr) {
                                const TcPass& pm = prog.pass[T_MR];
                                double sd[1] = {0.0};
#pragma unroll
                                for (int j = 0; j < TC_ROWCH; ++j)
                                    if (cbeg + j * cstep < pm.n_live) {
                                        float h[8];
                                        tmem_ld8(trow + pm.d_col + cbeg + j * cstep, h);
                                        tmem_ld_wait();
#pragma unroll
                                        for (int i = 0; i < 8; ++i) h[i] = __fsub_rn(q[j][i], h[i]);
                                        sd[0] += (double)sumsq4(h[0], h[1], h[2], h[3]);
                                        sd[0] += (double)sumsq4(h[4], h[5], h[6], h[7]);
                                    }
                                row_allreduce_tc<1>(sd, red, r, split);
                                if (split == 0 && live) {
                                    const float sqdist = (float)sd[0], squnorm = sq;
                                    const float tt = __fdiv_rn(__fmul_rn(2.0f, sqdist), __fmul_rn(__fsub_rn(1.0f, squnorm), __fsub_rn(1.0f, sqvnorm)));
                                    const float xt = __fadd_rn(__fadd_rn(1.0f, tt), 1e-7f);
                                    P.out.rec[w0 + r] = (float)acosh((double)xt);
                                }
                            }
                            if (split == 0 && live && P.out.unorm) P.out.unorm[w0 + r] = sqrtf(sqvnorm);
                        }
                    }
                    // ---- the CriticX layer riding along with this pass -------------------------------------------
                    if (P.ride && ps.n2) {
                        const float post2 = sbias[prog.post_off + 4 * p + 3];
                        const float* __restrict__ b2 = sbias + ps.b_off2;
                        if (ps.epi2 == TE_CRITIC_OUT) {
                            critic_out_tc(trow + ps.d_col2, post2, b2, sbias + prog.critic5_off, prog.latent_c, live ? P.out.critic + w0 + r : nullptr, red, r, split);
                        } else {
                            const float sc2 = __int_as_float((127 + TC_CRITIC_SHIFT) << 23);
                            for (int c = cbeg; c < ps.n2_live; c += cstep) {
                                float v[8], bv[8];
                                tmem_ld8(trow + ps.d_col2 + c, v);
                                ldg8(b2 + c, bv);
                                tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float tv = fmaf(v[i], post2, bv[i]);
                                    v[i] = tv > 0.0f ? tv : __fmul_rn(tv, 0.2f);
                                }
                                check_range8(v, sc2, P.error_flag);
                                store_act8(act, r, TC_CRITIC_K0 + c, v, sc2);
                            }
                        }
                    }
                    if (dbg && blockIdx.x == 0 && tid == 0) P.debug[24 + p] += clock64() - ce0;
                    // ---- next step of this slot: the following pass of the tile, or the first pass of the slot's next tile
                    int pn = p + 1;
                    while (pn < T_COUNT && !((P.pass_mask >> pn) & 1u)) ++pn;
                    if (pn < T_COUNT) {
                        hand_over(sl, t, pn);
                    } else if (t + TC_TILES < my_tiles) {
                        has_x &= ~(1u << sl);  // a new tile: whatever window the buffer holds is the old tile's
                        staged &= ~(1u << sl);
                        hand_over(sl, t + TC_TILES, p_first);
                    }
                }
            }
        if (dbg && blockIdx.x == 0 && tid == 0) {
            P.debug[0] = clock64() - dbg_t0;
            P.debug[1] = dbg_wait;
            P.debug[5] = dbg_xload;
        }
    }
    if (DBG && blockIdx.x == 0 && tid == 0) P.debug[6] = my_tiles;
    tc_fence_before();
    __syncthreads();
    if (warp == TC_PRODUCER_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

// ------------------------------------------------------------------------------------------------------------
// weight packing: fp32 parameters -> scaled fp16 hi/lo stages [nblk][k16]{hi,lo}[2 chunks][n][8]
// ------------------------------------------------------------------------------------------------------------
// One CTA per pass: sw = the largest power of two that keeps max|w| 2^sw below 2^15; scale[0] = 2^sw,
// scale[1] = 2^-(sa+sw) (what the epilogue multiplies the accumulators by).
__global__ void __launch_bounds__(1024) tc_wscale_kernel(const ColSrc* __restrict__ cols, int ncols, int in_shift, float* __restrict__ scale) {
    // a warp per column (32 warps, <= 8 columns each): the first version walked the <= 256 columns one after the other with the
    // whole CTA, 46 us per pass -- 0.9 ms per model, three times what scoring a short signal costs.  A maximum does not depend on
    // the order, so the scales are bit for bit the same.
    __shared__ float smax[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float m = 0.0f;
    for (int cc = warp; cc < ncols; cc += nwarps) {
        const ColSrc s = cols[cc];
        if (s.w == nullptr) continue;
        for (int k = lane; k < s.K; k += 32) m = fmaxf(m, fabsf(s.w[(size_t)s.row * s.K + k]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) smax[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int w = 1; w < nwarps; ++w) smax[0] = fmaxf(smax[0], smax[w]);
    if (threadIdx.x == 0) {
        int e = 0;
        int sw = 11;
        if (smax[0] > 0.0f && smax[0] < 3.0e38f) {
            frexpf(smax[0], &e);  // max|w| < 2^e
            sw = 15 - e;
        }
        sw = sw > 24 ? 24 : (sw < -60 ? -60 : sw);
        scale[0] = ldexpf(1.0f, sw);
        scale[1] = ldexpf(1.0f, -(sw + in_shift));
    }
}

// f_base: operand feature of the panel's k = 0 (riders start at a later k-step of the shared operand buffer)
__global__ void pack_tc_kernel(const ColSrc* __restrict__ cols, int k16, int nblk, int n, __half* __restrict__ dst, float* __restrict__ bias,
                               const float* __restrict__ scale, int f_base) {
    const int ncols = nblk * n;
    const int total = k16 * 16 * ncols;
    const float wscale = scale[0];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int k = e / ncols, cc = e - k * ncols;
        const int blk = cc / n, c = cc - blk * n;
        const ColSrc s = cols[cc];
        const int ksrc = f_base + k - s.k0;  // index into the source row
        const float v = (s.w != nullptr && ksrc >= 0 && ksrc < s.K) ? s.w[(size_t)s.row * s.K + ksrc] : 0.0f;
        const float t = __fmul_rn(v, wscale);
        const __half hi = __float2half_rn(t), lo = __float2half_rn(t - __half2float(hi));
        const int ks = k >> 4, ch = (k >> 3) & 1, el = k & 7;
        // k-steps of one block are contiguous: a stage is TC_KSTAGE consecutive k-steps of a block
        const size_t off = ((size_t)blk * k16 + ks) * (size_t)n * 32 + (size_t)ch * n * 8 + (size_t)c * 8 + el;
        dst[off] = hi;
        dst[off + (size_t)n * 16] = lo;  // lo piece follows the hi piece: n*32 B = n*16 halves
        if (k == 0) {
            // one bias per column: b_ih + b_hh pre-added in fp32 (the reference adds them one after the other: the
            // results differ by at most one ulp of the gate pre-activation, below the contraction's own rounding)
            const float ba = s.b1 ? s.b1[s.bidx] : 0.0f, bb = s.b2 ? s.b2[s.bidx] : 0.0f;
            bias[cc] = __fadd_rn(ba, bb);
        }
    }
}

static inline int round8i(int v) { return (v + 7) / 8 * 8; }
static inline int round16i(int v) { return (v + 15) / 16 * 16; }

size_t forward_tc_smem_bytes() {
    return (size_t)TC_TILES * TC_ACT_BYTES + (size_t)TC_NSLOT * TC_STAGE_BYTES + TC_RED_BYTES + TC_BIAS_BYTES + (2 * TC_NSLOT + 2 * TC_TILES) * 8 + 16 +
           (size_t)TC_TILES * TC_XS_FLOATS * 4;
}

// Builds the tensor-core program and packs the weights (called from hypad_pack_weights).
int pack_tc(hypad_ctx* ctx, const hypad_weights* w, cudaStream_t stream) {
    const int S = w->signal_shape, L = w->latent_dim, C = w->critic_dim, hyp = w->hyperbolic != 0;
    const int S16 = round16i(S), NS = round16i(S), NL = round16i(L), NC = round16i(C);
    TcProgram prog;
    memset(&prog, 0, sizeof(prog));
    prog.S = S; prog.S16 = S16; prog.latent = L; prog.latent_c = C; prog.hyperbolic = hyp;

    std::vector<std::vector<ColSrc>> cols(T_COUNT);
    // in_shift = sa of the pass's A operand: 10 for the window, 11 for activations bounded by 1, 8 for unbounded ones
    const int SH_X = 10, SH_UNIT = 11, SH_FREE = 8;
    auto set_pass = [&](int idx, int K, int n, int live, int d_col, int epi, int needs_x, int in_shift, int consumer_shift) {
        TcPass& p = prog.pass[idx];
        p.k16 = round16i(K) / 16; p.n = n; p.n_live = round8i(live); p.d_col = d_col; p.epi = epi; p.needs_x = needs_x;
        p.in_shift = in_shift; p.out_scale = ldexpf(1.0f, consumer_shift);
        p.kstage = TC_STAGE_BYTES / (n * 64);
        if (p.kstage > p.k16) p.kstage = p.k16;
        cols[idx].assign((size_t)n, ColSrc{});
    };
    auto linear_cols = [&](int idx, const float* W, const float* b, int rows, int K) {
        for (int c = 0; c < rows; ++c) {
            ColSrc& s = cols[idx][c];
            s.w = W; s.row = c; s.K = K; s.b1 = b; s.bidx = c;
        }
    };
    // torch gate order in weight_ih: i | f | g | o (f is dead: c0 = 0).  GI pass: columns [0,nu) = g, [nu,2nu) = i.
    auto lstm_cols = [&](int idx_gi, int idx_o, int H, int n_units, const float* const* Wd, const float* const* bih, const float* const* bhh, int K) {
        const int nu = prog.pass[idx_o].n;
        for (int u = 0; u < n_units; ++u) {
            const int dir = u / H, j = u % H;
            const int rows[3] = {2 * H + j, j, 3 * H + j};  // g, i, o
            ColSrc* dst[3] = {&cols[idx_gi][u], &cols[idx_gi][(size_t)nu + u], &cols[idx_o][u]};
            for (int g = 0; g < 3; ++g) {
                ColSrc& s = *dst[g];
                s.w = Wd[dir]; s.row = rows[g]; s.K = K; s.b1 = bih[dir]; s.b2 = bhh[dir]; s.bidx = rows[g];
            }
        }
    };
    set_pass(T_ENC_GI, S, 224, 100, 0, TE_LSTM_GI, 1, SH_X, 0);
    set_pass(T_ENC_O, S, 112, 100, 0, TE_LSTM_O, 1, SH_X, SH_UNIT);
    lstm_cols(T_ENC_GI, T_ENC_O, 50, 100, w->enc_w_ih, w->enc_b_ih, w->enc_b_hh, S);
    set_pass(T_Z, 100, NL, L, 0, TE_Z, 0, SH_UNIT, SH_FREE);
    linear_cols(T_Z, w->enc_dense_w, w->enc_dense_b, L, 100);
    set_pass(T_D0, L, 64, 50, 0, TE_LINEAR, 0, SH_FREE, SH_FREE);
    linear_cols(T_D0, w->dec_dense1_w, w->dec_dense1_b, 50, L);
    set_pass(T_L0_GI, 50, 256, 128, 0, TE_LSTM_GI, 0, SH_FREE, 0);
    set_pass(T_L0_O, 50, 128, 128, 0, TE_LSTM_O, 0, SH_FREE, SH_UNIT);
    lstm_cols(T_L0_GI, T_L0_O, 64, 128, w->dec_w_ih[0], w->dec_b_ih[0], w->dec_b_hh[0], 50);
    set_pass(T_L1_GI, 128, 256, 128, 0, TE_LSTM_GI, 0, SH_UNIT, 0);
    set_pass(T_L1_O, 128, 128, 128, 0, TE_LSTM_O, 0, SH_UNIT, SH_UNIT);
    lstm_cols(T_L1_GI, T_L1_O, 64, 128, w->dec_w_ih[1], w->dec_b_ih[1], w->dec_b_hh[1], 128);
    set_pass(T_D2, 128, NS, S, 0, TE_TANH, 0, SH_UNIT, SH_UNIT);
    linear_cols(T_D2, w->dec_dense2_w, w->dec_dense2_b, S, 128);
    set_pass(T_MR, S, NS, S, 0, TE_MOB_R, 0, SH_UNIT, 0);  // stays in columns [0, NS) until the row statistics after T_MX
    if (hyp) linear_cols(T_MR, w->mobius_w, nullptr, S, S);
    set_pass(T_MX, S, NS, S, 128, TE_MOB_X, 1, SH_X, 0);
    if (hyp) linear_cols(T_MX, w->mobius_w, nullptr, S, S);
    // Stand-alone critic layers (CriticX.forward on its own, or S too wide for riders).  When the program has riders they
    // use the riders' geometry -- hidden state at features TC_CRITIC_K0.., k-steps from k_lo -- so that both routes issue the
    // same tensor instructions on the same operands and CriticX.forward equals the fused critic bit for bit.
    const bool ride_geo = S <= TC_CRITIC_K0 && TC_CRITIC_K0 + round8i(C) <= 128;
    const int ck0 = ride_geo ? TC_CRITIC_K0 : 0;
    set_pass(T_C1, S, NC, C, 0, TE_CRITIC_HID, 1, SH_X, SH_FREE);
    prog.pass[T_C1].out_k0 = ck0;
    linear_cols(T_C1, w->critic_w[0], w->critic_b[0], C, S);
    for (int i = 0; i < 3; ++i) {
        const int idx = T_C2 + i;
        set_pass(idx, C, NC, C, 0, i == 2 ? TE_CRITIC_OUT : TE_CRITIC_HID, 0, SH_FREE, SH_FREE);
        linear_cols(idx, w->critic_w[1 + i], w->critic_b[1 + i], C, C);
        if (ck0) {
            TcPass& p = prog.pass[idx];
            p.k_lo = ck0 / 16;
            p.k16 = (round16i(ck0 + C) - ck0 / 16 * 16) / 16;
            if (p.kstage > p.k16) p.kstage = p.k16;
            p.out_k0 = ck0;
            for (int c = 0; c < C; ++c) cols[idx][c].k0 = ck0;
        }
    }
    // the CriticX layers as riders of the first four passes (see TcPass): possible when the window and the encoder's hidden
    // state leave features TC_CRITIC_K0.. of the operand buffer alone
    std::vector<std::vector<ColSrc>> cols2(T_COUNT);
    prog.can_ride = ride_geo;
    auto set_rider = [&](int idx, int K_src, int k0, const float* W, const float* b, int epi2, int in_shift2) {
        TcPass& p = prog.pass[idx];
        p.n2 = NC; p.n2_live = round8i(C); p.d_col2 = 224; p.epi2 = epi2; p.in_shift2 = in_shift2;
        const int f_lo = k0 / 16 * 16, f_hi = round16i(k0 + K_src);
        p.k2_lo = f_lo / 16; p.k2_n = (f_hi - f_lo) / 16;
        p.kstage2 = TC_STAGE_BYTES / (NC * 64);
        if (p.kstage2 > p.k2_n) p.kstage2 = p.k2_n;
        cols2[idx].assign((size_t)NC, ColSrc{});
        for (int c = 0; c < C; ++c) {
            ColSrc& s = cols2[idx][c];
            s.w = W; s.row = c; s.K = K_src; s.b1 = b; s.bidx = c; s.k0 = k0;
        }
    };
    if (prog.can_ride) {
        set_rider(T_ENC_GI, S, 0, w->critic_w[0], w->critic_b[0], TE_CRITIC_HID, SH_X);
        set_rider(T_ENC_O, C, TC_CRITIC_K0, w->critic_w[1], w->critic_b[1], TE_CRITIC_HID, SH_FREE);
        set_rider(T_Z, C, TC_CRITIC_K0, w->critic_w[2], w->critic_b[2], TE_CRITIC_HID, SH_FREE);
        set_rider(T_D0, C, TC_CRITIC_K0, w->critic_w[3], w->critic_b[3], TE_CRITIC_OUT, SH_FREE);
    }
    if (NL > 32 || NC > 32 || NS > 128) {
        set_error("tensor-core path: latent_dim / critic_dim above 32 not supported (got %d / %d)", L, C);
        return HYPAD_EINVAL;
    }
    size_t wbytes = 0, sfloats = 0, ncols_total = 0;
    for (int i = 0; i < T_COUNT; ++i) {
        TcPass& p = prog.pass[i];
        p.w_off = (int32_t)wbytes;
        wbytes += (size_t)p.k16 * p.n * 64;
        p.b_off = (int32_t)sfloats;
        sfloats += (size_t)p.n;
        ncols_total += (size_t)p.n;
        if (p.n2) {
            p.w_off2 = (int32_t)wbytes;
            wbytes += (size_t)p.k2_n * p.n2 * 64;
            p.b_off2 = (int32_t)sfloats;
            sfloats += (size_t)p.n2;
            ncols_total += (size_t)p.n2;
        }
    }
    prog.mob_bias_off = (int32_t)sfloats; sfloats += 128;
    prog.mob_y2_off = (int32_t)sfloats; sfloats += 4;
    prog.critic5_off = (int32_t)sfloats; sfloats += 68;
    prog.post_off = (int32_t)sfloats; sfloats += 4 * T_COUNT;
    prog.bias_floats = (int32_t)sfloats;  // the whole small-parameter buffer lives in shared memory while the kernel runs
    if (sfloats * sizeof(float) > TC_BIAS_BYTES) {
        set_error("tensor-core path: %zu bytes of small parameters exceed the shared-memory copy (%d)", sfloats * sizeof(float), TC_BIAS_BYTES);
        return HYPAD_EINVAL;
    }
    const size_t need = wbytes + sfloats * sizeof(float) + 256;
    if (ctx->tc_bytes < need) {
        HYPAD_CUDA_TRY(cudaDeviceSynchronize());
        if (ctx->tc_packed) cudaFree(ctx->tc_packed);
        ctx->tc_packed = nullptr;
        ctx->tc_bytes = 0;
        HYPAD_CUDA_TRY(cudaMalloc(&ctx->tc_packed, need));
        ctx->tc_bytes = need;
    }
    if (!ctx->tc_error) HYPAD_CUDA_TRY(cudaMalloc(&ctx->tc_error, sizeof(int)));
    HYPAD_CUDA_TRY(cudaMemsetAsync(ctx->tc_error, 0, sizeof(int), stream));
    int rc = ensure_workspace(ctx, ncols_total * sizeof(ColSrc));
    if (rc != HYPAD_OK) return rc;
    std::vector<ColSrc> flat;
    std::vector<size_t> start(T_COUNT, 0), start2(T_COUNT, 0);
    for (int i = 0; i < T_COUNT; ++i) {
        start[i] = flat.size();
        flat.insert(flat.end(), cols[i].begin(), cols[i].end());
        start2[i] = flat.size();
        flat.insert(flat.end(), cols2[i].begin(), cols2[i].end());
    }
    HYPAD_CUDA_TRY(cudaMemcpyAsync(ctx->workspace, flat.data(), flat.size() * sizeof(ColSrc), cudaMemcpyHostToDevice, stream));
    HYPAD_CUDA_TRY(cudaMemsetAsync(ctx->tc_packed, 0, ctx->tc_bytes, stream));
    float* small = reinterpret_cast<float*>(ctx->tc_packed + ((wbytes + 255) / 256) * 256);
    ctx->tc_small_off = ((wbytes + 255) / 256) * 256;
    for (int i = 0; i < T_COUNT; ++i) {
        const TcPass& p = prog.pass[i];
        const int total = p.k16 * 16 * p.n;
        float* scale = small + prog.post_off + 4 * i;
        tc_wscale_kernel<<<1, 1024, 0, stream>>>((const ColSrc*)ctx->workspace + start[i], p.n, p.in_shift, scale);
        HYPAD_LAUNCH_CHECK();
        pack_tc_kernel<<<(total + 255) / 256, 256, 0, stream>>>((const ColSrc*)ctx->workspace + start[i], p.k16, 1, p.n,
                                                               reinterpret_cast<__half*>(ctx->tc_packed + p.w_off), small + p.b_off, scale,
                                                               p.k_lo * 16);
        HYPAD_LAUNCH_CHECK();
        if (p.n2) {
            const int total2 = p.k2_n * 16 * p.n2;
            tc_wscale_kernel<<<1, 1024, 0, stream>>>((const ColSrc*)ctx->workspace + start2[i], p.n2, p.in_shift2, scale + 2);
            HYPAD_LAUNCH_CHECK();
            pack_tc_kernel<<<(total2 + 255) / 256, 256, 0, stream>>>((const ColSrc*)ctx->workspace + start2[i], p.k2_n, 1, p.n2,
                                                                    reinterpret_cast<__half*>(ctx->tc_packed + p.w_off2), small + p.b_off2,
                                                                    scale + 2, p.k2_lo * 16);
            HYPAD_LAUNCH_CHECK();
        }
    }
    // Mobius bias / y2 / critic dense5 come from the FFMA context's packed buffer (already built by hypad_pack_weights)
    HYPAD_CUDA_TRY(cudaMemcpyAsync(small + prog.mob_bias_off, ctx->packed + ctx->prog.mob_bias_off, 128 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    HYPAD_CUDA_TRY(cudaMemcpyAsync(small + prog.mob_y2_off, ctx->packed + ctx->prog.mob_y2_off, sizeof(float), cudaMemcpyDeviceToDevice, stream));
    HYPAD_CUDA_TRY(cudaMemcpyAsync(small + prog.critic5_off, ctx->packed + ctx->prog.critic5_off, (C + 1) * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    HYPAD_CUDA_TRY(cudaStreamSynchronize(stream));
    static_assert(sizeof(TcProgram) <= sizeof(ctx->tc_prog_storage), "tc_prog_storage too small");
    memcpy(ctx->tc_prog_storage, &prog, sizeof(prog));
    return HYPAD_OK;
}

int launch_forward_tc(const hypad_ctx* ctx, const void* x, int x_is_f64, int64_t n, int64_t row_stride, const float* z_in,
                      int stages, const hypad_forward_out* out, cudaStream_t stream) {
    TcParams P;
    memset(&P, 0, sizeof(P));
    memcpy(&P.prog, ctx->tc_prog_storage, sizeof(TcProgram));
    P.x = x; P.z_in = z_in;
    P.wpacked = ctx->tc_packed;
    P.small = reinterpret_cast<const float*>(ctx->tc_packed + ctx->tc_small_off);
    P.n = n; P.row_stride = row_stride; P.x_is_f64 = x_is_f64; P.stages = stages;
    P.out = *out;
    P.error_flag = ctx->tc_error;
    P.debug = ctx->tc_debug;
    const bool hyp = P.prog.hyperbolic != 0;
    uint32_t mask = 0;
    if (stages & HYPAD_STAGE_ENCODER) mask |= (1u << T_ENC_GI) | (1u << T_ENC_O) | (1u << T_Z);
    if (stages & HYPAD_STAGE_DECODER) {
        mask |= (1u << T_D0) | (1u << T_L0_GI) | (1u << T_L0_O) | (1u << T_L1_GI) | (1u << T_L1_O) | (1u << T_D2);
        if (hyp) mask |= (1u << T_MR);
    }
    if (hyp && (stages & HYPAD_STAGE_MOBIUS_X)) mask |= (1u << T_MX);
    // the critic rides along when the encoder and decoder passes it rides on run anyway
    P.ride = P.prog.can_ride && (stages & HYPAD_STAGE_CRITIC) && (stages & HYPAD_STAGE_ENCODER) && (stages & HYPAD_STAGE_DECODER);
    if ((stages & HYPAD_STAGE_CRITIC) && !P.ride) mask |= (1u << T_C1) | (1u << T_C2) | (1u << T_C3) | (1u << T_C4);
    P.pass_mask = mask;
    if (!(hyp && (stages & HYPAD_STAGE_DECODER))) P.out.unorm = nullptr, P.out.hyper = nullptr;
    if (!(hyp && (stages & HYPAD_STAGE_DECODER) && (stages & HYPAD_STAGE_MOBIUS_X))) P.out.rec = nullptr;
    if (!(hyp && (stages & HYPAD_STAGE_MOBIUS_X))) P.out.hyper_x = nullptr;
    P.want_rowstats = (P.out.rec != nullptr) || (P.out.unorm != nullptr);
    const size_t smem = forward_tc_smem_bytes();
    static thread_local bool configured = false;
    if (!configured) {
        HYPAD_CUDA_TRY(cudaFuncSetAttribute(forward_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HYPAD_CUDA_TRY(cudaFuncSetAttribute(forward_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int sms = kNumSMs;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const int64_t ntiles = ceil_div(n, TC_M);
    const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
    if (P.debug) forward_tc_kernel<true><<<grid, TC_THREADS, smem, stream>>>(P);
    else forward_tc_kernel<false><<<grid, TC_THREADS, smem, stream>>>(P);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

}  // namespace hypad
