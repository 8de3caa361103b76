// Shared host/device helpers for libhypad_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hypad_b200.h"

namespace hypad {

void set_error(const char* fmt, ...);

#define HYPAD_CUDA_TRY(expr)                                                                         \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            ::hypad::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return HYPAD_ECUDA;                                                                      \
        }                                                                                            \
    } while (0)

#define HYPAD_REQUIRE(cond, ...)             \
    do {                                     \
        if (!(cond)) {                       \
            ::hypad::set_error(__VA_ARGS__); \
            return HYPAD_EINVAL;             \
        }                                    \
    } while (0)

void count_launch();

// placed after every <<<...>>>: counts the launch and surfaces launch-configuration errors
#define HYPAD_LAUNCH_CHECK()                      \
    do {                                          \
        ::hypad::count_launch();                  \
        HYPAD_CUDA_TRY(cudaPeekAtLastError());    \
    } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Packed network: a small "layer program" the fused forward kernel interprets.
// A GEMM pass computes 64 windows x (G*64) output columns from an activation buffer held
// transposed in shared memory ([k][64 windows]) and a packed weight panel [Kpad][G][64].
// ---------------------------------------------------------------------------------------------
enum EpilogueKind : int32_t {
    EPI_LSTM = 0,   // 3 column groups = i | g | o of 64 hidden units -> h (64 rows of the output buffer)
    EPI_LINEAR = 1  // out = acc + bias ; act: 0 none, 1 tanh, 2 LeakyReLU(0.2)
};

enum BufId : int32_t { BUF_X = 0, BUF_A = 1, BUF_B = 2 };

struct PassDesc {
    int32_t w_off;    // float offset of the panel [Kpad][G][64] in the packed buffer
    int32_t b_off;    // float offset of bias1 [G*64] followed by bias2 [G*64]
    int32_t kpad;     // multiple of 8
    int32_t groups;   // G in {1,2,3}
    int32_t src;      // BufId of the A operand
    int32_t dst;      // BufId of the output
    int32_t dst_row;  // first feature row written in dst
    int32_t epi;      // EpilogueKind
    int32_t act;      // EPI_LINEAR activation
    int32_t pad0, pad1, pad2;
};

constexpr int kMaxPasses = 16;

// indices into NetProgram::pass (models/tadgan.py layer by layer)
enum PassIdx : int32_t {
    P_C1 = 0, P_C2, P_C3, P_C4,            // CriticX dense1..4 (+LeakyReLU)        :91-104
    P_ENC0, P_ENC1, P_Z,                   // Encoder BiLSTM (2 x 64 unit slots), dense :23-27
    P_D0, P_L0A, P_L0B, P_L1A, P_L1B,      // Decoder dense1, BiLSTM layer 0, layer 1   :59-60
    P_DENSE2,                              // Decoder dense2 + tanh                     :61-62
    P_MOB_R, P_MOB_X,                      // MobiusLinear matmul on eucl / on the window (same panel)
    P_COUNT
};
static_assert(P_COUNT <= kMaxPasses, "layer program too long");

struct NetProgram {
    PassDesc pass[kMaxPasses];
    int32_t S;             // signal_shape
    int32_t S8;            // S rounded up to 8
    int32_t latent;        // Encoder/Decoder latent_space_dim
    int32_t latent_c;      // CriticX latent_space_dim
    int32_t hyperbolic;
    int32_t mob_bias_off;  // float offset of the Mobius bias [128] (zero padded)
    int32_t mob_y2_off;    // float offset of sum(bias^2) (fp32)
    int32_t critic5_off;   // float offset of critic dense5: W[latent_c] then b[1]
};

// One packed output column: where its weight row and its biases come from (weight packing kernels).
struct ColSrc {
    const float* w = nullptr;   // row-major (rows, K) source matrix or nullptr (zero column)
    const float* b1 = nullptr;  // bias source or nullptr
    const float* b2 = nullptr;
    int32_t row = 0;            // weight row
    int32_t K = 0;              // source row length
    int32_t bidx = 0;           // bias index
    int32_t k0 = 0;             // operand feature of the source row's element 0 (tensor-core packing: the critic chain reads
                                // its hidden state from features 104.. of the shared operand buffer)
};

}  // namespace hypad

struct hypad_ctx {
    int device;
    float* packed;        // device: packed panels + biases
    size_t packed_floats;
    hypad::NetProgram prog;
    bool has_weights;
    void* workspace;      // device scratch, grown on demand
    size_t workspace_bytes;
    int max_smem_optin;
    // tensor-core path (forward_tc.cu)
    unsigned char* tc_packed;   // TF32-split weight stages followed by the small-parameter block
    size_t tc_bytes;
    size_t tc_small_off;
    int* tc_error;              // device flags: 1 barrier wait timed out, 2 operand range left, 4 range fallback served the call
    bool strict_range;          // hypad_forward raises on a range violation instead of falling back to the FFMA kernel
    long long range_fallbacks;  // polls that found a call served by the fallback
    long long* tc_debug;        // optional device cycle counters (hypad_forward_debug_cycles)
    unsigned char tc_prog_storage[2048];
    void* fin_state;            // device: state of the staged statistics (critic_stats.cu), allocated on first use
};

namespace hypad {
int ensure_workspace(hypad_ctx* ctx, size_t bytes);
}
