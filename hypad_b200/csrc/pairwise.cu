// hyperspace/poincare_distance.py of the reference on the device (off the executed scoring path, used by
// hyperspace/losses.py:154):
//   square_norm(x)          :19-25   clamp(torch.norm(x, dim=-1)^2, min=1e-5)
//   pairwise_distances(x,y) :28-48   clamp(|x_i|^2 + |y_j|^2 - 2 <x_i, y_j>, 1e-7, inf)
//   poincare_distance(p,g)  :5-16    acosh(1 + 2 pairwise / ((1 - square_norm(p))_i (1 - square_norm(g))_j))
// fp32 like the reference (torch.mm on fp32 inputs).  The N x M x D contraction runs on the fp32 FFMA pipe: the squared
// distance is a difference of nearly equal numbers for close points, so the products must keep fp32 accuracy, and D is ~100
// -- the kernel is bound by the FFMA pipe (200 FLOP per pair) with the N x M fp32 output (4 B per pair) close behind; see DESIGN.md.
#include "common.cuh"

namespace hypad {

constexpr int PW_TILE = 128;  // rows of x and rows of y per CTA tile
constexpr int PW_K = 16;      // features per shared-memory stage

// per row: sum of squares (fp32 value of the fp64 sum) and the clamped squared norm of square_norm()
__global__ void pw_rownorm_kernel(const float* __restrict__ x, int64_t n, int D, float* __restrict__ sumsq, float* __restrict__ sqnorm) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
        double s = 0.0;
        for (int c = lane; c < D; c += 32) {
            const float v = x[row * D + c];
            s += (double)__fmul_rn(v, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            const float ss = (float)s;
            if (sumsq) sumsq[row] = ss;
            if (sqnorm) {
                const float r = sqrtf(ss);  // torch.norm(...) ** 2
                sqnorm[row] = fmaxf(__fmul_rn(r, r), 1e-5f);
            }
        }
    }
}

// One output from its dot product, in the reference's operation order.  Not inlined: the kernel calls it for each of a thread's 64
// outputs, and 64 inlined copies of acoshf pushed the epilogue out of the instruction cache (ncu: `no_instruction` was the top
// stall of the first version).
__device__ __noinline__ float pw_finish(float dot, float xn, float yn, float one_minus_xsq, float one_minus_ysq, int mode) {
    // x_norm + y_norm - 2.0 * mm, clamped to [1e-7, inf)   (:45-48)
    float v = __fsub_rn(__fadd_rn(xn, yn), __fmul_rn(2.0f, dot));
    v = fmaxf(v, 1e-7f);
    if (mode == 0) v = acoshf(__fadd_rn(1.0f, __fdiv_rn(__fmul_rn(2.0f, v), __fmul_rn(one_minus_xsq, one_minus_ysq))));  // :16
    return v;
}

// mode 0: poincare_distance, mode 1: pairwise_distances.  One CTA per 128 x 128 output tile (grid-stride), 256 threads, thread =
// 8 x 8 outputs (rows ty*4 + {0..3} and 64 + ty*4 + {0..3}, columns likewise with tx: every shared-memory read is a conflict-free
// LDS.128, 4 of them per 64 FFMA, and 16 neighbouring threads store 64 consecutive floats); operands staged k-major in shared
// memory 16 features at a time.
__global__ void __launch_bounds__(256, 2) pw_distance_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                             const float* __restrict__ xss, const float* __restrict__ yss,
                                                             const float* __restrict__ xsq, const float* __restrict__ ysq, int64_t N,
                                                             int64_t M, int D, int mode, float* __restrict__ out) {
    __shared__ __align__(16) float As[PW_K][PW_TILE + 4];
    __shared__ __align__(16) float Bs[PW_K][PW_TILE + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int lr = threadIdx.x >> 1, lk = (threadIdx.x & 1) * 8;  // staging: row lr, features lk .. lk+7 of the stage
    const int64_t tiles_m = (M + PW_TILE - 1) / PW_TILE, tiles_n = (N + PW_TILE - 1) / PW_TILE;
    for (int64_t tile = blockIdx.x; tile < tiles_n * tiles_m; tile += gridDim.x) {
        const int64_t i0 = (tile / tiles_m) * PW_TILE, j0 = (tile % tiles_m) * PW_TILE;
        float acc[8][8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.0f;
        const bool xr = i0 + lr < N, yr = j0 + lr < M;
        const float* xrow = x + (i0 + lr) * D;
        const float* yrow = y + (j0 + lr) * D;
        for (int k0 = 0; k0 < D; k0 += PW_K) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int k = k0 + lk + q;
                As[lk + q][lr] = (xr && k < D) ? xrow[k] : 0.0f;
                Bs[lk + q][lr] = (yr && k < D) ? yrow[k] : 0.0f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < PW_K; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
            }
            __syncthreads();
        }
        const bool vec = (M & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;  // rows of out are 16-byte aligned
        // the 8 columns' norms once per tile (clamped index: lanes beyond M compute a value nobody stores)
        float yn[8], yq[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int64_t j = j0 + (c >> 2) * 64 + tx * 4 + (c & 3);
            const int64_t jc = j < M ? j : M - 1;
            yn[c] = yss[jc];
            yq[c] = mode == 0 ? __fsub_rn(1.0f, ysq[jc]) : 0.0f;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int64_t i = i0 + (r >> 2) * 64 + ty * 4 + (r & 3);
            if (i >= N) continue;
            const float xn = xss[i];
            const float a1 = mode == 0 ? __fsub_rn(1.0f, xsq[i]) : 0.0f;
#pragma unroll
            for (int cg = 0; cg < 2; ++cg) {
                const int64_t jb = j0 + cg * 64 + tx * 4;
                float d[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) d[c] = pw_finish(acc[r][cg * 4 + c], xn, yn[cg * 4 + c], a1, yq[cg * 4 + c], mode);
                if (vec && jb + 3 < M) {
                    *reinterpret_cast<float4*>(out + i * M + jb) = make_float4(d[0], d[1], d[2], d[3]);
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (jb + c < M) out[i * M + jb + c] = d[c];
                }
            }
        }
    }
}

static int pairwise(hypad_ctx* ctx, const float* x, int64_t n, const float* y, int64_t m, int D, int mode, float* out, cudaStream_t stream) {
    HYPAD_CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_workspace(ctx, (size_t)(2 * (n + m)) * sizeof(float));
    if (rc != HYPAD_OK) return rc;
    float* xss = (float*)ctx->workspace;
    float* xsq = xss + n;
    float* yss = xsq + n;
    float* ysq = yss + m;
    auto rows_grid = [](int64_t rows) { return (unsigned)(ceil_div(rows, 8) < kNumSMs * 8 ? ceil_div(rows, 8) : kNumSMs * 8); };
    pw_rownorm_kernel<<<rows_grid(n), 256, 0, stream>>>(x, n, D, xss, xsq);
    HYPAD_LAUNCH_CHECK();
    if (y != x || m != n) {
        pw_rownorm_kernel<<<rows_grid(m), 256, 0, stream>>>(y, m, D, yss, ysq);
        HYPAD_LAUNCH_CHECK();
    } else {
        yss = xss, ysq = xsq;
    }
    const int64_t tiles = ceil_div(n, PW_TILE) * ceil_div(m, PW_TILE);
    const unsigned grid = (unsigned)(tiles < (int64_t)kNumSMs * 16 ? tiles : (int64_t)kNumSMs * 16);
    pw_distance_kernel<<<grid, 256, 0, stream>>>(x, y, xss, yss, xsq, ysq, n, m, D, mode, out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

}  // namespace hypad

using namespace hypad;

extern "C" {

int hypad_poincare_distance_pairwise(hypad_ctx* ctx, const float* pred, int64_t n_pred, const float* gt, int64_t n_gt, int D, float* out,
                                     void* stream) {
    HYPAD_REQUIRE(ctx && pred && gt && out, "hypad_poincare_distance_pairwise: NULL argument");
    HYPAD_REQUIRE(n_pred >= 0 && n_gt >= 0 && D >= 1, "hypad_poincare_distance_pairwise: bad shape");
    if (n_pred == 0 || n_gt == 0) return HYPAD_OK;
    return pairwise(ctx, pred, n_pred, gt, n_gt, D, 0, out, (cudaStream_t)stream);
}

int hypad_pairwise_sqdist(hypad_ctx* ctx, const float* x, int64_t n, const float* y, int64_t m, int D, float* out, void* stream) {
    HYPAD_REQUIRE(ctx && x && y && out, "hypad_pairwise_sqdist: NULL argument");
    HYPAD_REQUIRE(n >= 0 && m >= 0 && D >= 1, "hypad_pairwise_sqdist: bad shape");
    if (n == 0 || m == 0) return HYPAD_OK;
    return pairwise(ctx, x, n, y, m, D, 1, out, (cudaStream_t)stream);
}

int hypad_square_norm(const float* x, int64_t n, int D, float* out, void* stream) {
    HYPAD_REQUIRE(x && out, "hypad_square_norm: NULL argument");
    HYPAD_REQUIRE(n >= 0 && D >= 1, "hypad_square_norm: bad shape");
    if (n == 0) return HYPAD_OK;
    const unsigned grid = (unsigned)(ceil_div(n, 8) < kNumSMs * 8 ? ceil_div(n, 8) : kNumSMs * 8);
    pw_rownorm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, D, nullptr, out);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}

}  // extern "C"
