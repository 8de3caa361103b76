// Diagnostic: one 128 x N x K product on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulator in TMEM),
// with the fp32 operands split into 1, 2 or 3 TF32 pieces ("3xTF32" error-compensated product and a 6-term variant).
// It exists to MEASURE, on the B200, whether a tensor-core contraction can hold the score parity of the fp32 path
// (tests/test_gpu_tensor_probe.py compares it with an fp64 product and with the FFMA order the fused kernel uses).
//
// Operand layout in shared memory (K-major, no swizzle; UMMA "interleave" canonical layout):
//   element (row r, k) of a piece lives at ((k/4) * rows + r) * 16 B + (k%4) * 4 B
// i.e. core matrices of 8 rows x 16 B, SBO (next 8 rows) = 128 B, LBO (next 4 k) = rows * 16 B.
#include "common.cuh"

namespace hypad {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version for sm_100
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, both K-major, N>>3, M>>4
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                    // c_format = F32
    d |= 2u << 7;                    // a_format = TF32
    d |= 2u << 10;                   // b_format = TF32
    d |= (uint32_t)(N >> 3) << 17;   // n_dim
    d |= (uint32_t)(M >> 4) << 24;   // m_dim
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

struct TcProbeArgs {
    const float* A;  // [128][K]
    const float* B;  // [N][K]
    float* D;        // [128][N]
    int K, N, pieces, terms;
};

// pieces: how many TF32 pieces each operand is split into (1..3); terms: how many partial products are accumulated
//   terms 1: a1 b1 ; 3: a2 b1 + a1 b2 + a1 b1 ; 6: a3 b1 + a2 b2 + a1 b3 + a2 b1 + a1 b2 + a1 b1   (small terms first)
__global__ void __launch_bounds__(128) tc_probe_kernel(const TcProbeArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = a.K, N = a.N;
    const int a_piece = K * 128;  // floats per A piece: [K/4][128][4]
    const int b_piece = K * N;    // floats per B piece: [K/4][N][4]
    float* sA = smem;
    float* sB = smem + a.pieces * a_piece;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::);
    }
    // split and lay out the operands
    for (int e = tid; e < 128 * K; e += 128) {
        const int r = e / K, k = e - r * K;
        float v = a.A[e];
        const int off = ((k >> 2) * 128 + r) * 4 + (k & 3);
        for (int p = 0; p < a.pieces; ++p) {
            const float h = to_tf32_rna(v);
            sA[p * a_piece + off] = h;
            v -= h;
        }
    }
    for (int e = tid; e < N * K; e += 128) {
        const int r = e / K, k = e - r * K;
        float v = a.B[e];
        const int off = ((k >> 2) * N + r) * 4 + (k & 3);
        for (int p = 0; p < a.pieces; ++p) {
            const float h = to_tf32_rna(v);
            sB[p * b_piece + off] = h;
            v -= h;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_slot;

    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(128, N);
        // (a piece, b piece) per term, small products first
        const int pa6[6] = {2, 1, 0, 1, 0, 0}, pb6[6] = {0, 1, 2, 0, 1, 0};
        const int pa3[3] = {1, 0, 0}, pb3[3] = {0, 1, 0};
        uint32_t acc = 0;
        for (int t = 0; t < a.terms; ++t) {
            const int ia = a.terms == 6 ? pa6[t] : (a.terms == 3 ? pa3[t] : 0);
            const int ib = a.terms == 6 ? pb6[t] : (a.terms == 3 ? pb3[t] : 0);
            for (int ks = 0; ks < K / 8; ++ks) {
                // one MMA consumes 8 k = two 16-byte K chunks; chunk stride (LBO) = rows * 16 B, 8-row group stride (SBO) = 128 B
                const uint64_t ad = make_smem_desc(smem_u32(sA + ia * a_piece) + ks * 2 * 128 * 16, 128 * 16, 128);
                const uint64_t bd = make_smem_desc(smem_u32(sB + ib * b_piece) + ks * 2 * N * 16, N * 16, 128);
                umma_tf32(tmem, ad, bd, idesc, acc);
                acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the MMAs to retire (phase 0)
    {
        uint32_t done = 0;
        for (long long spin = 0; !done && spin < (1ll << 24); ++spin) {  // bounded: a bad descriptor must not hang the GPU
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(smem_u32(&bar)), "r"(0u)
                : "memory");
        }
        if (!done && tid == 0) a.D[0] = __int_as_float(0x7fc00001);  // never completed: flag with a NaN payload
        if (!done) {
            __syncthreads();
            if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
            return;
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // TMEM -> registers -> global: thread = row (lane of TMEM), 8 columns per load
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j) a.D[row * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
}

// Throughput microbenchmark: `reps` tcgen05.mma (M=128, K=8, tf32) issued back to back by one thread, operands in shared
// memory (contents irrelevant).  mode 0: every MMA accumulates into the same TMEM columns; mode 1: rotate over
// 512/N accumulators; mode 2: like 0 but A and B descriptors alternate between two buffers.  out[0] = cycles.
__global__ void __launch_bounds__(128) tc_bench_kernel(int N, int reps, int mode, long long* out) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t bar;
    __shared__ uint64_t bar2[2];
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 2 * (128 * 8 + 256 * 8); e += 128) smem[e] = 1.0f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_slot;
    if (warp == 1) {
        const uint32_t idesc = make_idesc_tf32(128, N);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 2 * 128 * 8);
        const uint32_t a1 = a0 + 128 * 8 * 4, b1 = b0 + 256 * 8 * 4;
        const int nacc = 512 / N;
        long long t0 = 0, t1 = 0;
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
        if (pred) {
            t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                const bool alt = (mode == 2) && (r & 1);
                const uint64_t ad = make_smem_desc(alt ? a1 : a0, 128 * 16, 128);
                const uint64_t bd = make_smem_desc(alt ? b1 : b0, N * 16, 128);
                const uint32_t d = tmem + (mode == 1 ? (uint32_t)((r % nacc) * N) : 0u);
                umma_tf32(d, ad, bd, idesc, 1);
                if (mode >= 3 && (r % 6) == 5) {
                    // mode 3: a commit (to a spare barrier) after every 6 MMAs; mode 4: plus a try_wait that is already satisfied
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[0])) : "memory");
                    if (mode >= 4) {
                        uint32_t ok;
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                                     : "=r"(ok) : "r"(smem_u32(&bar2[1])), "r"(1u) : "memory");
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                }
            }
            t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            out[1] = t1 - t0;  // issue time
        }
        __syncwarp();
        uint32_t done = 0;
        for (long long spin = 0; !done && spin < (1ll << 26); ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        if (pred) out[0] = clock64() - t0;  // until retired
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

}  // namespace hypad

using namespace hypad;

extern "C" int hypad_tc_probe_bench(int N, int reps, int mode, long long* h_out2, void* stream) {
    HYPAD_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && reps >= 1 && h_out2, "hypad_tc_probe_bench: bad argument");
    long long* d = nullptr;
    HYPAD_CUDA_TRY(cudaMalloc(&d, 2 * sizeof(long long)));
    const size_t smem = 2 * (128 * 8 + 256 * 8) * sizeof(float);
    tc_bench_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(N, reps, mode, d);
    HYPAD_LAUNCH_CHECK();
    HYPAD_CUDA_TRY(cudaMemcpy(h_out2, d, 2 * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(d);
    return HYPAD_OK;
}


extern "C" int hypad_tc_probe_gemm(const float* A, const float* B, float* D, int K, int N, int pieces, int terms, void* stream) {
    HYPAD_REQUIRE(A && B && D, "hypad_tc_probe_gemm: NULL argument");
    HYPAD_REQUIRE(K >= 8 && K % 8 == 0 && N >= 16 && N % 16 == 0 && N <= 256, "hypad_tc_probe_gemm: need K%%8==0, 16<=N<=256, N%%16==0");
    HYPAD_REQUIRE((pieces == 1 && terms == 1) || (pieces == 2 && terms == 3) || (pieces == 3 && terms == 6),
                  "hypad_tc_probe_gemm: (pieces, terms) must be (1,1), (2,3) or (3,6)");
    const size_t smem = (size_t)pieces * (K * 128 + K * N) * sizeof(float);
    HYPAD_REQUIRE(smem <= 220 * 1024, "hypad_tc_probe_gemm: operands need %zu bytes of shared memory", smem);
    HYPAD_CUDA_TRY(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TcProbeArgs a{A, B, D, K, N, pieces, terms};
    tc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}
