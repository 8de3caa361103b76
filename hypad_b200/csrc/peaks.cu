// Pipe-rate micro-benchmarks for the roofline denominators SURVEY.md 8(d) asks for: FP32 FFMA, FP64 DFMA, MUFU.EX2
// (ex2.approx.ftz.f32), MUFU.EX2 in its packed half form (ex2.approx.f16x2: two results per issue), and the 64-bit shuffle
// the KDE kernel's hand-over uses.  Each thread runs ILP independent dependency chains of `iters` instructions; the grid fills
// every SM with `ctas_per_sm` CTAs of 1024 threads.  The caller times the launch with CUDA events (scripts/measure_peaks.py)
// and writes profiles/peaks.json; bench.py reads that file for the KDE kernel's roofline.
#include <cuda_fp16.h>

#include "common.cuh"

namespace hypad {

constexpr int ILP = 8;

template <int KIND>
__global__ void __launch_bounds__(1024) peak_kernel(int iters, float seed, float* sink) {
    float acc = 0.f;
    if (KIND == 0) {  // FFMA
        float v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) v[j] = seed + (float)(threadIdx.x + j);
        const float a = 0.999f + seed, b = 1e-3f;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) v[j] = fmaf(v[j], a, b);
        }
#pragma unroll
        for (int j = 0; j < ILP; ++j) acc += v[j];
    } else if (KIND == 1) {  // DFMA
        double v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) v[j] = (double)seed + (double)(threadIdx.x + j);
        const double a = 0.999 + (double)seed, b = 1e-3;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) v[j] = fma(v[j], a, b);
        }
#pragma unroll
        for (int j = 0; j < ILP; ++j) acc += (float)v[j];
    } else if (KIND == 2) {  // MUFU.EX2 fp32
        float v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) v[j] = seed - 0.5f - 1e-3f * (float)((threadIdx.x + j) & 31);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[j]));  // 2^x of a value in (0, 1) stays there after the first step
        }
#pragma unroll
        for (int j = 0; j < ILP; ++j) acc += v[j];
    } else if (KIND == 3) {  // MUFU.EX2 f16x2
        unsigned v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            const __half2 h = __floats2half2_rn(seed - 0.5f - 1e-3f * (float)j, seed - 0.25f);
            v[j] = *reinterpret_cast<const unsigned*>(&h);
        }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[j]));
        }
#pragma unroll
        for (int j = 0; j < ILP; ++j) acc += __half2float(__low2half(*reinterpret_cast<const __half2*>(&v[j])));
    } else {  // SHFL.IDX of 32-bit registers
        float v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) v[j] = seed + (float)(threadIdx.x + j);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) v[j] = __shfl_sync(0xffffffffu, v[j], (i + j) & 31);
        }
#pragma unroll
        for (int j = 0; j < ILP; ++j) acc += v[j];
    }
    if (acc == 123456.789f) *sink = acc;  // keeps the chains alive
}

}  // namespace hypad

using namespace hypad;

extern "C" int hypad_peak_probe(int kind, int iters, int ctas_per_sm, float* sink, long long* instr_per_launch, void* stream_) {
    HYPAD_REQUIRE(kind >= 0 && kind <= 4 && iters > 0 && ctas_per_sm > 0 && sink, "hypad_peak_probe: bad argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    int dev = 0, sms = kNumSMs;
    HYPAD_CUDA_TRY(cudaGetDevice(&dev));
    HYPAD_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned grid = (unsigned)(sms * ctas_per_sm);
    switch (kind) {
        case 0: peak_kernel<0><<<grid, 1024, 0, stream>>>(iters, 0.f, sink); break;
        case 1: peak_kernel<1><<<grid, 1024, 0, stream>>>(iters, 0.f, sink); break;
        case 2: peak_kernel<2><<<grid, 1024, 0, stream>>>(iters, 0.f, sink); break;
        case 3: peak_kernel<3><<<grid, 1024, 0, stream>>>(iters, 0.f, sink); break;
        default: peak_kernel<4><<<grid, 1024, 0, stream>>>(iters, 0.f, sink); break;
    }
    HYPAD_LAUNCH_CHECK();
    if (instr_per_launch) *instr_per_launch = (long long)grid * 1024ll * (long long)iters * ILP;  // thread-level instructions
    return HYPAD_OK;
}
