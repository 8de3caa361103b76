// Unevaluated sums of two doubles ("double-double"): ~106-bit accumulation for the O(T) statistics of the finish.
// Why: every global statistic of the path (critic mean / std, z-score moments, prefix sums of the smoothing) is a sum over
// up to millions of positions that several GPUs each hold a slice of.  Accumulated in (hi, lo) pairs with error-free
// transformations the result no longer depends on the order of the partial sums to ~1e-30 relative, so the rounded double is
// the same whether one GPU or eight took part (the sharded runs are compared bit for bit with the single-GPU one), and flat
// stretches of a signal keep exactly flat statistics (see finish.cu).
#pragma once

namespace hypad {

struct dd {
    double hi, lo;
};

__host__ __device__ __forceinline__ dd dd_make(double hi, double lo = 0.0) {
    dd r;
    r.hi = hi;
    r.lo = lo;
    return r;
}

#ifdef __CUDA_ARCH__
#define HYPAD_DADD(a, b) __dadd_rn((a), (b))
#define HYPAD_DMUL(a, b) __dmul_rn((a), (b))
#define HYPAD_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define HYPAD_DADD(a, b) ((a) + (b))
#define HYPAD_DMUL(a, b) ((a) * (b))
#define HYPAD_FMA(a, b, c) fma((a), (b), (c))
#endif

__host__ __device__ __forceinline__ dd dd_add(dd a, dd b) {
    const double s = HYPAD_DADD(a.hi, b.hi), bb = HYPAD_DADD(s, -a.hi);
    double e = HYPAD_DADD(HYPAD_DADD(a.hi, -HYPAD_DADD(s, -bb)), HYPAD_DADD(b.hi, -bb));  // TwoSum error term
    e = HYPAD_DADD(e, HYPAD_DADD(a.lo, b.lo));
    dd r;
    r.hi = HYPAD_DADD(s, e);
    r.lo = HYPAD_DADD(e, -HYPAD_DADD(r.hi, -s));
    return r;
}
__host__ __device__ __forceinline__ dd dd_add(dd a, double b) { return dd_add(a, dd_make(b)); }
__host__ __device__ __forceinline__ dd dd_neg(dd a) { return dd_make(-a.hi, -a.lo); }
// a * b exactly (TwoProd through the fused multiply-add)
__host__ __device__ __forceinline__ dd dd_prod(double a, double b) {
    dd r;
    r.hi = HYPAD_DMUL(a, b);
    r.lo = HYPAD_FMA(a, b, -r.hi);
    return r;
}
__host__ __device__ __forceinline__ dd dd_mul(dd a, dd b) {
    dd p = dd_prod(a.hi, b.hi);
    p.lo = HYPAD_DADD(p.lo, HYPAD_DADD(HYPAD_DMUL(a.hi, b.lo), HYPAD_DMUL(a.lo, b.hi)));
    dd r;
    r.hi = HYPAD_DADD(p.hi, p.lo);
    r.lo = HYPAD_DADD(p.lo, -HYPAD_DADD(r.hi, -p.hi));
    return r;
}
__host__ __device__ __forceinline__ dd dd_div(dd a, double d) {
    const double q1 = a.hi / d;
    const dd r = dd_add(a, dd_neg(dd_prod(q1, d)));
    const double q2 = r.hi / d;
    dd q;
    q.hi = HYPAD_DADD(q1, q2);
    q.lo = HYPAD_DADD(q2, -HYPAD_DADD(q.hi, -q1));
    return q;
}
__host__ __device__ __forceinline__ double dd_value(dd a) { return HYPAD_DADD(a.hi, a.lo); }

#ifdef __CUDACC__
__device__ __forceinline__ dd dd_shfl_xor(dd v, int o) {
    dd r;
    r.hi = __shfl_xor_sync(0xffffffffu, v.hi, o);
    r.lo = __shfl_xor_sync(0xffffffffu, v.lo, o);
    return r;
}
// Sum over the CTA (blockDim.x a multiple of 32, <= 1024); sh holds 64 doubles.  Every thread gets the result.
__device__ __forceinline__ dd dd_block_sum(dd v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dd_add(v, dd_shfl_xor(v, o));
    if (lane == 0) {
        sh[warp] = v.hi;
        sh[32 + warp] = v.lo;
    }
    __syncthreads();
    dd t = dd_make(0.0);
    if (warp == 0) {
        if (lane < (int)(blockDim.x >> 5)) t = dd_make(sh[lane], sh[32 + lane]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = dd_add(t, dd_shfl_xor(t, o));
        if (lane == 0) {
            sh[0] = t.hi;
            sh[32] = t.lo;
        }
    }
    __syncthreads();
    t = dd_make(sh[0], sh[32]);
    __syncthreads();
    return t;
}
#endif

}  // namespace hypad
