// All-gather of a small per-rank record through NVLink peer memory, in ONE kernel launch and without a host-side collective call.
// Every rank owns a symmetric buffer (torch.distributed._symmetric_memory: the same allocation mapped into every process of the
// node); the kernel's CTA p copies this rank's record straight into peer p's buffer (remote stores over NVLink / NVSwitch), makes
// them visible system-wide, raises this rank's flag in peer p's buffer, and then waits for peer p's flag in the LOCAL buffer.
// When the kernel retires, the records of all ranks sit back to back in local memory for the next kernel of the stream.
// The stages of the sharded finish exchange a few KB each (digit histograms, partial sums, block summaries, run fragments): an
// NCCL all-gather of that size costs ~20-40 us of launch and protocol latency, this costs one launch and one NVLink round trip.
//
// Protocol: exchanges are numbered; exchange e uses slot e % slots with sequence number e / slots + 1.  A rank can only push
// exchange e + 1 after its wait of exchange e returned, i.e. after every peer pushed e, which every peer does (stream order)
// after its consumers of e - 1 ran: with >= 2 slots a slot is never overwritten while a peer still reads it.
#include "common.cuh"

namespace hypad {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) peer_exchange_kernel(const unsigned char* __restrict__ local, long long nbytes,
                                                            const long long* __restrict__ peer_base, int rank, int world, long long data_off,
                                                            long long flag_off, unsigned long long seq, int* error_flag) {
    const int p = blockIdx.x;  // the peer this CTA serves
    unsigned char* remote = reinterpret_cast<unsigned char*>(peer_base[p]);
    unsigned char* dst = remote + data_off + (long long)rank * nbytes;
    if ((nbytes & 15) == 0 && ((reinterpret_cast<uintptr_t>(local) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        const uint4* s = reinterpret_cast<const uint4*>(local);
        uint4* d = reinterpret_cast<uint4*>(dst);
        for (long long i = threadIdx.x; i < nbytes / 16; i += blockDim.x) d[i] = s[i];
    } else {
        const unsigned long long* s = reinterpret_cast<const unsigned long long*>(local);
        unsigned long long* d = reinterpret_cast<unsigned long long*>(dst);
        for (long long i = threadIdx.x; i < nbytes / 8; i += blockDim.x) d[i] = s[i];
    }
    __threadfence_system();  // every thread's stores are visible system-wide before the flag goes up
    __syncthreads();
    if (threadIdx.x == 0) {
        st_release_sys(reinterpret_cast<unsigned long long*>(remote + flag_off) + rank, seq);
        // wait for peer p's record in the local buffer (bounded: a lost peer must not hang this GPU)
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(reinterpret_cast<unsigned char*>(peer_base[rank]) + flag_off) + p;
        const long long t0 = clock64();
        while (ld_acquire_sys(mine) < seq) {
            if (clock64() - t0 > 20000000000ll) {  // ~10 s at 2 GHz
                if (error_flag) atomicExch(error_flag, 1);
                break;
            }
            __nanosleep(200);
        }
    }
}

}  // namespace hypad

extern "C" int hypad_peer_exchange(const void* local, int64_t nbytes, const int64_t* peer_base_dev, int rank, int world, int64_t data_off,
                                   int64_t flag_off, uint64_t seq, int* error_flag_dev, void* stream) {
    HYPAD_REQUIRE(local && peer_base_dev && nbytes > 0 && (nbytes & 7) == 0 && world >= 1 && rank >= 0 && rank < world,
                  "hypad_peer_exchange: bad argument (records are multiples of 8 bytes)");
    hypad::peer_exchange_kernel<<<(unsigned)world, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)local, (long long)nbytes,
                                                                                  (const long long*)peer_base_dev, rank, world,
                                                                                  (long long)data_off, (long long)flag_off,
                                                                                  (unsigned long long)seq, error_flag_dev);
    HYPAD_LAUNCH_CHECK();
    return HYPAD_OK;
}
