// Shared device helpers of the fusion kernels (sm_100a).  All translation units are compiled with
// -fmad=false: every fp32/fp64 expression below is evaluated exactly as written (one IEEE rounding
// per operator), which is what makes the results bit-comparable with NumPy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/sd_fusion.h"

#define SD_WARP 32
#define SD_FULL 0xffffffffu

#define SD_CUDA_TRY(expr)                                   \
    do {                                                    \
        cudaError_t _e = (expr);                            \
        if (_e != cudaSuccess) {                            \
            sd_set_last_cuda_error((int)_e, #expr);         \
            return SD_ERR_CUDA;                             \
        }                                                   \
    } while (0)

#define SD_LAUNCH_CHECK()  SD_CUDA_TRY(cudaGetLastError())

void sd_set_last_cuda_error(int code, const char* what);

namespace sd {

// ------------------------------------------------------------------------------------------------
// order-preserving fp32 <-> uint32 key (radix select, atomic min/max)
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f2key(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b; memcpy(&b, &f, 4);
#endif
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__host__ __device__ __forceinline__ float key2f(uint32_t k) {
    uint32_t b = k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu);
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}

// ------------------------------------------------------------------------------------------------
// warp / block primitives
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(SD_FULL, v, o);
        if (lane_id() >= o) v += t;
    }
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SD_FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SD_FULL, v, o);
    return v;
}
__device__ __forceinline__ uint32_t warp_min(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(SD_FULL, v, o));
    return v;
}
__device__ __forceinline__ uint32_t warp_max(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(SD_FULL, v, o));
    return v;
}

// Exclusive block scan of one int per thread (blockDim.x <= 1024, multiple of 32).
// Returns the exclusive prefix; *total receives the block sum.  `smem` needs 33 ints.
__device__ __forceinline__ int block_excl_scan(int v, int* smem, int* total) {
    int incl = warp_incl_scan(v);
    int w = warp_id(), l = lane_id(), nw = blockDim.x >> 5;
    if (l == 31) smem[w] = incl;
    __syncthreads();
    if (w == 0) {
        int s = (l < nw) ? smem[l] : 0;
        int si = warp_incl_scan(s);
        smem[l] = si - s;
        if (l == 31) smem[32] = si;
    }
    __syncthreads();
    int excl = incl - v + smem[w];
    *total = smem[32];
    __syncthreads();
    return excl;
}

// ------------------------------------------------------------------------------------------------
// Decoupled look-back (single-pass chained scan) over dynamically ticketed tiles.
// One 64-bit status word per tile: [63:62] flag (0 = empty, 1 = aggregate, 2 = inclusive prefix),
// [61:0] payload.  Words are self-cleaned by the last block of the launch, so a workspace that was
// zeroed once can be reused by every later launch (CUDA-graph friendly, no memset nodes).
// ------------------------------------------------------------------------------------------------
#define SD_FLAG_AGG  (1ull << 62)
#define SD_FLAG_INCL (2ull << 62)
#define SD_PAYLOAD   ((1ull << 62) - 1)

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// Called by ONE warp of the block that owns `tile` with that tile's aggregate payload (payloads add
// component-wise without carry between packed fields by construction).  Publishes the aggregate,
// walks back over the predecessors' words and returns the exclusive prefix (sum of all earlier
// tiles' aggregates); finally publishes the inclusive prefix.
__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long* status, int tile,
                                                                unsigned long long aggregate) {
    int l = lane_id();
    if (tile == 0) {
        if (l == 0) st_status(status, SD_FLAG_INCL | aggregate);
        return 0ull;
    }
    if (l == 0) st_status(status + tile, SD_FLAG_AGG | aggregate);
    unsigned long long excl = 0ull;
    int base = tile - 1;
    while (true) {
        int idx = base - l;
        unsigned long long w = 0ull;
        if (idx >= 0) {
            do { w = ld_status(status + idx); } while ((w >> 62) == 0ull);
        } else {
            w = SD_FLAG_INCL;  // virtual tile before the first: inclusive prefix 0
        }
        unsigned incl_mask = __ballot_sync(SD_FULL, (w >> 62) == 2ull);
        // lanes up to and including the first (closest) inclusive word contribute
        int first = incl_mask ? (__ffs(incl_mask) - 1) : 32;
        unsigned long long contrib = (l <= first) ? (w & SD_PAYLOAD) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(SD_FULL, contrib, o);
        excl += contrib;
        if (incl_mask) break;
        base -= 32;
    }
    if (l == 0) st_status(status + tile, SD_FLAG_INCL | (excl + aggregate));
    return excl;
}

// Per-scan bookkeeping that lives next to the status words.
struct ScanCtl {
    unsigned int ticket;   // next tile id
    unsigned int done;     // blocks that have left the kernel
    unsigned int aux0;     // kernel-specific accumulator (self-cleaned by its user)
    unsigned int aux1;
};

// Last-block cleanup: every block calls this exactly once when it has no more tiles.  The block that
// observes done == gridDim.x - 1 is the last: all other blocks have finished every status read, so
// it may zero the words in [0, ntiles) and the counters.  Returns true in the last block.
__device__ __forceinline__ bool scan_finish(ScanCtl* ctl, unsigned long long* status, int ntiles, int nblocks) {
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int d = atomicAdd(&ctl->done, 1u);
        s_last = (d == (unsigned)nblocks - 1u);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    for (int i = threadIdx.x; i < ntiles; i += blockDim.x) status[i] = 0ull;
    if (threadIdx.x == 0) { ctl->ticket = 0u; ctl->done = 0u; }
    return true;
}

// ------------------------------------------------------------------------------------------------
// misc
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ldg_f(const float* p) { return __ldg(p); }

template <typename T>
__host__ __device__ __forceinline__ T ceil_div(T a, T b) { return (a + b - 1) / b; }

}  // namespace sd
