// Exact medians on fp32 columns: np.median of pcl.mad (pcl.py:76-81).
//
// Three-pass MSD radix select (11 + 11 + 10 bits) on order-preserving keys.  Both middle order
// statistics are tracked at once (even n needs fl32((a+b)/2)).  A pass is one launch
// over all jobs: CTAs build a shared-memory histogram of the current digit of the keys that match
// the resolved prefix, merge it into a per-job global histogram with integer atomics (deterministic)
// and the last CTA of the job (ticket) picks the bucket holding each rank, then zeroes the
// histogram for the next pass.  Keys are either col[i] or |col[i] - *center| (the MAD pass), so the
// absolute deviations are never materialised.  The column is L2-resident (<= 8 MB), so the three
// reads cost L2 bandwidth, not HBM.
#include "sd_internal.cuh"

namespace sd {

#ifndef SD_SEL_THREADS
#define SD_SEL_THREADS 512
#endif
#ifndef SD_SEL_WAVES
#define SD_SEL_WAVES 2      // CTAs per SM over all jobs of a launch (per pass alone: 1 -> 28 us, 2 -> 18 us, 4 -> 15 us; the pipelined step is the same for 2 and 4)
#endif
constexpr int kSelThreads = SD_SEL_THREADS;
constexpr int kSelItems = 8;

template <int PASS> struct SelPass;
template <> struct SelPass<0> { static constexpr int shift = 21, bits = 11, hi_shift = 32; };
template <> struct SelPass<1> { static constexpr int shift = 10, bits = 11, hi_shift = 21; };
template <> struct SelPass<2> { static constexpr int shift = 0,  bits = 10, hi_shift = 10; };

__device__ __forceinline__ float sel_val(const float* col, int i, bool absdev, float c) {
    float v = __ldg(col + i);
    if (absdev) v = fabsf(v - c);     // abs(points1D - median), fp32 (pcl.py:79)
    return v;
}

template <int PASS>
__global__ void __launch_bounds__(kSelThreads)
select_pass_kernel(const SelJob* __restrict__ jobs) {
    using P = SelPass<PASS>;
    constexpr int NB = 1 << P::bits;
    __shared__ uint32_t s_hist[2][kSelBins];
    __shared__ uint32_t s_nan;
    __shared__ int s_last;

    const SelJob job = jobs[blockIdx.y];
    SelState* st = job.st;
    const int tid = threadIdx.x;
    const bool absdev = job.center != nullptr;
    const float c = absdev ? *job.center : 0.f;
    // masked columns: rows [0, n_it) exist, `n` of them are alive.  A marking job decides the alive bytes itself in pass 0
    // (the MAD filter that precedes this median in the reference, pcl.py:46-81) and learns n at the end of that pass.
    const bool marking = (PASS == 0) && (job.mark.col != nullptr);
    const int n_it = job.n_loop ? *job.n_loop : *job.n;
    int n = marking ? 0 : *job.n;
    const float mk_med = job.mark.col ? *job.mark.med : 0.f, mk_mad = job.mark.col ? *job.mark.mad : 1.f;
    uint8_t* __restrict__ flag = job.flag;
    uint32_t alive_local = 0, first_local = 0;

    uint32_t pre0 = 0, pre1 = 0;
    if (PASS > 0) { pre0 = st->prefix[0]; pre1 = st->prefix[1]; }
    const bool two = (PASS > 0) && (pre0 != pre1);

    for (int i = tid; i < NB; i += kSelThreads) { s_hist[0][i] = 0; s_hist[1][i] = 0; }
    if (tid == 0) s_nan = 0;
    __syncthreads();

    // grid-stride over tiles of kSelThreads*kSelItems keys; run-length aggregated smem atomics
    const int tile = kSelThreads * kSelItems;
    uint32_t nan_local = 0;
    for (int base = blockIdx.x * tile; base < n_it; base += gridDim.x * tile) {
        uint32_t run_d = 0xffffffffu, run_c = 0; int run_s = 0;
#pragma unroll
        for (int k = 0; k < kSelItems; ++k) {
            int i = base + k * kSelThreads + tid;
            bool alive = i < n_it;
            if (alive && marking) {
                const float ad = fabsf(__ldg(job.mark.col + i) - mk_med);      // abs(points1D - median)          pcl.py:79
                const float pen = (0.6745f * ad) / mk_mad;                      // 0.6745 * abs_diffs / mad_axis   pcl.py:63
                alive = pen < job.mark.thr;                                     // NaN / inf compare false         pcl.py:67
                flag[i] = alive ? 1 : 0;
                alive_local += alive ? 1u : 0u;
                if (alive) first_local = max(first_local, 0xffffffffu - (uint32_t)i);
            } else if (alive && flag) {
                alive = flag[i] != 0;
            }
            if (alive) {
                float val = sel_val(job.col, i, absdev, c);
                uint32_t key = f2key(val);
                if (PASS == 0) nan_local += (val != val) ? 1u : 0u;
                uint32_t hi = (PASS == 0) ? 0u : (key >> P::hi_shift);
                uint32_t d = (key >> P::shift) & (NB - 1);
                int s = -1;
                if (PASS == 0 || hi == pre0) s = 0;
                else if (two && hi == pre1) s = 1;
                if (s >= 0) {
                    if (d == run_d && s == run_s) { ++run_c; }
                    else {
                        if (run_c) atomicAdd(&s_hist[run_s][run_d], run_c);
                        run_d = d; run_s = s; run_c = 1;
                    }
                }
            }
        }
        if (run_c) atomicAdd(&s_hist[run_s][run_d], run_c);
    }
    if (PASS == 0 && nan_local) atomicAdd(&s_nan, nan_local);
    __syncthreads();
    for (int i = tid; i < NB; i += kSelThreads) {
        uint32_t h0 = s_hist[0][i];
        if (h0) atomicAdd(&st->hist[0][i], h0);
        if (two) { uint32_t h1 = s_hist[1][i]; if (h1) atomicAdd(&st->hist[1][i], h1); }
    }
    if (PASS == 0 && tid == 0 && s_nan) atomicAdd(&st->nan_count, s_nan);
    if (marking) {
        alive_local = (uint32_t)warp_sum((int)alive_local);
        if (lane_id() == 0 && alive_local) atomicAdd(&st->alive, alive_local);
        first_local = warp_max(first_local);
        if (lane_id() == 0 && first_local) atomicMax(&st->first_inv, first_local);
    }

    // ---- last CTA of the job resolves the digit of both ranks
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        uint32_t t = atomicAdd(&st->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (marking) {
        n = (int)__ldcg(&st->alive);
        if (tid == 0 && job.n_mark_out) *job.n_mark_out = n;
        if (tid == 0 && job.first_alive_out) { const uint32_t fi = __ldcg(&st->first_inv); *job.first_alive_out = fi ? (int)(0xffffffffu - fi) : 0; }
    }

    // copy the merged histogram(s) back to smem, zero the global copy
    for (int i = tid; i < NB; i += kSelThreads) {
        s_hist[0][i] = __ldcg(&st->hist[0][i]); st->hist[0][i] = 0;
        if (two) { s_hist[1][i] = __ldcg(&st->hist[1][i]); st->hist[1][i] = 0; }
    }
    __syncthreads();
    // every thread owns NB/kSelThreads consecutive bins; a block scan locates the bucket of each rank
    {
        constexpr int PER = (NB + kSelThreads - 1) / kSelThreads;
        __shared__ int s_scan[33];
        for (int s = 0; s < 2; ++s) {
            uint32_t rank;
            if (PASS == 0) rank = (s == 0) ? (uint32_t)((n > 0 ? n - 1 : 0) / 2) : (uint32_t)(n / 2);
            else rank = st->rank[s];
            const uint32_t* h = s_hist[(two && s == 1) ? 1 : 0];
            int mine = 0;
#pragma unroll
            for (int k = 0; k < PER; ++k) { int b = tid * PER + k; if (b < NB) mine += (int)h[b]; }
            int total;
            int excl = block_excl_scan(mine, s_scan, &total);
            if (n > 0 && rank >= (uint32_t)excl && rank < (uint32_t)(excl + mine)) {
                uint32_t cum = (uint32_t)excl; int b = tid * PER;
                for (int k = 0; k < PER; ++k) {
                    uint32_t hb = h[tid * PER + k];
                    if (rank < cum + hb) { b = tid * PER + k; break; }
                    cum += hb;
                }
                uint32_t pre = (PASS == 0) ? 0u : (s == 0 ? pre0 : pre1);
                st->prefix[s] = (PASS == 0) ? (uint32_t)b : ((pre << P::bits) | (uint32_t)b);
                st->rank[s] = rank - cum;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        st->ticket = 0;
        if (PASS == 2) {
            float a = key2f(st->prefix[0]), b = key2f(st->prefix[1]);
            float med;
            if (n == 0 || st->nan_count != 0) med = __uint_as_float(0x7fc00000u);   // np.median: NaN
            else if (n & 1) med = b;                   // ranks coincide for odd n
            else med = (a + b) * 0.5f;                 // np.mean of the two middle values, fp32
            *job.out = med;
            if (job.status && job.zero_bit && n > 0 && !(med > 0.0f)) atomicOr(job.status, job.zero_bit);
            st->nan_count = 0; st->alive = 0; st->first_inv = 0;
            st->prefix[0] = st->prefix[1] = 0; st->rank[0] = st->rank[1] = 0;
        }
    }
}

}  // namespace sd

int sd_launch_select_median(const sd::SelJob* d_jobs, int njobs, int cap, cudaStream_t st) {
    using namespace sd;
    if (njobs <= 0) return SD_OK;
    const int tile = kSelThreads * kSelItems;
    int per_job = ceil_div(cap, tile);
    int target = max(1, (148 * SD_SEL_WAVES) / njobs);
    per_job = max(1, min(per_job, target));
    dim3 grid(per_job, njobs);
    select_pass_kernel<0><<<grid, kSelThreads, 0, st>>>(d_jobs);
    select_pass_kernel<1><<<grid, kSelThreads, 0, st>>>(d_jobs);
    select_pass_kernel<2><<<grid, kSelThreads, 0, st>>>(d_jobs);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
