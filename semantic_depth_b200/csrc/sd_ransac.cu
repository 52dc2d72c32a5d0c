// RANSAC variant of the plane fit (north_star / SURVEY.md row 8-R; the reference itself only has the
// all-points least squares of pcl.py:118-120).  The host supplies seeded index triplets; each
// hypothesis is the plane through its three points in the regression form of pcl.py, and is scored
// by its inlier count under the reference's residual test (pcl.py:130-131):
//     abs((((C0*u) + (C1*v)) - w) + C2) < thr      fp64, no FMA, fixed operation order
// so a NumPy restatement reproduces every count bit for bit.  fp64-ALU bound (5 flop per test):
// each thread keeps 4 hypotheses in registers and streams points from shared memory (broadcast
// reads), counts stay in registers and are merged with integer atomics.
#include "sd_internal.cuh"

namespace sd {

constexpr int kRsThreads = 256;
constexpr int kRsHypPerThread = 4;
constexpr int kRsHypPerBlock = kRsThreads * kRsHypPerThread;   // 1024
constexpr int kRsPointTile = 256;
constexpr int kRsPointsPerBlock = 8192;

__device__ __forceinline__ void uvw_of(int axis, const float* x, const float* y, const float* z, int i,
                                       double& u, double& v, double& w) {
    const double px = (double)__ldg(x + i), py = (double)__ldg(y + i), pz = (double)__ldg(z + i);
    if (axis == 0) { u = py; v = pz; w = px; }
    else if (axis == 1) { u = px; v = pz; w = py; }
    else { u = px; v = py; w = pz; }
}

__global__ void ransac_prepare_kernel(const RansacJob* __restrict__ jobs) {
    const RansacJob J = jobs[blockIdx.y];
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= J.n_hyp) return;
    const int n = *J.n;
    const int i0 = J.triplets[3 * h], i1 = J.triplets[3 * h + 1], i2 = J.triplets[3 * h + 2];
    double c0 = 0, c1 = 0, c2 = 0, valid = 0;
    if (i0 >= 0 && i1 >= 0 && i2 >= 0 && i0 < n && i1 < n && i2 < n && i0 != i1 && i0 != i2 && i1 != i2) {
        double u0, v0, w0, u1, v1, w1, u2, v2, w2;
        uvw_of(J.axis, J.x, J.y, J.z, i0, u0, v0, w0);
        uvw_of(J.axis, J.x, J.y, J.z, i1, u1, v1, w1);
        uvw_of(J.axis, J.x, J.y, J.z, i2, u2, v2, w2);
        const double au = u1 - u0, av = v1 - v0, aw = w1 - w0;
        const double bu = u2 - u0, bv = v2 - v0, bw = w2 - w0;
        const double nu = av * bw - aw * bv;
        const double nv = aw * bu - au * bw;
        const double nw = au * bv - av * bu;
        if (nw != 0.0) {
            c0 = -nu / nw;
            c1 = -nv / nw;
            c2 = (w0 - c0 * u0) - c1 * v0;
            valid = 1.0;
        }
    }
    J.hyp_coeff[4 * h] = c0; J.hyp_coeff[4 * h + 1] = c1; J.hyp_coeff[4 * h + 2] = c2; J.hyp_coeff[4 * h + 3] = valid;
    J.hyp_counts[h] = 0;
}

__global__ void __launch_bounds__(kRsThreads)
ransac_score_kernel(const RansacJob* __restrict__ jobs) {
    __shared__ double s_u[kRsPointTile], s_v[kRsPointTile], s_w[kRsPointTile];
    const RansacJob J = jobs[blockIdx.z];
    const int n = *J.n;
    const int p_begin = blockIdx.x * kRsPointsPerBlock;
    if (p_begin >= n) return;
    const int p_end = min(n, p_begin + kRsPointsPerBlock);
    const int hb = blockIdx.y * kRsHypPerBlock;
    double c0[kRsHypPerThread], c1[kRsHypPerThread], c2[kRsHypPerThread];
    int cnt[kRsHypPerThread];
    bool ok[kRsHypPerThread];
#pragma unroll
    for (int k = 0; k < kRsHypPerThread; ++k) {
        const int h = hb + k * kRsThreads + threadIdx.x;
        ok[k] = h < J.n_hyp && J.hyp_coeff[4 * h + 3] != 0.0;
        c0[k] = ok[k] ? J.hyp_coeff[4 * h] : 0.0;
        c1[k] = ok[k] ? J.hyp_coeff[4 * h + 1] : 0.0;
        c2[k] = ok[k] ? J.hyp_coeff[4 * h + 2] : 0.0;
        cnt[k] = 0;
    }
    const double thr = J.thr;
    for (int base = p_begin; base < p_end; base += kRsPointTile) {
        const int m = min(kRsPointTile, p_end - base);
        __syncthreads();
        if ((int)threadIdx.x < m) {
            double u, v, w;
            uvw_of(J.axis, J.x, J.y, J.z, base + threadIdx.x, u, v, w);
            s_u[threadIdx.x] = u; s_v[threadIdx.x] = v; s_w[threadIdx.x] = w;
        }
        __syncthreads();
        for (int p = 0; p < m; ++p) {
            const double u = s_u[p], v = s_v[p], w = s_w[p];
#pragma unroll
            for (int k = 0; k < kRsHypPerThread; ++k) {
                const double a = ((c0[k] * u + c1[k] * v) - w) + c2[k];
                cnt[k] += (fabs(a) < thr) ? 1 : 0;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kRsHypPerThread; ++k) {
        const int h = hb + k * kRsThreads + threadIdx.x;
        if (ok[k] && cnt[k]) atomicAdd(&J.hyp_counts[h], cnt[k]);
    }
}

// arg max of the counts, lowest hypothesis index on ties; one CTA per job
__global__ void __launch_bounds__(1024)
ransac_argmax_kernel(const RansacJob* __restrict__ jobs) {
    __shared__ unsigned long long s_best[32];
    const RansacJob J = jobs[blockIdx.x];
    // key = (count << 32) | (0xffffffff - index): max key = max count, then min index
    unsigned long long best = 0ull;
    for (int h = threadIdx.x; h < J.n_hyp; h += blockDim.x) {
        unsigned long long key = ((unsigned long long)(uint32_t)J.hyp_counts[h] << 32) | (unsigned long long)(0xffffffffu - (uint32_t)h);
        best = key > best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(SD_FULL, best, o); best = t > best ? t : best; }
    if (lane_id() == 0) s_best[warp_id()] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best = s_best[w] > best ? s_best[w] : best;
        int idx = (J.n_hyp > 0) ? (int)(0xffffffffu - (uint32_t)(best & 0xffffffffull)) : -1;
        *J.best = idx;
        if (idx >= 0) { J.best_coeff[0] = J.hyp_coeff[4 * idx]; J.best_coeff[1] = J.hyp_coeff[4 * idx + 1]; J.best_coeff[2] = J.hyp_coeff[4 * idx + 2]; }
    }
}

}  // namespace sd

int sd_launch_ransac(const sd::RansacJob* d_jobs, int njobs, int cap, int n_hyp, cudaStream_t st) {
    using namespace sd;
    if (njobs <= 0 || n_hyp <= 0) return SD_OK;
    ransac_prepare_kernel<<<dim3(ceil_div(n_hyp, 256), njobs), 256, 0, st>>>(d_jobs);
    dim3 grid(ceil_div(cap, kRsPointsPerBlock), ceil_div(n_hyp, kRsHypPerBlock), njobs);
    ransac_score_kernel<<<grid, kRsThreads, 0, st>>>(d_jobs);
    ransac_argmax_kernel<<<njobs, 1024, 0, st>>>(d_jobs);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
