// Pixel stage of the fusion path: one pass over the two networks' raw outputs.
//
//   labels      softmax(logits)[:, c] > thr            semantic_depth.py:550-556,563-564
//   blend       DepthFrame.post_processing + cast       semantic_depth.py:656-664,676
//   scale       disparity * disparity_mult (fp32)       semantic_depth.py:145
//   reproject   cv2.reprojectImageTo3D(disp, Q)         semantic_depth.py:691-696
//   gather      points3D[road_mask] / [fence_mask]      semantic_depth.py:183-187  (raster order)
//   z cut       remove_from_to(road, 2, 0, 7)           semantic_depth.py:206, pcl.py:35-37
//
// HBM-bound: 20 B/pixel in (12 B logits + 2 x 4 B disparity), 16 B per surviving point out.
// A CTA owns a tile of 1024 consecutive pixels.  The logits tile (12 KB, AoS) is staged in shared
// memory by one TMA bulk copy (cp.async.bulk + mbarrier) so HBM sees full 128 B lines; disparities
// are read as 128-bit vectors (the flipped map mirrored).  Per-class stable compaction = warp
// shuffles + block scan + decoupled look-back over dynamically ticketed tiles: output order is the
// raster order NumPy boolean indexing produces, and the source pixel index of every point is kept.
#include "sd_internal.cuh"

namespace sd {

constexpr int kPixThreads = 256;
constexpr int kPixPer = 4;
constexpr int kPixTile = kPixThreads * kPixPer;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    } while (!ok);
}

// softmax(l)[c] > thr for c = 0 (road, bit 0) and c = 1 (fence, bit 1).
// Contract: the decision of an fp64 softmax.  Tier 1 evaluates an fp32 estimate
// whose error is < 1e-5 and accepts it when it is further than 1e-4 from the threshold; only the
// remaining sliver (and non-finite logits) runs the fp64 expression of the oracle:
//   e = exp(l - max l);  p = e / ((e0 + e1) + e2);  p_c > thr.
__device__ __forceinline__ int classify(float l0, float l1, float l2, double thr, float thr32) {
    float m = fmaxf(l0, fmaxf(l1, l2));
    float e0 = __expf(l0 - m), e1 = __expf(l1 - m), e2 = __expf(l2 - m);
    float inv = 1.0f / ((e0 + e1) + e2);
    float p0 = e0 * inv, p1 = e1 * inv;
    bool sure0 = fabsf(p0 - thr32) > 1e-4f, sure1 = fabsf(p1 - thr32) > 1e-4f;   // false for NaN
    int r = (p0 > thr32 ? 1 : 0) | (p1 > thr32 ? 2 : 0);
    if (sure0 && sure1) return r;
    if (isnan(l0) || isnan(l1) || isnan(l2)) return 0;   // np.max propagates NaN -> every p is NaN
    double dm = (double)m;
    double d0 = exp((double)l0 - dm), d1 = exp((double)l1 - dm), d2 = exp((double)l2 - dm);
    double s = (d0 + d1) + d2;
    return ((d0 / s) > thr ? 1 : 0) | ((d1 / s) > thr ? 2 : 0);
}

// fp32 rounding of num/den evaluated in fp64 (what OpenCV stores).  Fast path: num * (1/den) with a
// guard that detects the (2^-27-rare) case where the two-rounding product could sit on the other
// side of an fp32 rounding boundary than the correctly rounded quotient; those take the division.
__device__ __forceinline__ float div_to_f32(double num, double rden, double den) {
    double q = num * rden;
    uint32_t lo = (uint32_t)__double2loint(q) & 0x1fffffffu;      // bits below the fp32 mantissa
    uint32_t dist = lo > 0x10000000u ? lo - 0x10000000u : 0x10000000u - lo;
    double aq = fabs(q);
    bool safe = (dist > 8u) && (aq > 1e-37) && (aq < 1e38);        // NaN/inf/0/subnormal -> exact path
    if (!safe) q = num / den;
    return (float)q;
}

struct PixArgs {
    const float* logits; const float* disp; const double* lmask; const double* rmask;
    int height, width, hw;
    float q03, q13, q23, q32, mult;
    double thr; float road_z_cut; int raw_disp;
    SdCloudBuf road, fence; int cap_stride;
    int32_t* cnt_road_gather; int32_t* cnt_road_z; int32_t* cnt_fence; int cnt_stride;
    uint8_t* labels; float* points; float* disp_pp;
    unsigned long long* status; ScanCtl* ctl; int pix_tiles;
};

__global__ void __launch_bounds__(kPixThreads)
pixel_fuse_kernel(const __grid_constant__ PixArgs a) {
    __shared__ __align__(128) float s_logits[kPixTile * 3];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_scan[33];
    __shared__ int s_tile;
    __shared__ unsigned long long s_excl;

    const int f = blockIdx.y;
    const int tid = threadIdx.x;
    ScanCtl* ctl = a.ctl + f;
    unsigned long long* status = a.status + (size_t)f * a.pix_tiles;

    if (tid == 0) {
        s_tile = (int)atomicAdd(&ctl->ticket, 1u);
        mbar_init(&s_bar, 1);
    }
    __syncthreads();
    const int tile = s_tile;
    const int pix0 = tile * kPixTile;
    const int npix = min(kPixTile, a.hw - pix0);

    if (tid == 0) {
        uint32_t bytes = (uint32_t)npix * 12u;
        mbar_expect_tx(&s_bar, bytes);
        tma_bulk_g2s(s_logits, a.logits + ((size_t)f * a.hw + pix0) * 3, bytes, &s_bar);
    }

    const int p = pix0 + tid * kPixPer;           // first pixel of this thread (flat, in frame)
    const bool active = p < a.hw;
    float l4[4] = {0.f, 0.f, 0.f, 0.f}, r4[4] = {0.f, 0.f, 0.f, 0.f};
    int v = 0, u0 = 0;
    if (active) {
        v = p / a.width;
        u0 = p - v * a.width;
        const float* dl = a.disp + (size_t)f * 2 * a.hw;
        const float* dr = dl + a.hw;
        float4 L = __ldg(reinterpret_cast<const float4*>(dl + p));
        float4 R = __ldg(reinterpret_cast<const float4*>(dr + (size_t)v * a.width + (a.width - 4 - u0)));
        l4[0] = L.x; l4[1] = L.y; l4[2] = L.z; l4[3] = L.w;
        r4[0] = R.w; r4[1] = R.z; r4[2] = R.y; r4[3] = R.x;   // np.fliplr of the second map
    }
    mbar_wait(&s_bar, 0);

    const double q0 = (double)a.q03, q1 = (double)a.q13, q2 = (double)a.q23, q3 = (double)a.q32;
    const float thr32 = (float)a.thr;
    float X[4], Y[4], Z[4], D[4];
    int lab[4];
    int n_road = 0, n_roadz = 0, n_fence = 0;
    bool keep_road[4], keep_fence[4];
    const bool want_all = (a.points != nullptr);
    const double yh = q1 - (double)v;              // -1*v + cy   (row 1 of Q)
#pragma unroll
    for (int j = 0; j < kPixPer; ++j) {
        keep_road[j] = keep_fence[j] = false;
        lab[j] = 0;
        X[j] = Y[j] = Z[j] = D[j] = 0.f;
        if (!active) continue;
        const int u = u0 + j;
        const float* lg = s_logits + (tid * kPixPer + j) * 3;
        lab[j] = classify(lg[0], lg[1], lg[2], a.thr, thr32);
        // ---- blend (semantic_depth.py:660-664): m is fp32, the ramps are fp64
        const float l = l4[j], r = r4[j];
        const float m = 0.5f * (l + r);
        const double lm = __ldg(a.lmask + u), rm = __ldg(a.rmask + u);
        float dpp;
        if (a.raw_disp) {
            dpp = l;
        } else if (lm == 0.0 && rm == 0.0 && isfinite(m) && m != 0.0f) {
            dpp = m;                                // (0*l + 0*r) + (1-0-0)*m == m exactly
        } else {
            double t = (rm * (double)l + lm * (double)r) + ((1.0 - lm) - rm) * (double)m;
            dpp = (float)t;
        }
        D[j] = dpp;
        if (lab[j] != 0 || want_all) {
            const float d = a.raw_disp ? dpp : dpp * a.mult;   // semantic_depth.py:145 (fp32 product)
            const double wp = q3 * (double)d;       // W = Q[3][2]*d
            const double rwp = 1.0 / wp;
            const double xh = (double)u + q0;       // 1*u - cx
            X[j] = div_to_f32(xh, rwp, wp);
            Y[j] = div_to_f32(yh, rwp, wp);
            Z[j] = div_to_f32(q2, rwp, wp);
        }
        if (lab[j] & 1) {
            ++n_road;
            keep_road[j] = Z[j] < -a.road_z_cut;    // pcl.py:36
            n_roadz += keep_road[j] ? 1 : 0;
        }
        if (lab[j] & 2) { keep_fence[j] = true; ++n_fence; }
    }

    // ---- optional dense outputs (parity tests / facade): labels, points3D, blended disparity
    if (active) {
        const size_t gp = (size_t)f * a.hw + p;
        if (a.labels) {
            uchar4 lb = make_uchar4((unsigned char)lab[0], (unsigned char)lab[1], (unsigned char)lab[2], (unsigned char)lab[3]);
            *reinterpret_cast<uchar4*>(a.labels + gp) = lb;
        }
        if (a.disp_pp) *reinterpret_cast<float4*>(a.disp_pp + gp) = make_float4(D[0], D[1], D[2], D[3]);
        if (a.points) {
            float4* o = reinterpret_cast<float4*>(a.points + gp * 3);
            o[0] = make_float4(X[0], Y[0], Z[0], X[1]);
            o[1] = make_float4(Y[1], Z[1], X[2], Y[2]);
            o[2] = make_float4(Z[2], X[3], Y[3], Z[3]);
        }
    }

    // ---- stable compaction of both classes: block scan of packed counts + look-back across tiles
    int total;
    const int packed = n_roadz | (n_fence << 16);
    const int excl = block_excl_scan(packed, s_scan, &total);
    // block total of *all* road pixels (count only, no ordering needed)
    int ra = warp_sum(n_road);
    if (lane_id() == 0 && ra) atomicAdd(&ctl->aux0, (unsigned)ra);

    if (warp_id() == 0) {
        unsigned long long agg = (unsigned long long)(total & 0xffff) | ((unsigned long long)(total >> 16) << 31);
        unsigned long long e = lookback_exclusive(status, tile, agg);
        if (lane_id() == 0) s_excl = e;
    }
    __syncthreads();
    const unsigned long long ex = s_excl;
    int road_pos = (int)(ex & 0x7fffffffull) + (excl & 0xffff);
    int fence_pos = (int)((ex >> 31) & 0x7fffffffull) + (excl >> 16);
    const size_t cb = (size_t)f * a.cap_stride;
#pragma unroll
    for (int j = 0; j < kPixPer; ++j) {
        if (keep_road[j]) {
            a.road.x[cb + road_pos] = X[j]; a.road.y[cb + road_pos] = Y[j]; a.road.z[cb + road_pos] = Z[j];
            a.road.src[cb + road_pos] = p + j;
            ++road_pos;
        }
        if (keep_fence[j]) {
            a.fence.x[cb + fence_pos] = X[j]; a.fence.y[cb + fence_pos] = Y[j]; a.fence.z[cb + fence_pos] = Z[j];
            a.fence.src[cb + fence_pos] = p + j;
            ++fence_pos;
        }
    }
    if (tile == a.pix_tiles - 1 && tid == 0) {
        a.cnt_road_z[(size_t)f * a.cnt_stride] = (int)(ex & 0x7fffffffull) + (total & 0xffff);
        a.cnt_fence[(size_t)f * a.cnt_stride] = (int)((ex >> 31) & 0x7fffffffull) + (total >> 16);
    }
    // ---- last block of the frame: publish the road_gather count, clean the look-back words
    if (scan_finish(ctl, status, a.pix_tiles, a.pix_tiles)) {
        if (tid == 0) {
            a.cnt_road_gather[(size_t)f * a.cnt_stride] = (int)atomicExch(&ctl->aux0, 0u);
        }
    }
}

}  // namespace sd

int sd_launch_pixel(const float* d_logits, const float* d_disp, const double* d_lmask, const double* d_rmask,
                    int batch, int height, int width, SdCamera cam, double prob_thr, float road_z, int flags,
                    SdCloudBuf road, SdCloudBuf fence, int cap_stride,
                    int32_t* d_cnt_road_gather, int32_t* d_cnt_road_z, int32_t* d_cnt_fence, int cnt_stride,
                    uint8_t* d_labels, float* d_points, float* d_disp_pp,
                    unsigned long long* status, sd::ScanCtl* ctl, int pix_tiles, cudaStream_t st) {
    using namespace sd;
    if (width % 4 != 0 || width < 4 || height < 1 || batch < 1) return SD_ERR_INVALID;
    PixArgs a;
    a.logits = d_logits; a.disp = d_disp; a.lmask = d_lmask; a.rmask = d_rmask;
    a.height = height; a.width = width; a.hw = height * width;
    a.q03 = cam.q03; a.q13 = cam.q13; a.q23 = cam.q23; a.q32 = cam.q32; a.mult = cam.disparity_mult;
    a.thr = prob_thr; a.road_z_cut = road_z; a.raw_disp = (flags & SD_PIX_RAW_DISPARITY) ? 1 : 0;
    a.road = road; a.fence = fence; a.cap_stride = cap_stride;
    a.cnt_road_gather = d_cnt_road_gather; a.cnt_road_z = d_cnt_road_z; a.cnt_fence = d_cnt_fence;
    a.cnt_stride = cnt_stride;
    a.labels = d_labels; a.points = d_points; a.disp_pp = d_disp_pp;
    a.status = status; a.ctl = ctl; a.pix_tiles = pix_tiles;
    dim3 grid(pix_tiles, batch);
    pixel_fuse_kernel<<<grid, kPixThreads, 0, st>>>(a);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
