// Pixel stage of the fusion path: one pass over the two networks' raw outputs.
//
//   labels      softmax(logits)[:, c] > thr            semantic_depth.py:550-556,563-564
//   blend       DepthFrame.post_processing + cast       semantic_depth.py:656-664,676
//   scale       disparity * disparity_mult (fp32)       semantic_depth.py:145
//   reproject   cv2.reprojectImageTo3D(disp, Q)         semantic_depth.py:691-696
//   gather      points3D[road_mask] / [fence_mask]      semantic_depth.py:183-187  (raster order)
//   z cut       remove_from_to(road, 2, 0, 7)           semantic_depth.py:206, pcl.py:35-37
//
//   upsample    (optional) conv2d_transpose(second_skip, 16x16, stride 8)   fcn8s/fcn.py:207-213
//               the FCN-8s head evaluated inside the label kernel: 12 B/pixel of logits never exist
//
// HBM-bound: 20 B/pixel in (12 B logits + 2 x 4 B disparity), 16 B per surviving point out.
// Three streaming kernels without any dependency between CTAs:
//   pixel_label_kernel    a CTA owns 1024 consecutive pixels; its logits (12 KB, AoS) arrive in shared
//                         memory through one TMA bulk copy (cp.async.bulk + mbarrier) so HBM sees full
//                         lines, disparities are 128-bit loads (the flipped map mirrored); it decides the
//                         labels and the road z cut and writes 1 flag byte per pixel + per-tile counts
//   pixel_scan_kernel     exclusive scan of the tile counts of a frame (raster order = NumPy order)
//   pixel_scatter_kernel  blends / reprojects the kept pixels again (cheaper than storing them) and writes
//                         them at their final raster-ordered position, with their source pixel index
#include <cmath>
#include <cstring>
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

// north_star's alternative labelling: np.argmax over the three classes (first maximum wins; NaN wins like NumPy)
__device__ __forceinline__ int classify_argmax(float l0, float l1, float l2) {
    int a = 0; float m = l0;
    if (!(m != m) && (l1 > m || l1 != l1)) { a = 1; m = l1; }
    if (!(m != m) && (l2 > m || l2 != l2)) { a = 2; }
    return a == 0 ? 1 : (a == 1 ? 2 : 0);
}

struct PixArgs {
    const float* logits; const float* disp; const double* lmask; const double* rmask;
    int height, width, hw;
    float q03, q13, q23, q32, mult;
    double thr; float road_z_cut; int raw_disp;
    SdCloudBuf road, fence; int cap_stride;
    int32_t* cnt_road_gather; int32_t* cnt_road_z; int32_t* cnt_fence; int cnt_stride;
    uint8_t* labels; float* points; float* disp_pp;
    uint8_t* flags;            // [B][hw]   bit0 road, bit1 fence, bit2 road point that survives the z cut
    int32_t* tcounts;          // [B][tiles][4] per-tile counts: road (all), road kept, fence, -
    int32_t* toffs;            // [B][tiles][2] exclusive offsets of the tile's road / fence points
    int pix_tiles;
    // score-map mode: logits = conv2d_transpose(scores, upw, 16x16, stride 8, 'same') + upb
    const float* scores;       // [B][H/8][W/8][3] or nullptr
    const float* upw;          // [16][16][3 out][3 in]
    const float* upb;          // [3]
    float* logits_out;         // optional [B][hw][3]: the upsampled logits (parity tests)
    int label_mode;            // 0: softmax > thr (reference), 1: argmax (north_star wording)
    // z cut as a threshold on the scaled disparity: z(d) = fl32(q23 / (q32 * d)) is monotone in d > 0, so the host finds
    // the largest fp32 d with z(d) < -cut by bisection over the bit patterns, with the kernel's own arithmetic
    int use_dstar; uint32_t dstar_bits;
};

// blended + scaled disparity of one pixel (semantic_depth.py:660-664,676 and :145)
__device__ __forceinline__ float blend_px(const PixArgs& a, float l, float r, int u, float& dpp) {
    const float m = 0.5f * (l + r);
    const double lm = __ldg(a.lmask + u), rm = __ldg(a.rmask + u);
    if (a.raw_disp) {
        dpp = l;
    } else if (lm == 0.0 && rm == 0.0 && isfinite(m) && m != 0.0f) {
        dpp = m;                                // (0*l + 0*r) + (1-0-0)*m == m exactly
    } else {
        const double t = (rm * (double)l + lm * (double)r) + ((1.0 - lm) - rm) * (double)m;
        dpp = (float)t;
    }
    return a.raw_disp ? dpp : dpp * a.mult;     // fp32 product
}

__device__ __forceinline__ void load_disp4(const PixArgs& a, int f, int p, int v, int u0, float* l4, float* r4) {
    const float* dl = a.disp + (size_t)f * 2 * a.hw;
    const float* dr = dl + a.hw;
    const float4 L = __ldg(reinterpret_cast<const float4*>(dl + p));
    const float4 R = __ldg(reinterpret_cast<const float4*>(dr + (size_t)v * a.width + (a.width - 4 - u0)));
    l4[0] = L.x; l4[1] = L.y; l4[2] = L.z; l4[3] = L.w;
    r4[0] = R.w; r4[1] = R.z; r4[2] = R.y; r4[3] = R.x;   // np.fliplr of the second map
}

// ---- kernel 1: labels + z cut -> 1 byte per pixel and per-tile counts.  Pure streaming: every CTA is
//      independent (no scan, no look-back), 20 B read and 1 B written per pixel.
__global__ void __launch_bounds__(kPixThreads)
pixel_label_kernel(const __grid_constant__ PixArgs a) {
    __shared__ __align__(128) float s_logits[kPixTile * 3];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_cnt[kPixThreads / 32][3];
    __shared__ float s_upw[16 * 16 * 9];
    const int f = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int pix0 = tile * kPixTile;
    const int npix = min(kPixTile, a.hw - pix0);
    const bool from_scores = a.scores != nullptr;
    if (from_scores) {
        for (int i = tid; i < 16 * 16 * 9; i += kPixThreads) s_upw[i] = __ldg(a.upw + i);
    } else if (tid == 0) {
        mbar_init(&s_bar, 1);
        const uint32_t bytes = (uint32_t)npix * 12u;
        mbar_expect_tx(&s_bar, bytes);
        tma_bulk_g2s(s_logits, a.logits + ((size_t)f * a.hw + pix0) * 3, bytes, &s_bar);
    }
    __syncthreads();
    const int p = pix0 + tid * kPixPer;
    const bool active = p < a.hw;
    float l4[4] = {0.f, 0.f, 0.f, 0.f}, r4[4] = {0.f, 0.f, 0.f, 0.f};
    int v = 0, u0 = 0;
    if (active) { v = p / a.width; u0 = p - v * a.width; load_disp4(a, f, p, v, u0, l4, r4); }
    if (!from_scores) {
        mbar_wait(&s_bar, 0);
    } else if (active) {
        // logits of this thread's 4 pixels: taps ky = y + 4 - 8*iy over two low-res rows / columns, fp32, no FMA,
        // accumulated from 0.0 in the order (iy, ix, ci) ascending, bias last -- the oracle's order
        const int sh = a.height >> 3, sw = a.width >> 3;
        const float* sc = a.scores + (size_t)f * sh * sw * 3;
        const int iy0 = ((v + 4) >> 3) - 1;
        const float b0 = __ldg(a.upb), b1 = __ldg(a.upb + 1), b2 = __ldg(a.upb + 2);
#pragma unroll
        for (int j = 0; j < kPixPer; ++j) {
            const int x = u0 + j;
            const int ix0 = ((x + 4) >> 3) - 1;
            float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const int iy = iy0 + dy, ky = v + 4 - 8 * iy;
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const int ix = ix0 + dx, kx = x + 4 - 8 * ix;
                    const bool inside = iy >= 0 && iy < sh && ix >= 0 && ix < sw;
                    const float* sp = sc + ((size_t)(inside ? iy : 0) * sw + (inside ? ix : 0)) * 3;
                    const float* wp = s_upw + (ky * 16 + kx) * 9;
#pragma unroll
                    for (int ci = 0; ci < 3; ++ci) {
                        const float sv = inside ? __ldg(sp + ci) : 0.f;
                        acc0 = acc0 + sv * wp[0 * 3 + ci];
                        acc1 = acc1 + sv * wp[1 * 3 + ci];
                        acc2 = acc2 + sv * wp[2 * 3 + ci];
                    }
                }
            }
            float* lg = s_logits + (tid * kPixPer + j) * 3;
            lg[0] = acc0 + b0; lg[1] = acc1 + b1; lg[2] = acc2 + b2;
        }
        if (a.logits_out) {
            const float4* src = reinterpret_cast<const float4*>(s_logits + tid * kPixPer * 3);
            float4* dst = reinterpret_cast<float4*>(a.logits_out + ((size_t)f * a.hw + p) * 3);
            dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
        }
    }
    const double q2 = (double)a.q23, q3 = (double)a.q32;
    const float thr32 = (float)a.thr;
    int n_road = 0, n_roadz = 0, n_fence = 0;
    unsigned char fl[4] = {0, 0, 0, 0};
    if (active) {
#pragma unroll
        for (int j = 0; j < kPixPer; ++j) {
            const float* lg = s_logits + (tid * kPixPer + j) * 3;
            int lab = a.label_mode ? classify_argmax(lg[0], lg[1], lg[2]) : classify(lg[0], lg[1], lg[2], a.thr, thr32);
            if (lab & 1) {
                float dpp;
                const float d = blend_px(a, l4[j], r4[j], u0 + j, dpp);
                bool far_enough;
                if (a.use_dstar) {
                    far_enough = __float_as_uint(d) <= a.dstar_bits;     // +0 .. d*: z < -cut; negative, NaN: bits above
                } else {
                    const double wp = q3 * (double)d;
                    far_enough = div_to_f32(q2, 1.0 / wp, wp) < -a.road_z_cut;
                }
                ++n_road;
                if (far_enough) { lab |= 4; ++n_roadz; }                 // pcl.py:36
            }
            n_fence += (lab & 2) ? 1 : 0;
            fl[j] = (unsigned char)lab;
        }
        *reinterpret_cast<uchar4*>(a.flags + (size_t)f * a.hw + p) = make_uchar4(fl[0], fl[1], fl[2], fl[3]);
        if (a.labels) *reinterpret_cast<uchar4*>(a.labels + (size_t)f * a.hw + p) = make_uchar4(fl[0] & 3, fl[1] & 3, fl[2] & 3, fl[3] & 3);
    }
    n_road = warp_sum(n_road); n_roadz = warp_sum(n_roadz); n_fence = warp_sum(n_fence);
    if (lane_id() == 0) { s_cnt[warp_id()][0] = n_road; s_cnt[warp_id()][1] = n_roadz; s_cnt[warp_id()][2] = n_fence; }
    __syncthreads();
    if (tid < 3) {
        int t = 0;
        for (int w = 0; w < kPixThreads / 32; ++w) t += s_cnt[w][tid];
        a.tcounts[((size_t)f * a.pix_tiles + tile) * 4 + tid] = t;
    }
}

// ---- kernel 2: exclusive scan of the tile counts of each frame (one CTA per frame)
__global__ void __launch_bounds__(1024)
pixel_scan_kernel(const __grid_constant__ PixArgs a) {
    __shared__ int s_scan[33];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int32_t* tc = a.tcounts + (size_t)f * a.pix_tiles * 4;
    int32_t* to = a.toffs + (size_t)f * a.pix_tiles * 2;
    int run_road = 0, run_fence = 0, all_road = 0;
    for (int base = 0; base < a.pix_tiles; base += 1024) {
        const int t = base + tid;
        int cr = 0, cf = 0, ca = 0;
        if (t < a.pix_tiles) { ca = tc[t * 4]; cr = tc[t * 4 + 1]; cf = tc[t * 4 + 2]; }
        int tot_r, tot_f, tot_a;
        const int er = block_excl_scan(cr, s_scan, &tot_r);
        const int ef = block_excl_scan(cf, s_scan, &tot_f);
        block_excl_scan(ca, s_scan, &tot_a);
        if (t < a.pix_tiles) { to[t * 2] = run_road + er; to[t * 2 + 1] = run_fence + ef; }
        run_road += tot_r; run_fence += tot_f; all_road += tot_a;
    }
    if (tid == 0) {
        a.cnt_road_gather[(size_t)f * a.cnt_stride] = all_road;
        a.cnt_road_z[(size_t)f * a.cnt_stride] = run_road;
        a.cnt_fence[(size_t)f * a.cnt_stride] = run_fence;
    }
}

// ---- kernel 3: blend + reprojection of the kept pixels and their raster-ordered scatter
__global__ void __launch_bounds__(kPixThreads)
pixel_scatter_kernel(const __grid_constant__ PixArgs a) {
    __shared__ int s_scan[33];
    __shared__ float4 s_out[2][kPixTile];       // road / fence survivors of the tile in output order
    const int f = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const bool want_all = (a.points != nullptr) || (a.disp_pp != nullptr);
    const int32_t* tc = a.tcounts + ((size_t)f * a.pix_tiles + tile) * 4;
    if (!want_all && tc[1] == 0 && tc[2] == 0) return;          // nothing of this tile survives (sky)
    const int p = tile * kPixTile + tid * kPixPer;
    const bool active = p < a.hw;
    const double q0 = (double)a.q03, q1 = (double)a.q13, q2 = (double)a.q23, q3 = (double)a.q32;
    float X[4], Y[4], Z[4], D[4];
    unsigned char fl[4] = {0, 0, 0, 0};
    int n_roadz = 0, n_fence = 0;
    if (active) {
        const uchar4 fb = *reinterpret_cast<const uchar4*>(a.flags + (size_t)f * a.hw + p);
        fl[0] = fb.x; fl[1] = fb.y; fl[2] = fb.z; fl[3] = fb.w;
    }
    const bool any = want_all || ((fl[0] | fl[1] | fl[2] | fl[3]) & 6);
    if (active && any) {
        const int v = p / a.width, u0 = p - v * a.width;
        float l4[4], r4[4];
        load_disp4(a, f, p, v, u0, l4, r4);
        const double yh = q1 - (double)v;              // -1*v + cy   (row 1 of Q)
#pragma unroll
        for (int j = 0; j < kPixPer; ++j) {
            X[j] = Y[j] = Z[j] = D[j] = 0.f;
            if (want_all || (fl[j] & 6)) {
                const float d = blend_px(a, l4[j], r4[j], u0 + j, D[j]);
                const double wp = q3 * (double)d;       // W = Q[3][2]*d
                const double rwp = 1.0 / wp;
                X[j] = div_to_f32((double)(u0 + j) + q0, rwp, wp);
                Y[j] = div_to_f32(yh, rwp, wp);
                Z[j] = div_to_f32(q2, rwp, wp);
            }
            n_roadz += (fl[j] & 4) ? 1 : 0;
            n_fence += (fl[j] & 2) ? 1 : 0;
        }
        const size_t gp = (size_t)f * a.hw + p;
        if (a.disp_pp) *reinterpret_cast<float4*>(a.disp_pp + gp) = make_float4(D[0], D[1], D[2], D[3]);
        if (a.points) {
            float4* o = reinterpret_cast<float4*>(a.points + gp * 3);
            o[0] = make_float4(X[0], Y[0], Z[0], X[1]);
            o[1] = make_float4(Y[1], Z[1], X[2], Y[2]);
            o[2] = make_float4(Z[2], X[3], Y[3], Z[3]);
        }
    }
    int total;
    const int excl = block_excl_scan(n_roadz | (n_fence << 16), s_scan, &total);
    const int32_t* to = a.toffs + ((size_t)f * a.pix_tiles + tile) * 2;
    // stage the survivors in output order, then leave through coalesced stores
    int rp = excl & 0xffff, fp = excl >> 16;
#pragma unroll
    for (int j = 0; j < kPixPer; ++j) {
        if (fl[j] & 4) { s_out[0][rp] = make_float4(X[j], Y[j], Z[j], __int_as_float(p + j)); ++rp; }
        if (fl[j] & 2) { s_out[1][fp] = make_float4(X[j], Y[j], Z[j], __int_as_float(p + j)); ++fp; }
    }
    __syncthreads();
    const int n_r = total & 0xffff, n_f = total >> 16;
    const size_t cb = (size_t)f * a.cap_stride;
    {
        const size_t o = cb + to[0];
        for (int i = tid; i < n_r; i += kPixThreads) {
            const float4 q = s_out[0][i];
            a.road.x[o + i] = q.x; a.road.y[o + i] = q.y; a.road.z[o + i] = q.z; a.road.src[o + i] = __float_as_int(q.w);
        }
    }
    {
        const size_t o = cb + to[1];
        for (int i = tid; i < n_f; i += kPixThreads) {
            const float4 q = s_out[1][i];
            a.fence.x[o + i] = q.x; a.fence.y[o + i] = q.y; a.fence.z[o + i] = q.z; a.fence.src[o + i] = __float_as_int(q.w);
        }
    }
}

}  // namespace sd

// largest fp32 d >= +0 with fl32(q23 / (q32 * d)) < -cut, as a bit pattern (host restatement of the kernel's arithmetic)
static bool zcut_threshold(float q23, float q32, float cut, uint32_t* bits_out) {
    if (!(q23 < 0.f) || !(q32 > 0.f) || !(cut >= 0.f) || !std::isfinite(cut)) return false;
    auto pred = [&](uint32_t b) {
        float d; memcpy(&d, &b, 4);
        const double wp = (double)q32 * (double)d;
        const float z = (float)((double)q23 / wp);
        return z < -cut;
    };
    if (!pred(0u)) return false;                       // +0 -> z = -inf: always true for a sane camera
    uint32_t lo = 0u, hi = 0x7f800000u;                // pred(lo) true; find the last true pattern up to +inf
    if (pred(hi)) { *bits_out = hi; return true; }
    while (hi - lo > 1u) {
        const uint32_t mid = lo + (hi - lo) / 2u;
        if (pred(mid)) lo = mid; else hi = mid;
    }
    *bits_out = lo;
    return true;
}

int sd_launch_pixel(const float* d_logits, const float* d_disp, const double* d_lmask, const double* d_rmask,
                    int batch, int height, int width, SdCamera cam, double prob_thr, float road_z, int flags,
                    SdCloudBuf road, SdCloudBuf fence, int cap_stride,
                    int32_t* d_cnt_road_gather, int32_t* d_cnt_road_z, int32_t* d_cnt_fence, int cnt_stride,
                    uint8_t* d_labels, float* d_points, float* d_disp_pp,
                    uint8_t* d_flags, int32_t* d_tcounts, int32_t* d_toffs, int pix_tiles, cudaStream_t st,
                    const float* d_scores, const float* d_upw, const float* d_upb, float* d_logits_out, int label_mode) {
    using namespace sd;
    if (width % 4 != 0 || width < 4 || height < 1 || batch < 1) return SD_ERR_INVALID;
    if (d_scores && (width % 8 != 0 || height % 8 != 0 || !d_upw || !d_upb)) return SD_ERR_INVALID;
    if (!d_scores && !d_logits) return SD_ERR_INVALID;
    PixArgs a;
    a.logits = d_logits; a.disp = d_disp; a.lmask = d_lmask; a.rmask = d_rmask;
    a.height = height; a.width = width; a.hw = height * width;
    a.q03 = cam.q03; a.q13 = cam.q13; a.q23 = cam.q23; a.q32 = cam.q32; a.mult = cam.disparity_mult;
    a.thr = prob_thr; a.road_z_cut = road_z; a.raw_disp = (flags & SD_PIX_RAW_DISPARITY) ? 1 : 0;
    a.road = road; a.fence = fence; a.cap_stride = cap_stride;
    a.cnt_road_gather = d_cnt_road_gather; a.cnt_road_z = d_cnt_road_z; a.cnt_fence = d_cnt_fence;
    a.cnt_stride = cnt_stride;
    a.labels = d_labels; a.points = d_points; a.disp_pp = d_disp_pp;
    a.flags = d_flags; a.tcounts = d_tcounts; a.toffs = d_toffs; a.pix_tiles = pix_tiles;
    a.scores = d_scores; a.upw = d_upw; a.upb = d_upb; a.logits_out = d_logits_out; a.label_mode = label_mode;
    a.dstar_bits = 0u;
    a.use_dstar = zcut_threshold(cam.q23, cam.q32, road_z, &a.dstar_bits) ? 1 : 0;
    dim3 grid(pix_tiles, batch);
    pixel_label_kernel<<<grid, kPixThreads, 0, st>>>(a);
    pixel_scan_kernel<<<batch, 1024, 0, st>>>(a);
    pixel_scatter_kernel<<<grid, kPixThreads, 0, st>>>(a);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
