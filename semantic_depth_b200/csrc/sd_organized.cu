// Organized-cloud neighbour search for the fused path: the two Open3D filters of the road chain
// (semantic_depth.py:227-245) when the cloud still knows which pixel every point came from.
//
// A road point is p = s * r(u,v) with r(u,v) = (u - cx, cy - v, -f) the ray of its pixel (that is what
// cv2.reprojectImageTo3D produced, semantic_depth.py:691-696).  Any point p' on the ray of another pixel
// at pixel distance delta satisfies
//        |p - p'| >= |p| * sin(angle) >= |p| * f * delta / (|r| * (|r| + delta))
// (|r x r'| >= f * delta, |r'| <= |r| + delta, and the right-hand side grows with delta).  So every
// neighbour within distance rho of p lives in a pixel window whose radius depends only on rho / |p|:
// 5-9 pixels for the k-th neighbour of a road point at ANY depth, while a world-space grid sees a
// 1500x density range.  The cloud is kept as a per-pixel float4 image (x, y, z, bits(index); +inf where
// there is no point), so the search is a stencil: no sort, no cell lists, uniform work per query.
//
//   org_knn_kernel       one thread per point; disc of 293 pixels (du^2 + dv^2 <= 90); branch-free top-k
//                        network on packed 32-bit keys (23 bits of the fp32 squared distance | 9-bit window
//                        ordinal); the k survivors are re-evaluated in fp64 and the result is certified
//                        (a) against the truncated (k+1)-th key and (b) against the ray bound of the first
//                        unscanned pixel; uncertified queries go to a queue
//   org_knn_hard_kernel  one warp per queued query; square windows 16, 32, 64, ... until the ray bound
//                        holds; fp32 keys, band rescan, exact fp64 selection; the last warp finalises the
//                        cloud statistics (exact 128-bit fixed-point sums -> order independent)
//   org_apply_sor_kernel removes the statistical outliers from the per-pixel image
//   org_ror_kernel       one thread per survivor; radius count with early exit in the same disc
//   org_ror_hard_kernel  one warp per undecided query; windows up to the radius' own pixel bound
//
// Arithmetic of the accepted results is the oracle's: fp64 (dx*dx + dy*dy) + dz*dz, IEEE sqrt, ascending
// sum from 0.0; fp32 is only ever used to discard candidates with a proven margin.
#include "sd_internal.cuh"

namespace sd {

constexpr int kOrgThreads = 128;
constexpr int kOrgW = 11;                      // main window: |du|, |dv| <= 11 and du^2 + dv^2 <= kOrgR2
constexpr int kOrgR2 = 132;
constexpr int kOrgSide = 2 * kOrgW + 1;        // 23 -> ordinals < 529 < 1024 (10 bits)
constexpr int kOrgOrdBits = 10;
constexpr int kOrgKeyShift = 32 - 1 - (32 - kOrgOrdBits);   // fp32 bits dropped from the key: 9 (22-bit key)
constexpr double kOrgOut = 11.532562594670797; // sqrt(133): pixel distance of the nearest unscanned pixel
__constant__ int c_org_hw[kOrgW + 1] = {11, 11, 11, 11, 10, 10, 9, 9, 8, 7, 5, 3};   // floor(sqrt(132 - dv^2))
constexpr int kOrgInner = 4;                   // phase 0: the (2*4+1)^2 = 81 central pixels go through the network
constexpr int kOrgList = 32;                   // phase 1: per-query list of candidates under the phase-0 threshold
constexpr float kKeyErrF = 1.5e-6f;            // relative error bound of the fp32 squared distance
constexpr int kHardList = 64;

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

__device__ __forceinline__ float key_of(const float4& c, float qx, float qy, float qz) {
    const float fx = c.x - qx, fy = c.y - qy, fz = c.z - qz;
    return fmaf(fz, fz, fmaf(fy, fy, fx * fx));          // only ever compared with a 1.5e-6 margin
}
__device__ __forceinline__ double dist2_exact(const float4& c, float qx, float qy, float qz) {
    const double dx = (double)c.x - (double)qx, dy = (double)c.y - (double)qy, dz = (double)c.z - (double)qz;
    return (dx * dx + dy * dy) + dz * dz;
}

// squared lower bound of the distance from p to any point of a pixel at pixel distance >= dout
__device__ __forceinline__ double rho2_safe(double P, double rr, double f, double dout) {
    double rho = P * f * dout / (rr * (rr + dout));
    rho = rho * (1.0 - 1e-3) - 4e-7 * P;            // fp32 rounding of the stored coordinates
    return rho > 0.0 ? rho * rho : 0.0;
}

struct Query {
    float qx, qy, qz; int u, v; double P, rr;
};
__device__ __forceinline__ Query load_query(const OrgJob& J, int i) {
    Query q;
    q.qx = __ldg(J.x + i); q.qy = __ldg(J.y + i); q.qz = __ldg(J.z + i);
    const int pix = __ldg(J.src + i);
    q.v = pix / J.width; q.u = pix - q.v * J.width;
    const double a = (double)q.u + (double)J.q03, b = (double)J.q13 - (double)q.v, f = (double)J.q23;
    q.rr = sqrt(a * a + b * b + f * f);
    q.P = sqrt((double)q.qx * q.qx + (double)q.qy * q.qy + (double)q.qz * q.qz);
    return q;
}

template <int KS>
struct UNet {                               // ascending uint32 keys, branch-free insertion
    uint32_t d[KS];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int p = 0; p < KS; ++p) d[p] = 0xffffffffu;
    }
    __device__ __forceinline__ void feed(uint32_t v) {
#pragma unroll
        for (int p = 0; p < KS; ++p) { const uint32_t lo = min(d[p], v); v = max(d[p], v); d[p] = lo; }
    }
    __device__ __forceinline__ uint32_t get(int idx) const {
        uint32_t r = d[0];
#pragma unroll
        for (int p = 1; p < KS; ++p) r = (p == idx) ? d[p] : r;
        return r;
    }
};

template <int K>
struct DNet {                               // ascending fp64, branch-free insertion
    double d[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int p = 0; p < K; ++p) d[p] = __longlong_as_double(0x7ff0000000000000ll);
    }
    __device__ __forceinline__ void feed(double v) {
#pragma unroll
        for (int p = 0; p < K; ++p) { const double lo = fmin(d[p], v); v = fmax(d[p], v); d[p] = lo; }
    }
    __device__ __forceinline__ double get(int idx) const {
        double r = d[0];
#pragma unroll
        for (int p = 1; p < K; ++p) r = (p == idx) ? d[p] : r;
        return r;
    }
    __device__ __forceinline__ double sum_sqrt(int keff) const {
        double s = 0.0;
#pragma unroll
        for (int p = 0; p < K; ++p) if (p < keff) s = s + sqrt(d[p]);
        return s;
    }
};

struct StatAcc {
    U128 sum{0ull, 0ull}, sq{0ull, 0ull}; unsigned long long pos = 0ull;
    __device__ __forceinline__ void add(double avg) {
        if (avg > 0.0) { sum = add128(sum, to_fixed70(avg)); sq = add128(sq, to_fixed70(avg * avg)); ++pos; }
    }
    __device__ __forceinline__ void flush_warp(OrgState* st) {       // all lanes
        sum = warp_sum128(sum); sq = warp_sum128(sq);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pos += __shfl_xor_sync(SD_FULL, pos, o);
        if (lane_id() == 0 && pos) {
            atomic_add128(st->acc[0], sum); atomic_add128(st->acc[1], sq); atomicAdd(&st->acc[2][0], pos);
        }
    }
};

// ---------------------------------------------------------------------------------------------
__global__ void org_fill_kernel(float4* dense, size_t count) {
    const float inf = __int_as_float(0x7f800000);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        dense[i] = make_float4(inf, inf, inf, inf);
}

// ---------------------------------------------------------------------------------------------
// k-NN, main kernel
// ---------------------------------------------------------------------------------------------
template <bool INTERIOR>
__device__ __forceinline__ float4 load_px(const float4* dense, int W, int H, int row, int col) {
    if (INTERIOR) return ldg4(dense + (size_t)row * W + col);
    const float inf = __int_as_float(0x7f800000);
    if (row < 0 || row >= H || col < 0 || col >= W) return make_float4(inf, inf, inf, inf);
    return ldg4(dense + (size_t)row * W + col);
}

__device__ __forceinline__ uint32_t pack_key(float key, int ord) {
    return ((__float_as_uint(key) >> kOrgKeyShift) << kOrgOrdBits) | (uint32_t)ord;     // key >= 0 or +inf
}

// exclusive fp32 threshold two truncation cells above the network's current k-th key: everything below it
// stays in the race, everything at or above it is provably farther than the k-th (margin 2^-14 >> fp32 error)
template <int KS>
__device__ __forceinline__ float threshold_of(const UNet<KS>& net, int keff) {
    const uint32_t kth = (keff > 0) ? net.get(keff - 1) : 0u;
    const uint32_t cell = kth >> kOrgOrdBits;
    if (kth == 0xffffffffu || cell >= (0x7f800000u >> kOrgKeyShift) - 2u) return __int_as_float(0x7f800000);
    return __uint_as_float((cell + 2u) << kOrgKeyShift);
}

// One query per lane, the whole warp in lock step.  Phase 0 sends the central 9x9 pixels through the
// branch-free network; the rest of the disc is swept with a threshold compare + append (11 ops per
// candidate); the appended candidates are flushed through the network whenever a list could overflow
// during the next row, which also tightens the threshold.
template <int KS, bool INTERIOR>
__device__ __forceinline__ void org_knn_sweep(const float4* __restrict__ dense, int W, int H, const Query& q, int keff,
                                              uint32_t (*s_list)[kOrgThreads], UNet<KS>& net, unsigned long long& dbg_list) {
    const int tid = threadIdx.x;
    // ---- phase 0
#pragma unroll 1
    for (int dv = -kOrgInner; dv <= kOrgInner; ++dv) {
        uint32_t pk[2 * kOrgInner + 1];
#pragma unroll
        for (int s = 0; s < 2 * kOrgInner + 1; ++s) {
            const int du = s - kOrgInner;
            const float4 c = load_px<INTERIOR>(dense, W, H, q.v + dv, q.u + du);
            pk[s] = pack_key(key_of(c, q.qx, q.qy, q.qz), (dv + kOrgW) * kOrgSide + (du + kOrgW));
        }
#pragma unroll
        for (int s = 0; s < 2 * kOrgInner + 1; ++s) net.feed(pk[s]);
    }
    float thr = threshold_of<KS>(net, keff);
    int cnt = 0;
    auto flush = [&]() {
        const int cmax = __reduce_max_sync(SD_FULL, cnt);
        for (int e0 = 0; e0 < cmax; e0 += 4) {
            uint32_t pk[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                pk[h] = 0xffffffffu;
                if (e0 + h < cnt) {
                    const int ord = (int)s_list[e0 + h][tid];
                    const int dv = ord / kOrgSide - kOrgW, du = ord - (ord / kOrgSide) * kOrgSide - kOrgW;
                    const float4 c = load_px<INTERIOR>(dense, W, H, q.v + dv, q.u + du);
                    pk[h] = pack_key(key_of(c, q.qx, q.qy, q.qz), ord);
                }
            }
#pragma unroll
            for (int h = 0; h < 4; ++h) net.feed(pk[h]);
        }
        dbg_list += (unsigned long long)cnt;
        cnt = 0;
        thr = threshold_of<KS>(net, keff);
    };
    // ---- phase 1
#pragma unroll 1
    for (int t = 0; t < kOrgSide; ++t) {
        const int dv = (t == 0) ? 0 : ((t & 1) ? (t + 1) / 2 : -(t / 2));
        const int adv = dv < 0 ? -dv : dv;
        const int hw = c_org_hw[adv];
        const int row = q.v + dv;
        const int ordbase = (dv + kOrgW) * kOrgSide + kOrgW;
        const bool inner_row = adv <= kOrgInner;
#pragma unroll
        for (int du = -kOrgW; du <= kOrgW; ++du) {
            const int adu = du < 0 ? -du : du;
            if (adu <= hw && !(inner_row && adu <= kOrgInner)) {            // warp-uniform
                const float4 c = load_px<INTERIOR>(dense, W, H, row, q.u + du);
                const bool in = key_of(c, q.qx, q.qy, q.qz) < thr;          // +inf (empty pixel) never passes
                if (in) s_list[cnt][tid] = (uint32_t)(ordbase + du);
                cnt += in ? 1 : 0;
            }
        }
        if (__any_sync(SD_FULL, cnt > kOrgList - kOrgSide)) flush();        // room for one more full row
    }
    flush();
}

template <int KS>
__global__ void __launch_bounds__(kOrgThreads)
org_knn_kernel(const OrgJob* __restrict__ jobs) {
    constexpr int K = KS - 1;
    __shared__ uint32_t s_list[kOrgList][kOrgThreads];
    const OrgJob J = jobs[blockIdx.y];
    const int n = *J.n;
    const int keff = min(J.k, n);
    const int W = J.width, H = J.height;
    const double f = fabs((double)J.q23);
    const float4* __restrict__ dense = J.dense;
    const int lane = lane_id();
    const int warps_total = gridDim.x * (kOrgThreads / 32);
    StatAcc acc;
    unsigned long long dbg_list = 0ull;
    for (int base = (blockIdx.x * (kOrgThreads / 32) + warp_id()) * 32; base < n; base += warps_total * 32) {
        const bool active = base + lane < n;
        const int i = active ? base + lane : n - 1;
        const Query q = load_query(J, i);
        UNet<KS> net; net.init();
        const bool interior = q.u >= kOrgW && q.u + kOrgW < W && q.v >= kOrgW && q.v + kOrgW < H;
        if (__all_sync(SD_FULL, interior)) org_knn_sweep<KS, true>(dense, W, H, q, keff, s_list, net, dbg_list);
        else org_knn_sweep<KS, false>(dense, W, H, q, keff, s_list, net, dbg_list);
        const bool overflow = false;
        // ---- the k survivors in fp64, ascending
        DNet<(K > 0 ? K : 1)> ex; ex.init();
#pragma unroll
        for (int p = 0; p < K; ++p) {
            if (p < keff) {
                const uint32_t pk = net.d[p];
                double d2 = __longlong_as_double(0x7ff0000000000000ll);
                if (pk != 0xffffffffu) {
                    const int ord = (int)(pk & ((1u << kOrgOrdBits) - 1u));
                    const int dv = ord / kOrgSide - kOrgW, du = ord - (ord / kOrgSide) * kOrgSide - kOrgW;
                    const float4 c = ldg4(dense + (size_t)(q.v + dv) * W + (q.u + du));
                    d2 = dist2_exact(c, q.qx, q.qy, q.qz);
                }
                ex.feed(d2);
            }
        }
        const double dk = (keff > 0) ? ex.get(keff - 1) : 0.0;
        // (a) nothing else in the window can beat the k-th: every other candidate's fp32 key is >= the
        //     truncated (k+1)-th key, and its exact distance is within kKeyErrF of its key
        const uint32_t pk1 = net.get(keff < KS ? keff : KS - 1);
        const double t_next = (pk1 == 0xffffffffu) ? __longlong_as_double(0x7ff0000000000000ll)
                                                    : (double)__uint_as_float((pk1 >> kOrgOrdBits) << kOrgKeyShift);
        const bool ok_window = dk * (1.0 + 2.0 * (double)kKeyErrF) <= t_next;
        // (b) nothing outside the window can: ray bound of the nearest unscanned pixel (or the image is covered)
        const bool covered = (q.u - kOrgW <= 0) && (q.u + kOrgW >= W - 1) && (q.v - kOrgW <= 0) && (q.v + kOrgW >= H - 1) &&
                             (W <= kOrgSide) && (H <= kOrgSide) && false;   // the disc never covers a real image
        const bool ok_outside = covered || (dk <= rho2_safe(q.P, q.rr, f, kOrgOut));
        if (active) {
            if (lane == 0) atomicAdd(&J.st->dbg[0], 32ull);
            if (ok_window && ok_outside && !overflow && isfinite(dk)) {
                const double avg = (keff > 0) ? ex.sum_sqrt(keff) / (double)keff : -1.0;
                J.avg[i] = avg;
                acc.add(avg);
            } else {
                atomicAdd(&J.st->dbg[1], 1ull);
                if (!ok_window) atomicAdd(&J.st->dbg[2], 1ull);
                if (!ok_outside) atomicAdd(&J.st->dbg[3], 1ull);
                if (overflow) atomicAdd(&J.st->dbg[4], 1ull);
                // dk is the exact k-th distance among SOME k points, hence an upper bound of the true one
                const int pos = atomicAdd(&J.st->qn_knn, 1);
                J.queue_knn[pos] = i;
                J.queue_bound[pos] = isfinite(dk) ? __double2float_ru(dk) : __int_as_float(0x7f800000);
            }
        }
    }
    acc.flush_warp(J.st);
    if (dbg_list) atomicAdd(&J.st->dbg[7], dbg_list);
}

// ---------------------------------------------------------------------------------------------
// k-NN, hard queries: one warp per query, growing square windows
// ---------------------------------------------------------------------------------------------
template <typename F>
__device__ __forceinline__ void warp_scan_annulus(const float4* dense, int W, int H, int u, int v, int w, int wprev, F&& f) {
    const int lane = lane_id();
    const int r0 = max(v - w, 0), r1 = min(v + w, H - 1);
    const int cl = max(u - w, 0), ch = min(u + w, W - 1);
    for (int row = r0; row <= r1; ++row) {
        const int adv = row > v ? row - v : v - row;
        const float4* rp = dense + (size_t)row * W;
        if (wprev < 0 || adv > wprev) {
            for (int c0 = cl; c0 <= ch; c0 += 32) { const int col = c0 + lane; f(col <= ch, rp, row, col); }
        } else {
            const int le = min(u - wprev - 1, ch), rs = max(u + wprev + 1, cl);
            for (int c0 = cl; c0 <= le; c0 += 32) { const int col = c0 + lane; f(col <= le, rp, row, col); }
            for (int c0 = rs; c0 <= ch; c0 += 32) { const int col = c0 + lane; f(col <= ch, rp, row, col); }
        }
    }
}

template <int KS>
__global__ void __launch_bounds__(kOrgThreads)
org_knn_hard_kernel(const OrgJob* __restrict__ jobs) {
    constexpr int K = (KS - 1 > 0) ? KS - 1 : 1;
    __shared__ int s_pix[kOrgThreads / 32][kHardList];
    __shared__ double s_sorted[kOrgThreads / 32][kHardList];
    const OrgJob J = jobs[blockIdx.y];
    const int n = *J.n;
    const int keff = min(J.k, n);
    const int W = J.width, H = J.height;
    const double f = fabs((double)J.q23);
    const float4* __restrict__ dense = J.dense;
    const int lane = lane_id(), wid = warp_id();
    const int warps_total = gridDim.x * (kOrgThreads / 32);
    const int qn = J.st->qn_knn;
    const float finf = __int_as_float(0x7f800000);
    const double dinf = __longlong_as_double(0x7ff0000000000000ll);
    StatAcc acc;
    for (int qi = blockIdx.x * (kOrgThreads / 32) + wid; qi < qn; qi += warps_total) {
        const int i = J.queue_knn[qi];
        const Query q = load_query(J, i);
        const float ub = J.queue_bound[qi];
        if (ub < finf && keff > 0) {
            // ---- fast path: the true k nearest all lie within sqrt(ub); the ray bound turns that into a
            //      pixel radius, every candidate with a key below ub (+ fp32 margin) is collected, then fp64
            const double sneed = (sqrt((double)ub) + 4e-7 * q.P) / (1.0 - 1e-3);
            const double den = q.P * f - sneed * q.rr;
            int w = -1;
            if (den > 0.0) {
                const double dneed = sneed * q.rr * q.rr / den;       // rho_safe(dneed) >= sqrt(ub)
                if (dneed < 2000.0) w = max((int)ceil(dneed) - 1, 0);
            }
            if (w >= 0) {
                const float band = ub * (1.0f + 4.0f * kKeyErrF);
                int cnt = 0;
                {
                    const int r0 = max(q.v - w, 0), r1 = min(q.v + w, H - 1);
                    const int cl = max(q.u - w, 0), ch = min(q.u + w, W - 1);
                    for (int c0 = cl; c0 <= ch; c0 += 32) {
                        const int col = c0 + lane;
                        const bool cv = col <= ch;
                        for (int rb = r0; rb <= r1; rb += 4) {
                            float4 c[4];
#pragma unroll
                            for (int h = 0; h < 4; ++h)
                                c[h] = (cv && rb + h <= r1) ? ldg4(dense + (size_t)(rb + h) * W + col) : make_float4(finf, finf, finf, finf);
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                const bool in = key_of(c[h], q.qx, q.qy, q.qz) <= band;     // +inf for empty / masked
                                const unsigned mk = __ballot_sync(SD_FULL, in);
                                const int pos = cnt + __popc(mk & ((1u << lane) - 1u));
                                if (in && pos < kHardList) s_pix[wid][pos] = (rb + h) * W + col;
                                cnt += __popc(mk);
                            }
                        }
                    }
                }
                __syncwarp();
                if (cnt <= kHardList && cnt >= keff) {
                    double d[2]; int rank[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int e = lane + 32 * h;
                        d[h] = (e < cnt) ? dist2_exact(ldg4(dense + s_pix[wid][e]), q.qx, q.qy, q.qz) : dinf;
                        rank[h] = 0;
                    }
                    for (int j = 0; j < cnt; ++j) {
                        const double dj = __shfl_sync(SD_FULL, (j < 32) ? d[0] : d[1], j & 31);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int e = lane + 32 * h;
                            rank[h] += (dj < d[h] || (dj == d[h] && j < e)) ? 1 : 0;
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) if (lane + 32 * h < cnt) s_sorted[wid][rank[h]] = d[h];
                    __syncwarp();
                    double sacc = 0.0;
                    for (int p = 0; p < keff; ++p) sacc = sacc + sqrt(s_sorted[wid][p]);
                    const double avg = sacc / (double)keff;
                    __syncwarp();
                    if (lane == 0) { J.avg[i] = avg; acc.add(avg); }
                    continue;
                }
                __syncwarp();
            }
        }
        float net[KS];
#pragma unroll
        for (int p = 0; p < KS; ++p) net[p] = finf;
        int wprev = -1, w = 16;
        float wk = finf;
        while (true) {
            warp_scan_annulus(dense, W, H, q.u, q.v, w, wprev, [&](bool valid, const float4* rp, int, int col) {
                float v = finf;
                if (valid) v = key_of(ldg4(rp + col), q.qx, q.qy, q.qz);
#pragma unroll
                for (int p = 0; p < KS; ++p) { const float lo = fminf(net[p], v); v = fmaxf(net[p], v); net[p] = lo; }
            });
            // warp-wide k-th smallest key: keff rounds of min extraction over the lanes' sorted lists
            float cp[KS];
#pragma unroll
            for (int p = 0; p < KS; ++p) cp[p] = net[p];
            uint32_t m = 0x7f800000u;
            for (int rd = 0; rd < keff; ++rd) {
                const uint32_t head = __float_as_uint(cp[0]);
                m = __reduce_min_sync(SD_FULL, head);
                const unsigned win = __ballot_sync(SD_FULL, head == m);
                if (lane == __ffs(win) - 1) {
#pragma unroll
                    for (int p = 0; p < KS - 1; ++p) cp[p] = cp[p + 1];
                    cp[KS - 1] = finf;
                }
            }
            wk = __uint_as_float(m);
            const bool covered = (q.u - w <= 0) && (q.u + w >= W - 1) && (q.v - w <= 0) && (q.v + w >= H - 1);
            if (covered) break;
            if ((double)wk * (1.0 + 2.0 * (double)kKeyErrF) <= rho2_safe(q.P, q.rr, f, (double)(w + 1))) break;
            wprev = w; w <<= 1;
        }
        // ---- collect every candidate whose key lies within the fp32 error band of the k-th key
        const float band = wk * (1.0f + 4.0f * kKeyErrF);
        int cnt = 0;
        warp_scan_annulus(dense, W, H, q.u, q.v, w, -1, [&](bool valid, const float4* rp, int row, int col) {
            bool in = false;
            if (valid) in = key_of(ldg4(rp + col), q.qx, q.qy, q.qz) <= band;
            const unsigned mk = __ballot_sync(SD_FULL, in);
            const int pos = cnt + __popc(mk & ((1u << lane) - 1u));
            if (in && pos < kHardList) s_pix[wid][pos] = row * W + col;
            cnt += __popc(mk);
        });
        __syncwarp();
        double avg;
        if (keff == 0) {
            avg = -1.0;
        } else if (cnt > kHardList || cnt < keff) {
            // massive ties (duplicated points): sequential exact selection by lane 0 over the window
            DNet<K> ex; ex.init();
            if (lane == 0) {
                for (int row = max(q.v - w, 0); row <= min(q.v + w, H - 1); ++row)
                    for (int col = max(q.u - w, 0); col <= min(q.u + w, W - 1); ++col)
                        ex.feed(dist2_exact(ldg4(dense + (size_t)row * W + col), q.qx, q.qy, q.qz));
            }
            avg = ex.sum_sqrt(keff) / (double)keff;
        } else {
            // exact fp64 distances of the <= 64 collected candidates, ranked by the whole warp
            double d[2]; int rank[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = lane + 32 * h;
                d[h] = (e < cnt) ? dist2_exact(ldg4(dense + s_pix[wid][e]), q.qx, q.qy, q.qz) : dinf;
                rank[h] = 0;
            }
            for (int j = 0; j < cnt; ++j) {
                const double dj = __shfl_sync(SD_FULL, (j < 32) ? d[0] : d[1], j & 31);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int e = lane + 32 * h;
                    rank[h] += (dj < d[h] || (dj == d[h] && j < e)) ? 1 : 0;
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) if (lane + 32 * h < cnt) s_sorted[wid][rank[h]] = d[h];
            __syncwarp();
            double s = 0.0;
            for (int p = 0; p < keff; ++p) s = s + sqrt(s_sorted[wid][p]);
            avg = s / (double)keff;
        }
        __syncwarp();
        if (lane == 0) { J.avg[i] = avg; acc.add(avg); }
    }
    acc.flush_warp(J.st);
    // ---- the last warp of the job finalises the cloud statistics (Open3D: mean over avg > 0 divided by
    //      n, Bessel std) from the exact sums of both k-NN kernels
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(&J.st->ticket, 1u) == (unsigned)warps_total - 1u) {
            __threadfence();
            OrgState* st = J.st;
            const double S = fixed70_to_double(__ldcg(&st->acc[0][0]), __ldcg(&st->acc[0][1]));
            const double Q = fixed70_to_double(__ldcg(&st->acc[1][0]), __ldcg(&st->acc[1][1]));
            const double Pn = (double)__ldcg(&st->acc[2][0]);
            st->acc[0][0] = st->acc[0][1] = st->acc[1][0] = st->acc[1][1] = st->acc[2][0] = st->acc[2][1] = 0ull;
            const double nn = (double)n;
            const double mean = (n > 0) ? S / nn : 0.0;
            double sq = (Q - 2.0 * mean * S) + Pn * mean * mean;
            if (sq < 0.0) sq = 0.0;
            const double sd_ = (n > 1) ? sqrt(sq / (nn - 1.0)) : __longlong_as_double(0x7ff8000000000000ull);
            J.stats[0] = mean; J.stats[1] = sd_; J.stats[2] = mean + J.std_ratio * sd_;
            st->ticket = 0; st->qn_knn = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// statistical filter applied to the per-pixel image
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
org_apply_sor_kernel(const OrgJob* __restrict__ jobs) {
    const OrgJob J = jobs[blockIdx.y];
    const int n = *J.n;
    const double thr = J.stats[2];
    const float inf = __int_as_float(0x7f800000);
    int alive = 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const double a = J.avg[i];
        if (a > 0.0 && a < thr) ++alive;
        else J.dense[__ldg(J.src + i)] = make_float4(inf, inf, inf, inf);
    }
    alive = warp_sum(alive);
    if (lane_id() == 0 && alive && J.n_alive) atomicAdd(J.n_alive, alive);
}

// ---------------------------------------------------------------------------------------------
// radius count
// ---------------------------------------------------------------------------------------------
struct RadiusTest {
    float r2_in, r2_out; double r2;
    __device__ __forceinline__ bool operator()(const float4& c, float qx, float qy, float qz) const {
        const float k = key_of(c, qx, qy, qz);
        if (!(k <= r2_out)) return false;               // also rejects the +inf of empty pixels
        if (k < r2_in) return true;
        return dist2_exact(c, qx, qy, qz) <= r2;
    }
};

__global__ void __launch_bounds__(kOrgThreads)
org_ror_kernel(const OrgJob* __restrict__ jobs) {
    const OrgJob J = jobs[blockIdx.y];
    const int n = *J.n;
    const int W = J.width, H = J.height;
    const double f = fabs((double)J.q23);
    const double r2 = J.radius * J.radius;
    RadiusTest test{__double2float_rd(r2 * (1.0 - 3e-6)), __double2float_ru(r2 * (1.0 + 3e-6)), r2};
    const bool sor = J.use_sor != 0;
    const double thr = sor ? J.stats[2] : 0.0;
    const int cap = J.nb_points;
    const float4* __restrict__ dense = J.dense;
    for (int i = blockIdx.x * kOrgThreads + threadIdx.x; i < n; i += gridDim.x * kOrgThreads) {
        if (sor) { const double a = J.avg[i]; if (!(a > 0.0 && a < thr)) { J.cnt[i] = 0; continue; } }
        const Query q = load_query(J, i);
        // columns: the ball's projected half-width (plus margin), rows: +-kOrgW; both scanned centre-out so
        // that the common case (more than nb_points neighbours) stops after ~nb_points tests
        const int wu = min((int)ceil(J.radius * q.rr * q.rr / (q.P * f) * 1.05) + 1, W);
        int count = 0;
        for (int t = 0; t < kOrgSide && count <= cap; ++t) {
            const int dv = (t == 0) ? 0 : ((t & 1) ? (t + 1) / 2 : -(t / 2));
            const int row = q.v + dv;
            if (row < 0 || row >= H) continue;
            const float4* rp = dense + (size_t)row * W;
            const int c0 = max(q.u - wu, 0), c1 = min(q.u + wu, W - 1);
            const float4 empty = make_float4(__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000), 0.f);
            int jl = q.u - 1, jr = q.u;
            while (count <= cap && (jl >= c0 || jr <= c1)) {
                float4 c[8];                                   // 8 independent loads per trip
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    c[h] = (jr + h <= c1) ? ldg4(rp + jr + h) : empty;
                    c[4 + h] = (jl - h >= c0) ? ldg4(rp + jl - h) : empty;
                }
#pragma unroll
                for (int h = 0; h < 8; ++h) count += test(c[h], q.qx, q.qy, q.qz) ? 1 : 0;
                jr += 4; jl -= 4;
            }
        }
        if (count > cap) { J.cnt[i] = cap + 1; continue; }
        if (r2 <= rho2_safe(q.P, q.rr, f, (double)(min(kOrgW, wu) + 1))) { J.cnt[i] = count; continue; }   // the window held the whole ball
        J.cnt[i] = 0;
        atomicAdd(&J.st->dbg[6], 1ull);
        J.queue_ror[atomicAdd(&J.st->qn_ror, 1)] = i;
    }
}

__global__ void __launch_bounds__(kOrgThreads)
org_ror_hard_kernel(const OrgJob* __restrict__ jobs) {
    const OrgJob J = jobs[blockIdx.y];
    const int W = J.width, H = J.height;
    const double f = fabs((double)J.q23);
    const double r2 = J.radius * J.radius;
    RadiusTest test{__double2float_rd(r2 * (1.0 - 3e-6)), __double2float_ru(r2 * (1.0 + 3e-6)), r2};
    const int cap = J.nb_points;
    const float4* __restrict__ dense = J.dense;
    const int lane = lane_id();
    const int warps_total = gridDim.x * (kOrgThreads / 32);
    const int qn = J.st->qn_ror;
    for (int qi = blockIdx.x * (kOrgThreads / 32) + warp_id(); qi < qn; qi += warps_total) {
        const int i = J.queue_ror[qi];
        const Query q = load_query(J, i);
        int count = 0, wprev = -1, w = 16;
        while (true) {
            // rows nearest first so that dense queries leave early
            const int r0 = max(q.v - w, 0), r1 = min(q.v + w, H - 1);
            const int cl = max(q.u - w, 0), ch = min(q.u + w, W - 1);
            const float hinf = __int_as_float(0x7f800000);
            for (int t0 = 0; t0 <= 2 * w && count <= cap; t0 += 4) {
                // four rows (nearest first) in flight per column chunk
                int rows[4]; bool inner[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int t = t0 + h;
                    const int dv = (t == 0) ? 0 : ((t & 1) ? (t + 1) / 2 : -(t / 2));
                    const int row = q.v + dv;
                    const int adv = dv < 0 ? -dv : dv;
                    rows[h] = (t <= 2 * w && row >= r0 && row <= r1) ? row : -1;
                    inner[h] = (wprev >= 0 && adv <= wprev);          // only the parts outside the previous square
                }
                for (int c0 = cl; c0 <= ch && count <= cap; c0 += 32) {
                    const int col = c0 + lane;
                    float4 c[4];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const bool skip = inner[h] && col >= q.u - wprev && col <= q.u + wprev;
                        c[h] = (rows[h] >= 0 && col <= ch && !skip) ? ldg4(dense + (size_t)rows[h] * W + col)
                                                                     : make_float4(hinf, hinf, hinf, 0.f);
                    }
#pragma unroll
                    for (int h = 0; h < 4; ++h) count += __popc(__ballot_sync(SD_FULL, test(c[h], q.qx, q.qy, q.qz)));
                }
            }
            if (count > cap) break;
            const bool covered = (q.u - w <= 0) && (q.u + w >= W - 1) && (q.v - w <= 0) && (q.v + w >= H - 1);
            if (covered || r2 <= rho2_safe(q.P, q.rr, f, (double)(w + 1))) break;
            wprev = w; w <<= 1;
        }
        if (lane == 0) J.cnt[i] = min(count, cap + 1);
    }
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(&J.st->ticket, 1u) == (unsigned)warps_total - 1u) { J.st->ticket = 0; J.st->qn_ror = 0; }
    }
}

}  // namespace sd

static dim3 org_grid(int cap, int threads, int njobs, int waves) {
    int per = (cap + threads - 1) / threads;
    int target = (148 * waves) / (njobs > 0 ? njobs : 1);
    if (target < 1) target = 1;
    if (per > target) per = target;
    if (per < 1) per = 1;
    return dim3(per, njobs);
}

int sd_launch_org_fill(float4* dense, size_t count, cudaStream_t st) {
    sd::org_fill_kernel<<<148 * 8, 256, 0, st>>>(dense, count);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

int sd_launch_org_knn(const sd::OrgJob* d_jobs, int njobs, int cap, int k, cudaStream_t st) {
    using namespace sd;
    if (njobs <= 0) return SD_OK;
    if (k < 1 || k > 32) return SD_ERR_INVALID;
    dim3 grid = org_grid(cap, kOrgThreads, njobs, 16), hard = org_grid(cap, kOrgThreads, njobs, 16);
#define SD_ORG_KNN(KS) do { org_knn_kernel<KS><<<grid, kOrgThreads, 0, st>>>(d_jobs); \
                            org_knn_hard_kernel<KS><<<hard, kOrgThreads, 0, st>>>(d_jobs); } while (0)
    if (k <= 7) SD_ORG_KNN(8);
    else if (k <= 10) SD_ORG_KNN(11);
    else if (k <= 16) SD_ORG_KNN(17);
    else if (k <= 20) SD_ORG_KNN(21);
    else SD_ORG_KNN(33);
#undef SD_ORG_KNN
    SD_LAUNCH_CHECK();
    return SD_OK;
}

int sd_launch_org_apply_sor(const sd::OrgJob* d_jobs, int njobs, int cap, cudaStream_t st) {
    using namespace sd;
    if (njobs <= 0) return SD_OK;
    org_apply_sor_kernel<<<org_grid(cap, 256 * 4, njobs, 4), 256, 0, st>>>(d_jobs);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

int sd_launch_org_ror(const sd::OrgJob* d_jobs, int njobs, int cap, cudaStream_t st) {
    using namespace sd;
    if (njobs <= 0) return SD_OK;
    org_ror_kernel<<<org_grid(cap, kOrgThreads, njobs, 16), kOrgThreads, 0, st>>>(d_jobs);
    org_ror_hard_kernel<<<org_grid(cap, kOrgThreads, njobs, 16), kOrgThreads, 0, st>>>(d_jobs);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
