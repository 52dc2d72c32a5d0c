// Uniform-grid neighbour search: the two Open3D filters of the road chain
// (semantic_depth.py:227-245; Open3D <= 0.7 RemoveStatisticalOutliers / RemoveRadiusOutliers).
//
// The reference runs a FLANN KD-tree on the host.  Here the cloud is binned into a 2-D uniform grid
// over its two widest axes (road clouds are 2.5-D slabs; the third axis is collapsed, which keeps
// every bound below a valid lower bound of the 3-D distance), points are counting-sorted by cell
// (integer atomics only: cell order inside a cell is arbitrary but the k smallest distances and the
// radius counts do not depend on it), and each point searches rings / rows of cells outwards until
// the current k-th distance (or the radius) proves that no unvisited cell can matter.  Distances are
// fp64 `(dx*dx + dy*dy) + dz*dz` without FMA, sqrt is IEEE, the k distances are summed in ascending
// order from 0.0 -- the arithmetic of FLANN's L2 functor + std::accumulate -- so the per-point
// mean distance is bit-identical to the CPU oracle.  An fp32 pre-test with a proven error margin
// rejects most candidates before the fp64 evaluation.
//
// Not HBM-bound: the sorted cloud (<= 8 MB) lives in L2; the cost is L2 latency + fp64 ALU.
#include "sd_internal.cuh"

namespace sd {

constexpr int kGridThreads = 256;
constexpr int kKnnThreads = 128;
constexpr double kSlackRel = 1e-9;     // cells: geometric slack for bounds, relative to the cell size

__device__ __forceinline__ float pick_axis(int a, float x, float y, float z) { return a == 0 ? x : (a == 1 ? y : z); }

__device__ __forceinline__ int cell_coord(double p, double o, double inv, int d) {
    double t = floor((p - o) * inv);
    int c = (t < 0.0) ? 0 : ((t >= (double)d) ? d - 1 : (int)t);
    return c;
}

// ---------------------------------------------------------------------------------------------
// 1. bounding box + grid geometry
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGridThreads)
grid_bbox_kernel(const KnnJob* __restrict__ jobs) {
    __shared__ int s_last;
    const KnnJob J = jobs[blockIdx.y];
    GridState* gs = J.gs;
    const int n = *J.n;
    uint32_t mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    for (int i = blockIdx.x * kGridThreads + threadIdx.x; i < n; i += gridDim.x * kGridThreads) {
        uint32_t k0 = f2key(__ldg(J.x + i)), k1 = f2key(__ldg(J.y + i)), k2 = f2key(__ldg(J.z + i));
        mn[0] = min(mn[0], k0); mx[0] = max(mx[0], k0);
        mn[1] = min(mn[1], k1); mx[1] = max(mx[1], k1);
        mn[2] = min(mn[2], k2); mx[2] = max(mx[2], k2);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { mn[a] = warp_min(mn[a]); mx[a] = warp_max(mx[a]); }
    if (lane_id() == 0 && n > 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(&gs->bbox[a], mn[a]); atomicMax(&gs->bbox[3 + a], mx[a]); }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&gs->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    double lo[3], ext[3];
    for (int a = 0; a < 3; ++a) {
        uint32_t kmn = __ldcg(&gs->bbox[a]), kmx = __ldcg(&gs->bbox[3 + a]);
        double l = (n > 0) ? (double)key2f(kmn) : 0.0, h = (n > 0) ? (double)key2f(kmx) : 0.0;
        if (!isfinite(l) || !isfinite(h)) { l = 0.0; h = 0.0; }   // non-finite clouds are unsupported
        lo[a] = l; ext[a] = h - l;
        gs->bbox[a] = 0xffffffffu; gs->bbox[3 + a] = 0u;
    }
    int a0 = 0, a1 = 1, a2 = 2;     // sort axes by extent, descending (stable)
    if (ext[a1] > ext[a0]) { int t = a0; a0 = a1; a1 = t; }
    if (ext[a2] > ext[a1]) { int t = a1; a1 = a2; a2 = t; }
    if (ext[a1] > ext[a0]) { int t = a0; a0 = a1; a1 = t; }
    double cell = J.cell_scale * sqrt(ext[a0] * ext[a1] / (double)(n > 0 ? n : 1));
    if (!(cell > 0.0)) cell = ext[a0] / (double)(n > 0 ? n : 1);
    if (!(cell > 0.0)) cell = 1.0;
    long long d0, d1;
    for (int it = 0; it < 64; ++it) {
        d0 = (long long)floor(ext[a0] / cell) + 1; d1 = (long long)floor(ext[a1] / cell) + 1;
        if (d0 * d1 <= (long long)J.cell_cap && d0 < (1 << 24) && d1 < (1 << 24)) break;
        cell *= 1.25;
    }
    gs->a0 = a0; gs->a1 = a1; gs->a2 = a2;
    gs->d0 = (int)d0; gs->d1 = (int)d1; gs->ncells = (int)(d0 * d1);
    gs->o0 = lo[a0]; gs->o1 = lo[a1];
    gs->cell = cell; gs->inv_cell = 1.0 / cell; gs->ext2 = ext[a2];
    gs->n = n;
    gs->ticket = 0;
}

// ---------------------------------------------------------------------------------------------
// 2. counting sort by cell: count -> exclusive scan (look-back) -> scatter
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGridThreads)
grid_count_kernel(const KnnJob* __restrict__ jobs) {
    const KnnJob J = jobs[blockIdx.y];
    const GridState g = *J.gs;
    for (int i = blockIdx.x * kGridThreads + threadIdx.x; i < g.n; i += gridDim.x * kGridThreads) {
        const float x = __ldg(J.x + i), y = __ldg(J.y + i), z = __ldg(J.z + i);
        const int c0 = cell_coord((double)pick_axis(g.a0, x, y, z), g.o0, g.inv_cell, g.d0);
        const int c1 = cell_coord((double)pick_axis(g.a1, x, y, z), g.o1, g.inv_cell, g.d1);
        const int c = c1 * g.d0 + c0;
        J.cell_of[i] = c;
        atomicAdd(&J.cell_count[c], 1);
    }
}

constexpr int kScanThreads = 512;
constexpr int kScanItems = kScanTile / kScanThreads;   // 8

__global__ void __launch_bounds__(kScanThreads)
grid_scan_kernel(const KnnJob* __restrict__ jobs) {
    __shared__ int s_scan[33];
    __shared__ int s_tile;
    __shared__ unsigned long long s_excl;
    const KnnJob J = jobs[blockIdx.y];
    const int ncells = J.gs->ncells;
    const int ntiles = ceil_div(ncells, kScanTile);
    const int tid = threadIdx.x;
    while (true) {
        if (tid == 0) s_tile = (int)atomicAdd(&J.scan_ctl->ticket, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= ntiles) break;
        const int base = tile * kScanTile + tid * kScanItems;
        int v[kScanItems]; int sum = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) { v[k] = (base + k < ncells) ? J.cell_count[base + k] : 0; sum += v[k]; }
        int total;
        const int excl = block_excl_scan(sum, s_scan, &total);
        if (warp_id() == 0) {
            unsigned long long e = lookback_exclusive(J.scan_status, tile, (unsigned long long)total);
            if (lane_id() == 0) s_excl = e;
        }
        __syncthreads();
        int run = (int)s_excl + excl;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) { if (base + k < ncells) J.cell_start[base + k] = run; run += v[k]; }
        if (tile == ntiles - 1 && tid == 0) J.cell_start[ncells] = (int)s_excl + total;
        __syncthreads();
    }
    scan_finish(J.scan_ctl, J.scan_status, ntiles, gridDim.x);
}

__global__ void __launch_bounds__(kGridThreads)
grid_scatter_kernel(const KnnJob* __restrict__ jobs) {
    const KnnJob J = jobs[blockIdx.y];
    const int n = J.gs->n;
    for (int i = blockIdx.x * kGridThreads + threadIdx.x; i < n; i += gridDim.x * kGridThreads) {
        const int c = J.cell_of[i];
        const int k = atomicSub(&J.cell_count[c], 1) - 1;     // leaves cell_count all-zero again
        const int pos = J.cell_start[c] + k;
        J.sx[pos] = __ldg(J.x + i); J.sy[pos] = __ldg(J.y + i); J.sz[pos] = __ldg(J.z + i);
        J.sorig[pos] = i;
    }
}

// ---------------------------------------------------------------------------------------------
// 3. exact kNN mean distance (Open3D RemoveStatisticalOutliers, per point)
// ---------------------------------------------------------------------------------------------
struct GridRt {
    int a0, a1, d0, d1, n;
    double o0, o1, cell, inv_cell, slack;
};
__device__ __forceinline__ GridRt load_grid(const GridState* gs) {
    GridRt g;
    g.a0 = gs->a0; g.a1 = gs->a1; g.d0 = gs->d0; g.d1 = gs->d1; g.n = gs->n;
    g.o0 = gs->o0; g.o1 = gs->o1; g.cell = gs->cell; g.inv_cell = gs->inv_cell;
    g.slack = gs->cell * kSlackRel;
    return g;
}

template <int KCAP>
struct Best {
    double d[KCAP];      // descending: d[0] is the current k-th smallest (the worst kept)
    int k;
    __device__ __forceinline__ void init(int k_) {
        k = k_;
#pragma unroll
        for (int p = 0; p < KCAP; ++p) d[p] = __longlong_as_double(0x7ff0000000000000ll);
    }
    __device__ __forceinline__ void insert(double v) {     // requires v < d[0]
#pragma unroll
        for (int p = 0; p < KCAP; ++p) {
            if (p < k) {
                const bool shift = (p + 1 < k) && (d[(p + 1 < KCAP) ? p + 1 : p] > v);
                d[p] = shift ? d[(p + 1 < KCAP) ? p + 1 : p] : (d[p] > v ? v : d[p]);
            }
        }
    }
    __device__ __forceinline__ double sum_sqrt_ascending() const {
        double s = 0.0;
#pragma unroll
        for (int p = KCAP - 1; p >= 0; --p) if (p < k) s = s + sqrt(d[p]);
        return s;
    }
};

// ---- slow exact path: fp64 keys, ring by ring (used when the fp32-keyed search cannot certify its set)
template <int KCAP>
__device__ __forceinline__ void scan_range_exact(const KnnJob& J, int s, int e, float qx, float qy, float qz, Best<KCAP>& best) {
    for (int j = s; j < e; ++j) {
        const float px = __ldg(J.sx + j), py = __ldg(J.sy + j), pz = __ldg(J.sz + j);
        const double dx = (double)px - (double)qx, dy = (double)py - (double)qy, dz = (double)pz - (double)qz;
        const double d2 = (dx * dx + dy * dy) + dz * dz;
        if (d2 < best.d[0]) best.insert(d2);
    }
}

template <int KCAP>
__device__ __noinline__ double knn_exact_sum(const KnnJob& J, const GridRt& g, float qx, float qy, float qz, double q0, double q1,
                                              int c0, int c1, int keff) {
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    Best<KCAP> best; best.init(keff);
    for (int r = 0;; ++r) {
        const int lo0 = c0 - r, hi0 = c0 + r, lo1 = c1 - r, hi1 = c1 + r;
        const int cl0 = max(lo0, 0), ch0 = min(hi0, g.d0 - 1);
        for (int row = max(lo1, 0); row <= min(hi1, g.d1 - 1); ++row) {
            const int rb = row * g.d0;
            if (row == lo1 || row == hi1) {
                scan_range_exact<KCAP>(J, J.cell_start[rb + cl0], J.cell_start[rb + ch0 + 1], qx, qy, qz, best);
            } else {
                if (lo0 >= 0) scan_range_exact<KCAP>(J, J.cell_start[rb + lo0], J.cell_start[rb + lo0 + 1], qx, qy, qz, best);
                if (hi0 < g.d0) scan_range_exact<KCAP>(J, J.cell_start[rb + hi0], J.cell_start[rb + hi0 + 1], qx, qy, qz, best);
            }
        }
        const double e_lo0 = (lo0 <= 0) ? inf : q0 - (g.o0 + (double)lo0 * g.cell);
        const double e_hi0 = (hi0 >= g.d0 - 1) ? inf : (g.o0 + (double)(hi0 + 1) * g.cell) - q0;
        const double e_lo1 = (lo1 <= 0) ? inf : q1 - (g.o1 + (double)lo1 * g.cell);
        const double e_hi1 = (hi1 >= g.d1 - 1) ? inf : (g.o1 + (double)(hi1 + 1) * g.cell) - q1;
        double lb = fmin(fmin(e_lo0, e_hi0), fmin(e_lo1, e_hi1));
        if (lb == inf) break;
        lb = lb - g.slack;
        if (lb > 0.0 && best.d[0] <= lb * lb) break;
    }
    return best.sum_sqrt_ascending();
}

// ---- fast path -------------------------------------------------------------------------------------
// Phase 1: every candidate of the search square goes through a branch-free sorted-insertion network on
//          fp32 keys (2 FMNMX per slot, all lanes active, no divergence): the k+1 smallest keys survive.
// Phase 2: the square is rescanned and the candidates whose key is within the fp32 error band of the
//          k-th key are collected (indices, shared memory, <= kListCap per query).  Every true member of
//          the exact k-set is in that list (see DESIGN.md "k-NN exactness").
// Phase 3: fp64 distances of the collected candidates, exact branch-free selection of the k smallest,
//          sqrt and ascending sum -- the arithmetic of the oracle.
// A query whose list overflows (massive ties) is redone by the fp64-keyed slow path.
constexpr int kListCap = 16;
constexpr float kKeyErr = 1.5e-6f;     // relative error bound of the fp32 squared distance

template <int KS>
struct FNet {
    float d[KS];                         // ascending: d[0] smallest
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int p = 0; p < KS; ++p) d[p] = __int_as_float(0x7f800000);
    }
    __device__ __forceinline__ void feed(float v) {
#pragma unroll
        for (int p = 0; p < KS; ++p) { const float lo = fminf(d[p], v); v = fmaxf(d[p], v); d[p] = lo; }
    }
    __device__ __forceinline__ float get(int idx) const {     // d[idx] with a runtime index (unrolled select)
        float r = d[0];
#pragma unroll
        for (int p = 1; p < KS; ++p) r = (p == idx) ? d[p] : r;
        return r;
    }
};

__device__ __forceinline__ float key_f32(const KnnJob& J, int j, float qx, float qy, float qz) {
    const float fx = __ldg(J.sx + j) - qx, fy = __ldg(J.sy + j) - qy, fz = __ldg(J.sz + j) - qz;
    return (fx * fx + fy * fy) + fz * fz;               // identical expression in phases 1 and 2
}

// visit the cells that square(r) adds to square(rold) (rold < 0: everything), row by row
template <typename F>
__device__ __forceinline__ void for_new_segments(const KnnJob& J, const GridRt& g, int c0, int c1, int r, int rold, F&& f) {
    const int lo0 = c0 - r, hi0 = c0 + r, lo1 = c1 - r, hi1 = c1 + r;
    const int cl0 = max(lo0, 0), ch0 = min(hi0, g.d0 - 1);
    for (int row = max(lo1, 0); row <= min(hi1, g.d1 - 1); ++row) {
        const int rb = row * g.d0;
        if (rold < 0 || row < c1 - rold || row > c1 + rold) {
            f(J.cell_start[rb + cl0], J.cell_start[rb + ch0 + 1]);
        } else {
            const int le = min(c0 - rold - 1, ch0), rs = max(c0 + rold + 1, cl0);
            if (le >= cl0) f(J.cell_start[rb + cl0], J.cell_start[rb + le + 1]);
            if (rs <= ch0) f(J.cell_start[rb + rs], J.cell_start[rb + ch0 + 1]);
        }
    }
}

template <int KS>
__global__ void __launch_bounds__(kKnnThreads)
knn_kernel(const KnnJob* __restrict__ jobs) {
    constexpr int K = KS - 1;
    __shared__ int s_list[kListCap][kKnnThreads];
    const KnnJob J = jobs[blockIdx.y];
    const GridRt g = load_grid(J.gs);
    const int keff = min(J.k, g.n);
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const int tid = threadIdx.x;
    U128 acc_sum{0ull, 0ull}, acc_sq{0ull, 0ull};
    unsigned long long acc_pos = 0ull;

    // warps claim 32 consecutive (cell-sorted) queries at a time from a per-job counter: sparse-region
    // queries cost 10-100x more than dense ones, static partitioning left half of the SMs idle at the end
    while (true) {
        int wbase = 0;
        if (lane_id() == 0) wbase = atomicAdd(&J.gs->work, 32);
        wbase = __shfl_sync(SD_FULL, wbase, 0);
        if (wbase >= g.n) break;
        const int i = wbase + lane_id();
        if (i >= g.n) continue;
        const float qx = __ldg(J.sx + i), qy = __ldg(J.sy + i), qz = __ldg(J.sz + i);
        const double q0 = (double)pick_axis(g.a0, qx, qy, qz), q1 = (double)pick_axis(g.a1, qx, qy, qz);
        const int c0 = cell_coord(q0, g.o0, g.inv_cell, g.d0), c1 = cell_coord(q1, g.o1, g.inv_cell, g.d1);
        // ---- phase 1
        FNet<KS> net; net.init();
        int rold = -1, r = 1;
        for (;; r <<= 1) {
            for_new_segments(J, g, c0, c1, r, rold, [&](int s, int e) {
                for (int j = s; j < e; ++j) net.feed(key_f32(J, j, qx, qy, qz));
            });
            const int lo0 = c0 - r, hi0 = c0 + r, lo1 = c1 - r, hi1 = c1 + r;
            const double e_lo0 = (lo0 <= 0) ? inf : q0 - (g.o0 + (double)lo0 * g.cell);
            const double e_hi0 = (hi0 >= g.d0 - 1) ? inf : (g.o0 + (double)(hi0 + 1) * g.cell) - q0;
            const double e_lo1 = (lo1 <= 0) ? inf : q1 - (g.o1 + (double)lo1 * g.cell);
            const double e_hi1 = (hi1 >= g.d1 - 1) ? inf : (g.o1 + (double)(hi1 + 1) * g.cell) - q1;
            double lb = fmin(fmin(e_lo0, e_hi0), fmin(e_lo1, e_hi1));
            if (lb == inf) break;                       // the square covers the whole grid
            lb = lb - g.slack;
            // every unvisited point is at least lb away: stop when the k-th key (inflated by its error) is closer
            if (lb > 0.0 && keff > 0 && (double)net.get(keff - 1) * 1.000002 <= lb * lb) break;
            rold = r;
        }
        // ---- phase 2
        const float wk = (keff > 0) ? net.get(keff - 1) : 0.f;
        const float band = wk * (1.0f + 4.0f * kKeyErr);
        int cnt = 0;
        for_new_segments(J, g, c0, c1, r, -1, [&](int s, int e) {
            for (int j = s; j < e; ++j) {
                const bool in = key_f32(J, j, qx, qy, qz) <= band;
                if (in && cnt < kListCap) s_list[cnt][tid] = j;
                cnt += in ? 1 : 0;
            }
        });
        // ---- phase 3
        double sum;
        if (cnt > kListCap || cnt < keff) {
            sum = knn_exact_sum<(K > 0 ? K : 1)>(J, g, qx, qy, qz, q0, q1, c0, c1, keff);
        } else {
            double bd[K > 0 ? K : 1];
#pragma unroll
            for (int p = 0; p < K; ++p) bd[p] = inf;
            for (int e = 0; e < cnt; ++e) {
                const int j = s_list[e][tid];
                const double dx = (double)__ldg(J.sx + j) - (double)qx, dy = (double)__ldg(J.sy + j) - (double)qy,
                             dz = (double)__ldg(J.sz + j) - (double)qz;
                double v = (dx * dx + dy * dy) + dz * dz;
#pragma unroll
                for (int p = 0; p < K; ++p) { const double lo = fmin(bd[p], v); v = fmax(bd[p], v); bd[p] = lo; }
            }
            sum = 0.0;
#pragma unroll
            for (int p = 0; p < K; ++p) if (p < keff) sum = sum + sqrt(bd[p]);
        }
        const double avg = (keff > 0) ? sum / (double)keff : -1.0;
        J.avg[__ldg(J.sorig + i)] = avg;
        J.savg[i] = avg;
        if (avg > 0.0) { acc_sum = add128(acc_sum, to_fixed70(avg)); acc_sq = add128(acc_sq, to_fixed70(avg * avg)); ++acc_pos; }
    }

    // ---- cloud statistics (Open3D: mean over avg > 0 divided by n, Bessel std): exact integer sums.
    //      Warps retire independently (no CTA barrier): the last WARP of the job finalises.
    acc_sum = warp_sum128(acc_sum); acc_sq = warp_sum128(acc_sq);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc_pos += __shfl_xor_sync(SD_FULL, acc_pos, o);
    int last = 0;
    if (lane_id() == 0) {
        if (acc_pos) {
            atomic_add128(J.gs->acc[0], acc_sum); atomic_add128(J.gs->acc[1], acc_sq);
            atomicAdd(&J.gs->acc[2][0], acc_pos);
        }
        __threadfence();
        last = (atomicAdd(&J.gs->ticket, 1u) == gridDim.x * (kKnnThreads / 32) - 1);
        if (last) {
            __threadfence();
            GridState* gs = J.gs;
            const double S = fixed70_to_double(__ldcg(&gs->acc[0][0]), __ldcg(&gs->acc[0][1]));
            const double Q = fixed70_to_double(__ldcg(&gs->acc[1][0]), __ldcg(&gs->acc[1][1]));
            const double Pn = (double)__ldcg(&gs->acc[2][0]);
            gs->acc[0][0] = gs->acc[0][1] = gs->acc[1][0] = gs->acc[1][1] = gs->acc[2][0] = gs->acc[2][1] = 0ull;
            const double n = (double)g.n;
            const double mean = (g.n > 0) ? S / n : 0.0;
            // sum over avg>0 of (avg-mean)^2 = Q - 2*mean*S + Pn*mean^2
            double sq = (Q - 2.0 * mean * S) + Pn * mean * mean;
            if (sq < 0.0) sq = 0.0;
            const double sd_ = (g.n > 1) ? sqrt(sq / (n - 1.0)) : __longlong_as_double(0x7ff8000000000000ull);
            J.stats[0] = mean; J.stats[1] = sd_; J.stats[2] = mean + J.std_ratio * sd_;
            gs->ticket = 0; gs->work = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 4. radius count with early exit (Open3D RemoveRadiusOutliers, per point)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kKnnThreads)
radius_kernel(const KnnJob* __restrict__ jobs) {
    const KnnJob J = jobs[blockIdx.y];
    const GridRt g = load_grid(J.gs);
    const double r = J.radius, r2 = r * r;
    const float r2_in = __double2float_rd(r2 * (1.0 - 3e-6)), r2_out = __double2float_ru(r2 * (1.0 + 3e-6));
    const bool sor = J.use_sor != 0;
    const double thr = sor ? J.stats[2] : 0.0;
    const int cap = J.count_cap;
    const int rows = (int)ceil(r * g.inv_cell) + 1;
    int alive_local = 0;
    while (true) {
        int wbase = 0;
        if (lane_id() == 0) wbase = atomicAdd(&J.gs->work, 32);
        wbase = __shfl_sync(SD_FULL, wbase, 0);
        if (wbase >= g.n) break;
        const int i = wbase + lane_id();
        if (i >= g.n) continue;
        const int orig = __ldg(J.sorig + i);
        if (sor) {
            const double a = J.savg[i];
            if (!(a > 0.0 && a < thr)) { J.cnt[orig] = 0; continue; }
            ++alive_local;
        }
        const float qx = __ldg(J.sx + i), qy = __ldg(J.sy + i), qz = __ldg(J.sz + i);
        const double q0 = (double)pick_axis(g.a0, qx, qy, qz), q1 = (double)pick_axis(g.a1, qx, qy, qz);
        const int c1 = cell_coord(q1, g.o1, g.inv_cell, g.d1);
        int count = 0;
        bool done = false;
        for (int t = 0; t <= 2 * rows && !done; ++t) {
            const int dr = (t == 0) ? 0 : ((t & 1) ? (t + 1) / 2 : -(t / 2));
            const int row = c1 + dr;
            if (row < 0 || row >= g.d1) continue;
            double gap = 0.0;                                // distance from q to the row's slab along a1
            if (dr > 0) gap = (g.o1 + (double)row * g.cell) - q1;
            else if (dr < 0) gap = q1 - (g.o1 + (double)(row + 1) * g.cell);
            gap -= g.slack;
            if (gap > r) continue;
            if (gap < 0.0) gap = 0.0;
            const double half = sqrt(r2 - gap * gap) + g.slack;
            const int ca = cell_coord(q0 - half, g.o0, g.inv_cell, g.d0), cb = cell_coord(q0 + half, g.o0, g.inv_cell, g.d0);
            const int rb = row * g.d0;
            const int s = J.cell_start[rb + ca], e = J.cell_start[rb + cb + 1];
            // walk outwards from the query's own position (its own index in its row, the cell straight
            // above / below it in the other rows): near candidates first, so dense queries stop after ~cap tests
            const int c0q = cell_coord(q0, g.o0, g.inv_cell, g.d0);
            int mid = (dr == 0) ? i : J.cell_start[rb + c0q];
            mid = min(max(mid, s), e);
            auto test = [&](int j) {
                const float px = __ldg(J.sx + j), py = __ldg(J.sy + j), pz = __ldg(J.sz + j);
                const float fx = px - qx, fy = py - qy, fz = pz - qz;
                const float d2f = (fx * fx + fy * fy) + fz * fz;       // fp32 pre-test, relative error < 1.5e-6
                if (d2f > r2_out) return;
                bool in = d2f < r2_in;
                if (!in) {
                    const double dx = (double)px - (double)qx, dy = (double)py - (double)qy, dz = (double)pz - (double)qz;
                    in = ((dx * dx + dy * dy) + dz * dz) <= r2;
                }
                if (in && sor) { const double a = J.savg[j]; in = (a > 0.0 && a < thr); }
                if (in) {
                    ++count;
                    if (cap >= 0 && count > cap) done = true;
                }
            };
            int jl = mid - 1, jr = mid;
            while (!done && (jl >= s || jr < e)) {
                if (jr < e) { test(jr); ++jr; }
                if (!done && jl >= s) { test(jl); --jl; }
            }
        }
        J.cnt[orig] = count;
    }
    alive_local = warp_sum(alive_local);
    if (lane_id() == 0) {
        if (sor && J.n_alive && alive_local) atomicAdd(J.n_alive, alive_local);
        __threadfence();
        if (atomicAdd(&J.gs->ticket, 1u) == gridDim.x * (kKnnThreads / 32) - 1) { J.gs->ticket = 0; J.gs->work = 0; }
    }
}

}  // namespace sd

static dim3 grid_for(int cap, int threads, int items, int njobs, int waves) {
    int per = sd::ceil_div(cap, threads * items);
    int target = (148 * waves) / (njobs > 0 ? njobs : 1);
    if (target < 1) target = 1;
    if (per > target) per = target;
    if (per < 1) per = 1;
    return dim3(per, njobs);
}

int sd_launch_grid_build(const sd::KnnJob* d_jobs, int njobs, int cap, cudaStream_t st) {
    using namespace sd;
    if (njobs <= 0) return SD_OK;
    grid_bbox_kernel<<<grid_for(cap, kGridThreads, 8, njobs, 4), kGridThreads, 0, st>>>(d_jobs);
    grid_count_kernel<<<grid_for(cap, kGridThreads, 4, njobs, 8), kGridThreads, 0, st>>>(d_jobs);
    grid_scan_kernel<<<grid_for(cap, kScanThreads, kScanItems, njobs, 3), kScanThreads, 0, st>>>(d_jobs);
    grid_scatter_kernel<<<grid_for(cap, kGridThreads, 4, njobs, 8), kGridThreads, 0, st>>>(d_jobs);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

int sd_launch_knn(const sd::KnnJob* d_jobs, int njobs, int cap, int k, cudaStream_t st) {
    using namespace sd;
    if (njobs <= 0) return SD_OK;
    if (k < 1 || k > kMaxKnnK) return SD_ERR_INVALID;
    dim3 grid = grid_for(cap, kKnnThreads, 1, njobs, 16);   // <= kKnnMaxBlocks CTAs per job
    if (k <= 3) knn_kernel<4><<<grid, kKnnThreads, 0, st>>>(d_jobs);
    else if (k <= 7) knn_kernel<8><<<grid, kKnnThreads, 0, st>>>(d_jobs);
    else if (k <= 10) knn_kernel<11><<<grid, kKnnThreads, 0, st>>>(d_jobs);
    else if (k <= 16) knn_kernel<17><<<grid, kKnnThreads, 0, st>>>(d_jobs);
    else if (k <= 20) knn_kernel<21><<<grid, kKnnThreads, 0, st>>>(d_jobs);
    else if (k <= 32) knn_kernel<33><<<grid, kKnnThreads, 0, st>>>(d_jobs);
    else knn_kernel<65><<<grid, kKnnThreads, 0, st>>>(d_jobs);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

int sd_launch_sor_stats(const sd::KnnJob*, int, cudaStream_t) { return SD_OK; }   // folded into knn_kernel

int sd_launch_radius(const sd::KnnJob* d_jobs, int njobs, int cap, cudaStream_t st) {
    using namespace sd;
    if (njobs <= 0) return SD_OK;
    radius_kernel<<<grid_for(cap, kKnnThreads, 1, njobs, 16), kKnnThreads, 0, st>>>(d_jobs);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
