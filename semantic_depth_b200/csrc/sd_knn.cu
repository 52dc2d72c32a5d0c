// Multi-resolution uniform-grid neighbour search: the two Open3D filters of the road chain
// (semantic_depth.py:227-245; Open3D <= 0.7 RemoveStatisticalOutliers / RemoveRadiusOutliers).
//
// The reference runs a FLANN KD-tree on the host.  Here the cloud is binned into a 2-D uniform grid
// over its two widest axes (road clouds are 2.5-D slabs; the third axis is collapsed, which keeps
// every bound below a valid lower bound of the 3-D distance).  A road cloud seen through a pinhole
// camera has a point density that falls with the cube of the depth (1500x between 7 m and 80 m), so
// one cell size cannot fit: the grid is kept at kLevels resolutions (cell edge x4 per level, cells of
// level L+1 are exact unions of 4x4 cells of level L), each with its own counting-sorted copy of the
// cloud (integer atomics only: the order inside a cell is arbitrary, but the k smallest distances and
// the radius counts do not depend on it).  Every query works at the finest level whose 3x3 cells
// around it hold enough points, so that its search disc always spans a handful of grid rows, each of
// which is ONE contiguous run of the level's sorted copy.
//
// Distances that reach a result are fp64 `(dx*dx + dy*dy) + dz*dz` without FMA, sqrt is IEEE, the k
// distances are summed in ascending order from 0.0 -- the arithmetic of FLANN's L2 functor +
// std::accumulate -- so the per-point mean distance is bit-identical to the CPU oracle.  fp32 keys
// are only ever used to discard candidates with a proven error margin.
//
// Not HBM-bound: the sorted copies (<= 3 x 5.4 MB per frame) live in L2/L1; the cost is instruction
// issue + L1/L2 latency.
#include "sd_internal.cuh"
#include <cstdlib>

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
    {   // the level-1 cell statistics of the previous cloud are no longer needed: back to (0, +max, 0)
        const int prev1 = gs->ld0[1] * gs->ld1[1];               // read before any block can pass the ticket below
        for (int c = blockIdx.x * kGridThreads + threadIdx.x; c < prev1; c += gridDim.x * kGridThreads)
            J.cell_box[c] = make_uint4(0u, 0xffffffffu, 0u, 0u);
        const int prevc = gs->ncells - gs->loff[1];               // coarse cells (levels >= 1) of the previous grid
        for (int c = blockIdx.x * kGridThreads + threadIdx.x; c < prevc; c += gridDim.x * kGridThreads)
            J.ybox[c] = make_uint2(0xffffffffu, 0u);
    }
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
        // all levels together hold < 1.07x the level-0 cells (+ one partial cell per row / column and level)
        if (d0 * d1 + (d0 * d1) / 8 + d0 + d1 + 64 <= (long long)J.cell_cap && d0 < (1 << 24) && d1 < (1 << 24)) break;
        cell *= 1.25;
    }
    gs->a0 = a0; gs->a1 = a1; gs->a2 = a2;
    gs->d0 = (int)d0; gs->d1 = (int)d1;
    int off = 0;
    for (int L = 0; L < kLevels; ++L) {
        const int e0 = (((int)d0 - 1) >> (2 * L)) + 1, e1 = (((int)d1 - 1) >> (2 * L)) + 1;
        gs->ld0[L] = e0; gs->ld1[L] = e1; gs->loff[L] = off;
        off += e0 * e1;
    }
    gs->ncells = off;
    gs->o0 = lo[a0]; gs->o1 = lo[a1];
    gs->cell = cell; gs->inv_cell = 1.0 / cell; gs->ext2 = ext[a2];
    gs->n = n;
    gs->ticket = 0;
}

// ---------------------------------------------------------------------------------------------
// 2. counting sort by cell at every level: count -> exclusive scan (look-back) -> scatter.
//    Levels are concatenated: cell arrays [level 0 | level 1 | level 2], sorted copies likewise, so one
//    scan yields absolute positions (level L occupies [L*n, (L+1)*n)).  Coarse cells receive thousands
//    of points each: their atomics are aggregated per warp with match_any.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int aggregated_add(int32_t* addr, int sign) {
    // the lanes that target `addr` elect a leader; returns what a private atomicAdd(addr, sign) would have
    // returned to this lane had the lanes gone one after the other
    const unsigned mask = __match_any_sync(__activemask(), (unsigned long long)addr);
    const int leader = __ffs(mask) - 1, lane = lane_id();
    const int rank = __popc(mask & ((1u << lane) - 1u)), total = __popc(mask);
    int old = 0;
    if (lane == leader) old = atomicAdd(addr, sign * total);
    old = __shfl_sync(mask, old, leader);
    return old + sign * rank;
}

__global__ void __launch_bounds__(kGridThreads)
grid_count_kernel(const KnnJob* __restrict__ jobs) {
    const KnnJob J = jobs[blockIdx.y];
    const GridState g = *J.gs;
    for (int i = blockIdx.x * kGridThreads + threadIdx.x; i < g.n; i += gridDim.x * kGridThreads) {
        const float x = __ldg(J.x + i), y = __ldg(J.y + i), z = __ldg(J.z + i);
        const int c0 = cell_coord((double)pick_axis(g.a0, x, y, z), g.o0, g.inv_cell, g.d0);
        const int c1 = cell_coord((double)pick_axis(g.a1, x, y, z), g.o1, g.inv_cell, g.d1);
        J.cell_of[i] = c1 * g.d0 + c0;
        atomicAdd(&J.cell_count[c1 * g.d0 + c0], 1);
#pragma unroll
        for (int L = 1; L < kLevels; ++L)
            aggregated_add(&J.cell_count[g.loff[L] + (c1 >> (2 * L)) * g.ld0[L] + (c0 >> (2 * L))], 1);
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
    const GridState g = *J.gs;
    for (int i = blockIdx.x * kGridThreads + threadIdx.x; i < g.n; i += gridDim.x * kGridThreads) {
        const int c = J.cell_of[i];
        const int c1 = c / g.d0, c0 = c - c1 * g.d0;
        const float x = __ldg(J.x + i), y = __ldg(J.y + i), z = __ldg(J.z + i);
        const float4 pt = make_float4(x, y, z, __int_as_float(i));
        {   // atomicSub leaves cell_count all-zero again
            const int pos = J.cell_start[c] + atomicSub(&J.cell_count[c], 1) - 1;
            J.sp[pos] = pt;
        }
        const uint32_t akey = f2key(pick_axis(g.a2, x, y, z));
#pragma unroll
        for (int L = 1; L < kLevels; ++L) {
            const int cl = g.loff[L] + (c1 >> (2 * L)) * g.ld0[L] + (c0 >> (2 * L));
            // the lanes of a coarse cell elect a leader for the position counter and for the cell's extent along the
            // collapsed axis (ordered keys; a superset bound for any later subset of the cell's points)
            const unsigned mask = __match_any_sync(__activemask(), cl);
            const int leader = __ffs(mask) - 1, lane = lane_id();
            const uint32_t kmin = __reduce_min_sync(mask, akey), kmax = __reduce_max_sync(mask, akey);
            int old = 0;
            if (lane == leader) {
                old = atomicSub(&J.cell_count[cl], __popc(mask));
                uint2* yb = J.ybox + (cl - g.loff[1]);
                atomicMin(&yb->x, kmin); atomicMax(&yb->y, kmax);
            }
            old = __shfl_sync(mask, old, leader);
            const int pos = J.cell_start[cl] + old - __popc(mask & ((1u << lane) - 1u)) - 1;
            J.sp[pos] = pt;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 3. exact kNN mean distance (Open3D RemoveStatisticalOutliers, per point)
// ---------------------------------------------------------------------------------------------
struct GridRt {
    int a0, a1, n;
    int d0[kLevels], d1[kLevels], off[kLevels];
    double o0, o1, cell, inv_cell, slack;
};
__device__ __forceinline__ GridRt load_grid(const GridState* gs) {
    GridRt g;
    g.a0 = gs->a0; g.a1 = gs->a1; g.n = gs->n;
#pragma unroll
    for (int L = 0; L < kLevels; ++L) { g.d0[L] = gs->ld0[L]; g.d1[L] = gs->ld1[L]; g.off[L] = gs->loff[L]; }
    g.o0 = gs->o0; g.o1 = gs->o1; g.cell = gs->cell; g.inv_cell = gs->inv_cell;
    g.slack = gs->cell * kSlackRel;
    return g;
}
// one level of the grid as the search loops see it
struct LevelRt {
    int d0, d1; const int32_t* cs;       // cell_start of this level (absolute positions into the sorted copies)
    double cell, inv_cell; int shift;
};
__device__ __forceinline__ LevelRt level_of(const KnnJob& J, const GridRt& g, int L) {
    LevelRt v;
    v.d0 = g.d0[0]; v.d1 = g.d1[0]; int off = 0;
#pragma unroll
    for (int l = 1; l < kLevels; ++l) if (l == L) { v.d0 = g.d0[l]; v.d1 = g.d1[l]; off = g.off[l]; }
    v.cs = J.cell_start + off;
    v.shift = 2 * L;
    const double s = (double)(1 << (2 * L));
    v.cell = g.cell * s; v.inv_cell = g.inv_cell / s;      // exact power-of-two scaling
    return v;
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

// ---- slow exact path: fp64 keys, ring by ring on the finest level (only reached on massive ties)
template <int KCAP>
__device__ __forceinline__ void scan_range_exact(const KnnJob& J, int s, int e, float qx, float qy, float qz, Best<KCAP>& best) {
    for (int j = s; j < e; ++j) {
        const float4 p = __ldg(J.sp + j);
        const double dx = (double)p.x - (double)qx, dy = (double)p.y - (double)qy, dz = (double)p.z - (double)qz;
        const double d2 = (dx * dx + dy * dy) + dz * dz;
        if (d2 < best.d[0]) best.insert(d2);
    }
}

template <int KCAP>
__device__ __noinline__ double knn_exact_sum(const KnnJob& J, const GridRt& g, float qx, float qy, float qz, double q0, double q1,
                                              int c0, int c1, int keff) {
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const int D0 = g.d0[0], D1 = g.d1[0];
    Best<KCAP> best; best.init(keff);
    for (int r = 0;; ++r) {
        const int lo0 = c0 - r, hi0 = c0 + r, lo1 = c1 - r, hi1 = c1 + r;
        const int cl0 = max(lo0, 0), ch0 = min(hi0, D0 - 1);
        for (int row = max(lo1, 0); row <= min(hi1, D1 - 1); ++row) {
            const int rb = row * D0;
            if (row == lo1 || row == hi1) {
                scan_range_exact<KCAP>(J, J.cell_start[rb + cl0], J.cell_start[rb + ch0 + 1], qx, qy, qz, best);
            } else {
                if (lo0 >= 0) scan_range_exact<KCAP>(J, J.cell_start[rb + lo0], J.cell_start[rb + lo0 + 1], qx, qy, qz, best);
                if (hi0 < D0) scan_range_exact<KCAP>(J, J.cell_start[rb + hi0], J.cell_start[rb + hi0 + 1], qx, qy, qz, best);
            }
        }
        const double e_lo0 = (lo0 <= 0) ? inf : q0 - (g.o0 + (double)lo0 * g.cell);
        const double e_hi0 = (hi0 >= D0 - 1) ? inf : (g.o0 + (double)(hi0 + 1) * g.cell) - q0;
        const double e_lo1 = (lo1 <= 0) ? inf : q1 - (g.o1 + (double)lo1 * g.cell);
        const double e_hi1 = (hi1 >= D1 - 1) ? inf : (g.o1 + (double)(hi1 + 1) * g.cell) - q1;
        double lb = fmin(fmin(e_lo0, e_hi0), fmin(e_lo1, e_hi1));
        if (lb == inf) break;
        lb = lb - g.slack;
        if (lb > 0.0 && best.d[0] <= lb * lb) break;
    }
    return best.sum_sqrt_ascending();
}

// ---- fast path -------------------------------------------------------------------------------------
// One query per lane; every loop a lane runs is either short and uniform or *flattened* (one trip per
// candidate, whatever grid row it comes from), so lanes of a warp stay busy until the lane with the most
// candidates is done.
// Level:   the finest grid level whose 3x3 cells around the query hold at least 2k points.
// Phase 1 (bound): up to 2k+48 of those points -- a window of the level's sorted copy in the query's row and
//          one in each adjacent row -- go through a branch-free sorted-insertion network on fp32 keys.  The
//          k-th smallest key U of ANY k-subset of the cloud is an upper bound of the true k-th squared
//          distance.  (Fewer than 2k points even at the coarsest level: rings of coarse cells grow.)
// Phase 2 (collect): the disc of radius sqrt(U) around the query is swept (one contiguous run of the sorted
//          copy per grid row; run bounds are staged in shared memory so the sweep is a single flattened loop)
//          with a compare + append: candidates whose key is within the fp32 error band of U go to a per-query
//          list in shared memory.  Every true member of the exact k-set is in that list (DESIGN.md "k-NN
//          exactness").  A list overflow tightens U from the listed keys and sweeps again.
// Phase 3 (select): the listed keys, truncated by 7 bits and tagged with their list slot, go through an
//          integer network (2 VIMNMX per slot); the k winners are re-evaluated in fp64 -- (dx*dx+dy*dy)+dz*dz,
//          IEEE sqrt, ascending sum from 0.0: the arithmetic of the oracle.  If the k-th and (k+1)-th
//          truncated keys are closer than two truncation cells (3e-5 relative, far above the fp32 error) or the
//          fp64 values are not ascending, the whole list is re-selected in fp64.
// A query whose list still overflows after three tightenings (massive ties) is redone by the fp64 slow path.
constexpr float kKeyErr = 1.5e-6f;     // relative error bound of the fp32 squared distance
constexpr int kSlotBits = 7;
#ifndef SD_KNN_ROWCELLS
#define SD_KNN_ROWCELLS 3
#endif
constexpr int kRowCells = SD_KNN_ROWCELLS;   // sweep level: the finest one on which the disc radius is below this many cells
constexpr int kMaxRows = 2 * (kRowCells + 1) + 2;   // row runs staged per query; wider discs go to the heavy kernel

template <int K> struct KnnCfg {
#ifndef SD_KNN_OWN
#define SD_KNN_OWN 16
#endif
#ifndef SD_KNN_SIDE
#define SD_KNN_SIDE 16
#endif
    static constexpr int own_win = K + SD_KNN_OWN;                        // phase-1 window in the query's own row
    static constexpr int side_win = K / 2 + SD_KNN_SIDE;                  // ... in each adjacent row
    static constexpr int min_fed = K + 2;                                 // fewer points than this in the 3x3 cells: coarser level
    static constexpr int feed_all = own_win + 2 * side_win;               // up to this many points in the 3x3 cells: all of them are fed
    static constexpr int feed_cap = 4 * K + 24;                           // ... ring mode
    static constexpr int list_raw = 3 * K + 10;
    static constexpr int list_cap = list_raw > 126 ? 126 : list_raw;      // phase-2 list entries per query (smem, 7-bit slots)
    // list entries are 16 bits: (row run of the query << 12) | offset inside that run (runs hold at most kExtreme = 4096
    // candidates); a small shared-memory footprint leaves the L1 to the sorted copies, which is what the sweeps wait on
    static constexpr size_t smem_bytes = ((size_t)(list_cap + 2) * sizeof(uint16_t) + (size_t)kMaxRows * sizeof(int2)) * kKnnThreads;   // + 2 spare rows
};

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

template <int KS>
struct UNetK {
    uint32_t d[KS];                      // ascending
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

__device__ __forceinline__ float key_of(const float4& p, float qx, float qy, float qz) {
    const float fx = p.x - qx, fy = p.y - qy, fz = p.z - qz;
    // fused multiply-adds: fewer roundings than the error bound kKeyErr assumes, two instructions less per candidate;
    // the SAME expression in every phase (bound, collect, select) and in the radius pre-test
    return __fmaf_rn(fz, fz, __fmaf_rn(fy, fy, fx * fx));
}
__device__ __forceinline__ float key_f32(const KnnJob& J, int j, float qx, float qy, float qz) {
    return key_of(__ldg(J.sp + j), qx, qy, qz);
}
__device__ __forceinline__ double dist2_of(const float4& p, float qx, float qy, float qz) {
    const double dx = (double)p.x - (double)qx, dy = (double)p.y - (double)qy, dz = (double)p.z - (double)qz;
    return (dx * dx + dy * dy) + dz * dz;
}
__device__ __forceinline__ double dist2_f64(const KnnJob& J, int j, float qx, float qy, float qz) {
    return dist2_of(__ldg(J.sp + j), qx, qy, qz);
}
#ifndef SD_KNN_BATCH
#define SD_KNN_BATCH 4
#endif
constexpr int kBatch = SD_KNN_BATCH;             // candidates whose loads are issued together (the loops are latency-bound)
#ifndef SD_KNN_PREFETCH
#define SD_KNN_PREFETCH 0   // bit 0: the bound phase's windows, bit 1: the collect phase's row runs are prefetched into L1 line by line
                            // before the first dependent load (memory-level parallelism instead of one miss after the other)
#endif
// prefetch the cache lines of sorted-copy entries [s, e) (16 bytes each)
__device__ __forceinline__ void prefetch_run(const float4* sp, int s, int e) {
    const char* p = reinterpret_cast<const char*>(sp + s);
    const char* const end = reinterpret_cast<const char*>(sp + e);
    for (p = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)127); p < end; p += 128)
        asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
}

// visit the cells that square(r) adds to square(rold) (rold < 0: everything) at one level, row by row
template <typename F>
__device__ __forceinline__ void for_new_segments(const LevelRt& lv, int c0, int c1, int r, int rold, F&& f) {
    const int lo0 = c0 - r, hi0 = c0 + r, lo1 = c1 - r, hi1 = c1 + r;
    const int cl0 = max(lo0, 0), ch0 = min(hi0, lv.d0 - 1);
    for (int row = max(lo1, 0); row <= min(hi1, lv.d1 - 1); ++row) {
        const int rb = row * lv.d0;
        if (rold < 0 || row < c1 - rold || row > c1 + rold) {
            f(lv.cs[rb + cl0], lv.cs[rb + ch0 + 1]);
        } else {
            const int le = min(c0 - rold - 1, ch0), rs = max(c0 + rold + 1, cl0);
            if (le >= cl0) f(lv.cs[rb + cl0], lv.cs[rb + le + 1]);
            if (rs <= ch0) f(lv.cs[rb + rs], lv.cs[rb + ch0 + 1]);
        }
    }
}

// runs of the 3x3 cells around level-0 cell (c0, c1) at level L: own row, row + 1, row - 1
__device__ __forceinline__ int block3_runs(const KnnJob& J, const GridRt& g, int L, int c0, int c1, int* s3, int* e3) {
    const int d0 = g.d0[L], d1 = g.d1[L];
    const int32_t* cs = J.cell_start + g.off[L];
    const int k0 = c0 >> (2 * L), k1 = c1 >> (2 * L);
    const int cl = max(k0 - 1, 0), ch = min(k0 + 1, d0 - 1);
    int total = 0;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int row = k1 + (t == 0 ? 0 : (t == 1 ? 1 : -1));
        int s = 0, e = 0;
        if (row >= 0 && row < d1) { s = __ldg(cs + row * d0 + cl); e = __ldg(cs + row * d0 + ch + 1); }
        s3[t] = s; e3[t] = e; total += e - s;
    }
    return total;
}

// run [s, e) of the level's sorted copy that covers the disc (q, rad) inside grid row `row` (empty when the row misses it)
__device__ __forceinline__ void row_run(const GridRt& g, const LevelRt& lv, double q0, double q1, double rad, int k1, int row,
                                        int& s, int& e) {
    double gap = 0.0;                                // distance from q to the row's slab along a1
    if (row > k1) gap = (g.o1 + (double)row * lv.cell) - q1;
    else if (row < k1) gap = q1 - (g.o1 + (double)(row + 1) * lv.cell);
    gap -= g.slack;
    s = 0; e = 0;
    if (gap > rad) return;
    if (gap < 0.0) gap = 0.0;
    const double half = (double)sqrtf(__double2float_ru(rad * rad - gap * gap)) * (1.0 + 2e-7) + g.slack;
    const int ca = cell_coord(q0 - half, g.o0, g.inv_cell, g.d0[0]) >> lv.shift,
              cb = cell_coord(q0 + half, g.o0, g.inv_cell, g.d0[0]) >> lv.shift;
    s = __ldg(lv.cs + row * lv.d0 + ca); e = __ldg(lv.cs + row * lv.d0 + cb + 1);
}
// warp-reduce the exact partial sums of the cloud statistics and add them to the job's accumulators
__device__ __forceinline__ void flush_stats(GridState* gs, U128 acc_sum, U128 acc_sq, unsigned long long acc_pos) {
    acc_sum = warp_sum128(acc_sum); acc_sq = warp_sum128(acc_sq);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc_pos += __shfl_xor_sync(SD_FULL, acc_pos, o);
    if (lane_id() == 0 && acc_pos) {
        atomic_add128(gs->acc[0], acc_sum); atomic_add128(gs->acc[1], acc_sq);
        atomicAdd(&gs->acc[2][0], acc_pos);
    }
}

#ifndef SD_KNN_HEAVY
#define SD_KNN_HEAVY 256
#endif
constexpr int kHeavy = SD_KNN_HEAVY;            // a disc with more candidates than this is swept by the whole warp
constexpr int kExtreme = 4096;                  // ... and beyond this it goes to knn_heavy_kernel (3-D cell pruning, exact re-bounding)

#ifndef SD_KNN_CLAIM
#define SD_KNN_CLAIM 1      // a warp claims 32 * SD_KNN_CLAIM consecutive cell-sorted queries and works through them chunk by chunk
#endif
constexpr int kClaim = SD_KNN_CLAIM;
#ifndef SD_KNN_WAVES
#define SD_KNN_WAVES 7      // CTAs launched per SM = the resident set (CTAs claim work until the queue is empty).  Measured: 16 waves give
                            // the same kernel time alone (1.02 vs 1.03 ms) but a slower pipelined step (2 825 vs 2 888 frames/s): surplus
                            // CTAs keep every SM's slots taken until the queue is empty and delay the other batches' kernels
#endif
#ifndef SD_KNN_GRID_SMEM
#define SD_KNN_GRID_SMEM 1  // the CTA-uniform grid geometry lives in shared memory (an LDS at each use) instead of ~20 registers per thread:
                            // 80 / 72 registers without / with 24 bytes of spill instead of 96, i.e. 6 / 7 resident CTAs per SM instead of 5
#endif
#ifndef SD_KNN_MINB
#define SD_KNN_MINB 7       // resident CTAs per SM.  The search kernels are bound by dependent-load latency at low occupancy, so warps in
                            // flight are what pays.  Measured on B200 (5-frame batch, kernel alone / pipelined step): geometry in registers,
                            // 5 CTAs at 96 registers 1.01-1.03 ms / 2 944-2 953 frames/s (6 CTAs spill: 1.12 ms); geometry in shared memory,
                            // 5 CTAs 0.975 ms / 2 991, 6 CTAs 0.990 / 3 032, 7 CTAs 0.976-0.98 / 3 070, 8 CTAs (64 registers, 77 KB of L1 left)
                            // 1.009 / 2 934
#endif
template <int KS>
__global__ void __launch_bounds__(kKnnThreads, (KS <= 11 ? SD_KNN_MINB : 1))
knn_kernel(const KnnJob* __restrict__ jobs) {
    constexpr int K = KS - 1;
    constexpr int KN = K > 0 ? K : 1;
    using Cfg = KnnCfg<KN>;
    constexpr int kListCap = Cfg::list_cap;
    extern __shared__ int2 s_dyn[];
    int2 (*s_seg)[kKnnThreads] = reinterpret_cast<int2 (*)[kKnnThreads]>(s_dyn);                   // (start, end) of a row run
    uint16_t (*s_list)[kKnnThreads] = reinterpret_cast<uint16_t (*)[kKnnThreads]>(s_dyn + kMaxRows * kKnnThreads);   // candidates: (run << 12) | offset
    auto listed = [&](int e, int t) -> int { const unsigned v = s_list[e][t]; return s_seg[v >> 12][t].x + (int)(v & 4095u); };
#if SD_KNN_GRID_SMEM >= 2
    __shared__ KnnJob s_job;
    if (threadIdx.x == 0) s_job = jobs[blockIdx.y];
    __syncthreads();
    const KnnJob& J = s_job;
#else
    const KnnJob J = jobs[blockIdx.y];
#endif
#if SD_KNN_GRID_SMEM
    // the grid geometry is CTA-uniform: kept in shared memory it costs an LDS where it is used instead of ~20 registers
    // per thread for the whole kernel
    __shared__ GridRt s_grid;
    if (threadIdx.x == 0) s_grid = load_grid(J.gs);
    __syncthreads();
    const GridRt& g = s_grid;
#else
    const GridRt g = load_grid(J.gs);
#endif
    const int keff = min(J.k, g.n);
    const int need = min(Cfg::min_fed, g.n);
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const int tid = threadIdx.x;
    U128 acc_sum{0ull, 0ull}, acc_sq{0ull, 0ull};
    unsigned long long acc_pos = 0ull;
    // warps claim 32 consecutive (cell-sorted) queries at a time from a per-job counter
    while (true) {
      int cbase = 0;
      if (lane_id() == 0) cbase = atomicAdd(&J.gs->work, 32 * kClaim);
      cbase = __shfl_sync(SD_FULL, cbase, 0);
      if (cbase >= g.n) break;
      for (int sub = 0; sub < kClaim; ++sub) {
        const int wbase = cbase + 32 * sub;
        if (wbase >= g.n) break;
        const bool valid = wbase + lane_id() < g.n;              // lanes past the end shadow the last query (warp stays converged)
        const int i = valid ? wbase + lane_id() : g.n - 1;
        const float4 qp = __ldg(J.sp + i);
        const float qx = qp.x, qy = qp.y, qz = qp.z;
        const double q0 = (double)pick_axis(g.a0, qx, qy, qz), q1 = (double)pick_axis(g.a1, qx, qy, qz);
        const int c0 = cell_coord(q0, g.o0, g.inv_cell, g.d0[0]), c1 = cell_coord(q1, g.o1, g.inv_cell, g.d1[0]);
        // ---- sampling level: finest one with at least k + 2 points in the 3x3 cells (all levels are probed at once)
        int s3[3], e3[3], L = 0, tot3 = 0;
        {
            int sa[kLevels][3], ea[kLevels][3], tot[kLevels];
#pragma unroll
            for (int l = 0; l < kLevels; ++l) tot[l] = block3_runs(J, g, l, c0, c1, sa[l], ea[l]);
            L = kLevels - 1;
#pragma unroll
            for (int l = kLevels - 2; l >= 0; --l) if (tot[l] >= need) L = l;
            tot3 = tot[0];
#pragma unroll
            for (int l = 1; l < kLevels; ++l) if (l == L) tot3 = tot[l];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                s3[t] = sa[0][t]; e3[t] = ea[0][t];
#pragma unroll
                for (int l = 1; l < kLevels; ++l) if (l == L) { s3[t] = sa[l][t]; e3[t] = ea[l][t]; }
            }
        }
        // ---- phase 1: upper bound of the k-th squared distance from the nearest cells
        FNet<KN> net; net.init();
        int fed = 0;
        {
            // few points in the 3x3 cells: all of them; otherwise three windows: own row centred on the query's
            // cell (on the query itself at level 0), adjacent rows centred on their run
            int ws[3], we[3];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const int win = (tot3 <= Cfg::feed_all) ? (e3[t] - s3[t]) : ((t == 0) ? Cfg::own_win : Cfg::side_win);
                int mid = (s3[t] + e3[t]) >> 1;
                if (t == 0 && L == 0) mid = i;
                const int a = max(s3[t], min(mid - win / 2, e3[t] - win));
                ws[t] = a; we[t] = min(e3[t], a + win);
            }
            const int n0 = we[0] - ws[0], n1 = we[1] - ws[1], n2 = we[2] - ws[2];
            fed = n0 + n1 + n2;
            if (SD_KNN_PREFETCH & 1) {
#pragma unroll
                for (int t = 0; t < 3; ++t) prefetch_run(J.sp, ws[t], we[t]);
            }
            for (int t0 = 0; t0 < fed; t0 += kBatch) {            // flattened over the three windows, kBatch loads in flight
                float4 c[kBatch];
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const int t = min(t0 + u, fed - 1);
                    const int j = (t < n0) ? ws[0] + t : ((t < n0 + n1) ? ws[1] + (t - n0) : ws[2] + (t - n0 - n1));
                    c[u] = __ldg(J.sp + j);
                }
#pragma unroll
                for (int u = 0; u < kBatch; ++u)
                    net.feed((t0 + u < fed) ? key_of(c[u], qx, qy, qz) : __int_as_float(0x7f800000));
            }
            if (fed < need) {                                     // sparse even at the coarsest level: grow rings there
                const LevelRt lv = level_of(J, g, L);
                const int k0 = c0 >> lv.shift, k1 = c1 >> lv.shift;
                net.init(); fed = 0;
                int rold = -1;
                for (int r = 1;; r += max(1, r >> 2)) {
                    for_new_segments(lv, k0, k1, r, rold, [&](int s, int e) {
                        const int e2 = min(e, s + (Cfg::feed_cap - fed));
                        for (int j = s; j < e2; ++j) net.feed(key_f32(J, j, qx, qy, qz));
                        fed += max(e2 - s, 0);
                    });
                    if (fed >= need) break;
                    if (k0 - r <= 0 && k0 + r >= lv.d0 - 1 && k1 - r <= 0 && k1 + r >= lv.d1 - 1) break;   // whole grid seen
                    rold = r;
                }
            }
        }
        // ---- phase 2: collect everything inside the bound.  Lanes with an ordinary disc sweep it themselves; a lane
        //      whose disc holds many candidates (an outlier above / below a dense region: the grid is 2-D) is served by
        //      the whole warp, 32 candidates per step.  A list overflow tightens the bound from the listed keys and
        //      sweeps once more.  Extreme discs (more than kExtreme candidates or more than kMaxRows rows) and lists that
        //      still overflow go to knn_heavy_kernel through the job's queue, with the bound.
        int cnt = 0;
        bool queued = false;
        float queue_band = __int_as_float(0x7f800000);
        {
            bool todo = (keff > 0 && fed >= keff);
            float band = todo ? net.get(keff - 1) * (1.0f + 4.0f * kKeyErr) : __int_as_float(0x7f800000);
            if (!todo && keff > 0) queued = true;          // no bound at all (cannot happen for keff <= n)
            for (int attempt = 0;; ++attempt) {
                const double rad = sqrt((double)band) * (1.0 + 1e-9) + g.slack;
                // sweep level: the finest one on which the disc spans at most 7 rows
                int Ls = kLevels - 1;
#pragma unroll
                for (int l = kLevels - 2; l >= 0; --l) if (rad * g.inv_cell < (double)(kRowCells << (2 * l))) Ls = l;
                const LevelRt lv = level_of(J, g, Ls);
                const int k1 = c1 >> lv.shift;
                const int R = (int)(rad * lv.inv_cell) + 1;
                const int rlo = max(k1 - R, 0), rhi = min(k1 + R, lv.d1 - 1);
                int nruns = 0, total = 0;
                bool heavy = false;
                if (todo) {
                    if (rhi - rlo < kMaxRows) {
                        for (int row = rlo; row <= rhi; ++row) {
                            int s, e; row_run(g, lv, q0, q1, rad, k1, row, s, e);
                            if (e > s) { s_seg[nruns][tid] = make_int2(s, e); ++nruns; total += e - s; }
                        }
                        heavy = total > kHeavy;
                        if (total > kExtreme) { queued = true; todo = false; }
                    } else { queued = true; todo = false; }
                }
                if (todo) cnt = 0;
                if (todo && !heavy) {
                    // the list is filled through a saturating pointer: entry k goes to row min(k, kListCap + 1), so a pointer
                    // that ends beyond row kListCap means "more than kListCap hits" (rows kListCap, kListCap + 1 are spare)
                    uint16_t* const lbase = &s_list[0][tid];
                    int lp = 0;                                              // element offset of the next entry
                    if (SD_KNN_PREFETCH & 2) {
                        for (int ri = 0; ri < nruns; ++ri) { const int2 se = s_seg[ri][tid]; prefetch_run(J.sp, se.x, se.y); }
                    }
                    constexpr int lp_max = (kListCap + 1) * kKnnThreads;
                    for (int ri = 0; ri < nruns; ++ri) {
                        const int2 se = s_seg[ri][tid];
                        const int tag = (ri << 12) - se.x;                   // entry = (run << 12) | offset = tag + candidate index
                        int j0 = se.x;
                        for (; j0 + kBatch <= se.y; j0 += kBatch) {          // full batches need no bound checks
                            const float4* __restrict__ pc = J.sp + j0;
                            float4 c[kBatch];
#pragma unroll
                            for (int u = 0; u < kBatch; ++u) c[u] = __ldg(pc + u);
#pragma unroll
                            for (int u = 0; u < kBatch; ++u) {
                                if (key_of(c[u], qx, qy, qz) <= band) { lbase[lp] = (uint16_t)(tag + j0 + u); lp = min(lp + kKnnThreads, lp_max); }
                            }
                        }
                        if (j0 < se.y) {                                     // the tail as one masked batch: one round trip, not up to kBatch - 1
                            float4 c[kBatch];
#pragma unroll
                            for (int u = 0; u < kBatch - 1; ++u) c[u] = __ldg(J.sp + min(j0 + u, se.y - 1));
#pragma unroll
                            for (int u = 0; u < kBatch - 1; ++u) {
                                if (j0 + u < se.y && key_of(c[u], qx, qy, qz) <= band) { lbase[lp] = (uint16_t)(tag + j0 + u); lp = min(lp + kKnnThreads, lp_max); }
                            }
                        }
                    }
                    cnt = lp / kKnnThreads;                                  // min(hits, kListCap + 1)
                }
                // heavy lanes, one after the other, all 32 lanes on each
                for (unsigned hm = __ballot_sync(SD_FULL, todo && heavy); hm; hm &= hm - 1) {
                    const int ld = __ffs(hm) - 1;
                    const float hx = __shfl_sync(SD_FULL, qx, ld), hy = __shfl_sync(SD_FULL, qy, ld), hz = __shfl_sync(SD_FULL, qz, ld);
                    const float hband = __shfl_sync(SD_FULL, band, ld);
                    const int hnr = __shfl_sync(SD_FULL, nruns, ld);
                    const int htid = (tid & ~31) + ld;
                    int hcnt = 0;
                    __syncwarp();                                        // the lane's staged runs are read by the whole warp
                    for (int ri = 0; ri < hnr; ++ri) {
                        const int2 se = s_seg[ri][htid];
                        for (int j0 = se.x; j0 < se.y; j0 += 32) {
                            const int j = j0 + lane_id();
                            const bool in = (j < se.y) && key_of(__ldg(J.sp + min(j, se.y - 1)), hx, hy, hz) <= hband;
                            const unsigned bm = __ballot_sync(SD_FULL, in);
                            const int slot = hcnt + __popc(bm & ((1u << lane_id()) - 1u));
                            if (in && slot < kListCap) s_list[slot][htid] = (uint16_t)((ri << 12) | (j - se.x));
                            hcnt += __popc(bm);
                        }
                    }
                    __syncwarp();
                    if (lane_id() == ld) cnt = hcnt;
                }
                // overflow: the k-th smallest of the listed kListCap (>= k) keys is a tighter bound (one retry, then the queue)
                if (todo && cnt > kListCap && attempt >= 1) { queued = true; todo = false; }
                todo = todo && cnt > kListCap;
                if (!__any_sync(SD_FULL, todo)) break;
                if (todo) {
                    net.init();
                    for (int e = 0; e < kListCap; ++e) net.feed(key_f32(J, listed(e, tid), qx, qy, qz));
                    band = fminf(band, net.get(keff - 1) * (1.0f + 4.0f * kKeyErr));
                }
            }
            queue_band = band;
        }
        {
            queued = queued && valid;
            const unsigned qm = __ballot_sync(SD_FULL, queued);
            if (qm) {
                int qbase = 0;
                if (lane_id() == __ffs(qm) - 1) qbase = atomicAdd(&J.gs->qn, __popc(qm));
                qbase = __shfl_sync(SD_FULL, qbase, __ffs(qm) - 1);
                if (queued) {
                    const int slot = qbase + __popc(qm & ((1u << lane_id()) - 1u));
                    J.queue[slot] = i; J.queue_band[slot] = queue_band;
                }
            }
        }
        if (queued || !valid) continue;             // lanes past the end only shadowed the last query
        // ---- phase 3
        double sum = 0.0;
        if (keff > 0) {
            UNetK<KN + 1> un; un.init();
            for (int e0 = 0; e0 < cnt; e0 += kBatch) {
                float4 c[kBatch];
#pragma unroll
                for (int u = 0; u < kBatch; ++u) c[u] = __ldg(J.sp + listed(min(e0 + u, cnt - 1), tid));
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const uint32_t kb = __float_as_uint(key_of(c[u], qx, qy, qz));
                    un.feed((e0 + u < cnt) ? (((kb >> kSlotBits) << kSlotBits) | (uint32_t)(e0 + u)) : 0xffffffffu);
                }
            }
            double bd[KN];
            bool ok = true;
            sum = 0.0;
#pragma unroll
            for (int p0 = 0; p0 < K; p0 += kBatch) {              // the winners, kBatch loads in flight
                float4 w[kBatch];
#pragma unroll
                for (int u = 0; u < kBatch; ++u)
                    if (p0 + u < K) w[u] = __ldg(J.sp + listed((p0 + u < keff) ? (int)(un.d[p0 + u] & ((1u << kSlotBits) - 1u)) : 0, tid));
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    if (p0 + u < K) {
                        const int p = p0 + u;
                        bd[p] = inf;
                        if (p < keff) {
                            bd[p] = dist2_of(w[u], qx, qy, qz);
                            if (p > 0) ok = ok && (bd[p - 1] <= bd[p]);
                            sum = sum + sqrt(bd[p]);
                        }
                    }
                }
            }
            // membership is certain when the (k+1)-th truncated key is at least two cells above the k-th
            const uint32_t ka = un.get(keff - 1) >> kSlotBits, kb1 = un.get(keff) >> kSlotBits;
            ok = ok && (cnt == keff || kb1 >= ka + 2u);
            if (!ok) {                                            // near tie: exact selection over the whole list
#pragma unroll
                for (int p = 0; p < K; ++p) bd[p] = inf;
                for (int e = 0; e < cnt; ++e) {
                    double v = dist2_f64(J, listed(e, tid), qx, qy, qz);
#pragma unroll
                    for (int p = 0; p < K; ++p) { const double lo = fmin(bd[p], v); v = fmax(bd[p], v); bd[p] = lo; }
                }
                sum = 0.0;
#pragma unroll
                for (int p = 0; p < K; ++p) if (p < keff) sum = sum + sqrt(bd[p]);
            }
        }
        const double avg = (keff > 0) ? sum / (double)keff : -1.0;
        if (valid) J.avg[__float_as_int(qp.w)] = avg;
        if (valid && avg > 0.0) { acc_sum = add128(acc_sum, to_fixed70(avg)); acc_sq = add128(acc_sq, to_fixed70(avg * avg)); ++acc_pos; }
      }
    }

    // ---- partial cloud statistics (exact integer sums); knn_heavy_kernel adds its queries and finalises
    flush_stats(J.gs, acc_sum, acc_sq, acc_pos);
}

// ---- heavy queries ----------------------------------------------------------------------------------------------------
// One queued query per warp at a time.  The cells of the disc's bounding square (level 1 for discs up to 6 level-1 cells in
// radius, else level 2) are tested against the query as 3-D boxes -- cell x extent of its points along the collapsed axis --
// one cell per lane; only cells whose box reaches into the ball are swept, 32 candidates per step.
//   collect : in-band candidates are ballot-appended to the warp's list;
//   select  : (after an overflow) every in-band key goes through per-lane networks, the global k-th smallest key is popped
//             from the 32 networks and becomes the exact bound for a second collect;
//   answer  : fp64 distances of the listed candidates, one or two per lane; the k smallest are popped in ascending order
//             and their square roots summed from 0.0 by lane 0 -- the oracle's arithmetic.
// A list that still overflows (more than kHeavyList - k exact ties) goes to the fp64 ring search.  The last warp finalises
// the cloud statistics of the job.
constexpr int kHeavyList = 64;
constexpr int kHeavyWarps = 4;

template <typename Visit>
__device__ __forceinline__ void heavy_for_cells(const KnnJob& J, const GridRt& g, int ga2, float hx, float hy, float hz,
                                                double h0, double h1, float band, Visit&& visit) {
    const double rad = sqrt((double)band) * (1.0 + 1e-9) + g.slack;
    if (rad * g.inv_cell < 6.0) {
        // small disc (under 1.5 level-1 cells): one run per grid row on the finest level where it spans at most 7 rows
        int Ls = kLevels - 1;
#pragma unroll
        for (int l = kLevels - 2; l >= 0; --l) if (rad * g.inv_cell < (double)(kRowCells << (2 * l))) Ls = l;
        const LevelRt lv = level_of(J, g, Ls);
        const int k1 = cell_coord(h1, g.o1, g.inv_cell, g.d1[0]) >> lv.shift;
        const int R = (int)(rad * lv.inv_cell) + 1;
        const int rlo = max(k1 - R, 0), rhi = min(k1 + R, lv.d1 - 1);
        int ms = 0, me = 0;
        if (rlo + lane_id() <= rhi) row_run(g, lv, h0, h1, rad, k1, rlo + lane_id(), ms, me);     // at most 9 rows
        for (unsigned am = __ballot_sync(SD_FULL, me > ms); am; am &= am - 1) {
            const int r = __ffs(am) - 1;
            visit(__shfl_sync(SD_FULL, ms, r), __shfl_sync(SD_FULL, me, r));
        }
        return;
    }
    const int hL = (rad * g.inv_cell < 24.0) ? 1 : 2;
    const LevelRt hv = level_of(J, g, hL);
    const uint2* ybox = J.ybox + (g.off[hL] - g.off[1]);
    const double ha = (double)pick_axis(ga2, hx, hy, hz);
    const int x0 = cell_coord(h0 - rad, g.o0, g.inv_cell, g.d0[0]) >> hv.shift, x1 = cell_coord(h0 + rad, g.o0, g.inv_cell, g.d0[0]) >> hv.shift;
    const int y0 = cell_coord(h1 - rad, g.o1, g.inv_cell, g.d1[0]) >> hv.shift, y1 = cell_coord(h1 + rad, g.o1, g.inv_cell, g.d1[0]) >> hv.shift;
    const int ncw = x1 - x0 + 1, total_cells = ncw * (y1 - y0 + 1);
    for (int base = 0; base < total_cells; base += 32) {
        const int ci = base + lane_id();
        int ms = 0, me = 0;
        if (ci < total_cells) {
            const int cy = y0 + ci / ncw, cx = x0 + (ci - (ci / ncw) * ncw);
            const int cell = cy * hv.d0 + cx;
            const int s = __ldg(hv.cs + cell), e = __ldg(hv.cs + cell + 1);
            if (e > s) {
                const uint2 yb = __ldg(ybox + cell);
                const double lo0 = g.o0 + (double)cx * hv.cell, lo1 = g.o1 + (double)cy * hv.cell;
                const double d0 = fmax(0.0, fmax(lo0 - h0, h0 - (lo0 + hv.cell)) - g.slack);
                const double d1 = fmax(0.0, fmax(lo1 - h1, h1 - (lo1 + hv.cell)) - g.slack);
                const double d2 = fmax(0.0, fmax((double)key2f(yb.x) - ha, ha - (double)key2f(yb.y)));
                if ((d0 * d0 + d1 * d1 + d2 * d2) * (1.0 - 1e-9) <= (double)band) { ms = s; me = e; }
            }
        }
        for (unsigned am = __ballot_sync(SD_FULL, me > ms); am; am &= am - 1) {
            const int r = __ffs(am) - 1;
            visit(__shfl_sync(SD_FULL, ms, r), __shfl_sync(SD_FULL, me, r));
        }
    }
}

#ifndef SD_KNN_HEAVY_MINB
#define SD_KNN_HEAVY_MINB 6   // resident CTAs per SM the register allocation is bounded for: 80 registers, the launch (6 CTAs per SM) is one wave
                              // (measured: 3 143-3 155 frames/s against 3 140-3 142 unbounded at 112 registers)
#endif
template <int KS>
__global__ void __launch_bounds__(kHeavyWarps * 32, (KS <= 11 ? SD_KNN_HEAVY_MINB : 1))
knn_heavy_kernel(const KnnJob* __restrict__ jobs) {
    constexpr int K = KS - 1;
    constexpr int KN = K > 0 ? K : 1;
    __shared__ int s_hl[kHeavyWarps][kHeavyList];
    const KnnJob J = jobs[blockIdx.y];
#if SD_KNN_GRID_SMEM
    __shared__ GridRt s_grid;
    if (threadIdx.x == 0) s_grid = load_grid(J.gs);
    __syncthreads();
    const GridRt& g = s_grid;
#else
    const GridRt g = load_grid(J.gs);
#endif
    const int keff = min(J.k, g.n);
    const int ga2 = J.gs->a2;
    const int w = warp_id(), lane = lane_id();
    const int qn = J.gs->qn;                                      // complete: the main kernel has finished
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    U128 acc_sum{0ull, 0ull}, acc_sq{0ull, 0ull};
    unsigned long long acc_pos = 0ull;
    while (true) {
        int e = 0;
        if (lane == 0) e = atomicAdd(&J.gs->qhead, 1);
        e = __shfl_sync(SD_FULL, e, 0);
        if (e >= qn) break;
        const int i = J.queue[e];
        float band = J.queue_band[e];
        const float4 qp = __ldg(J.sp + i);
        const float hx = qp.x, hy = qp.y, hz = qp.z;
        const double h0 = (double)pick_axis(g.a0, hx, hy, hz), h1 = (double)pick_axis(g.a1, hx, hy, hz);
        int cnt = 0;
        for (int pass = 0; pass < 3; ++pass) {
            if (pass == 1) {
                // select: exact k-th smallest key of the whole disc
                FNet<KN> net; net.init();
                heavy_for_cells(J, g, ga2, hx, hy, hz, h0, h1, band, [&](int s, int e2) {
                    for (int j = s + lane; j < e2; j += 32) {
                        const float key = key_of(__ldg(J.sp + j), hx, hy, hz);
                        if (key <= band) net.feed(key);
                    }
                });
                int head = 0; uint32_t kth = 0x7f800000u;
                for (int t = 0; t < keff; ++t) {
                    const uint32_t mine = (head < KN) ? __float_as_uint(net.get(head)) : 0x7f800000u;   // keys >= 0: bit order = value order
                    kth = __reduce_min_sync(SD_FULL, mine);
                    const unsigned who = __ballot_sync(SD_FULL, mine == kth);
                    if (lane == __ffs(who) - 1) ++head;
                }
                if (kth < 0x7f800000u) band = fminf(band, __uint_as_float(kth) * (1.0f + 4.0f * kKeyErr));
                continue;
            }
            cnt = 0;
            heavy_for_cells(J, g, ga2, hx, hy, hz, h0, h1, band, [&](int s, int e2) {
                for (int j0 = s; j0 < e2; j0 += 32) {
                    const int j = j0 + lane;
                    const bool in = (j < e2) && key_of(__ldg(J.sp + min(j, e2 - 1)), hx, hy, hz) <= band;
                    const unsigned bm = __ballot_sync(SD_FULL, in);
                    const int slot = cnt + __popc(bm & ((1u << lane) - 1u));
                    if (in && slot < kHeavyList) s_hl[w][slot] = j;
                    cnt += __popc(bm);
                }
            });
            __syncwarp();
            if (cnt <= kHeavyList) break;
        }
        double sum = 0.0;
        if (cnt > kHeavyList || cnt < keff) {                     // massive ties (or no finite bound): fp64 ring search
            if (lane == 0) {
                const int c0 = cell_coord(h0, g.o0, g.inv_cell, g.d0[0]), c1 = cell_coord(h1, g.o1, g.inv_cell, g.d1[0]);
                sum = knn_exact_sum<KN>(J, g, hx, hy, hz, h0, h1, c0, c1, keff);
            }
        } else {
            // answer: up to two listed candidates per lane in fp64, k pops of the global minimum in ascending order
            double v0 = inf, v1 = inf;
            if (lane < cnt) v0 = dist2_f64(J, s_hl[w][lane], hx, hy, hz);
            if (lane + 32 < cnt) v1 = dist2_f64(J, s_hl[w][lane + 32], hx, hy, hz);
            if (v1 < v0) { const double t = v0; v0 = v1; v1 = t; }
            for (int t = 0; t < keff; ++t) {
                double m = v0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(SD_FULL, m, o));
                const unsigned who = __ballot_sync(SD_FULL, v0 == m);
                if (lane == __ffs(who) - 1) { v0 = v1; v1 = inf; }
                sum = sum + sqrt(m);                              // every lane keeps the same running sum
            }
        }
        sum = __shfl_sync(SD_FULL, sum, 0);
        if (lane == 0) {
            const double avg = (keff > 0) ? sum / (double)keff : -1.0;
            J.avg[__float_as_int(qp.w)] = avg;
            if (avg > 0.0) { acc_sum = add128(acc_sum, to_fixed70(avg)); acc_sq = add128(acc_sq, to_fixed70(avg * avg)); ++acc_pos; }
        }
        __syncwarp();
    }
    // ---- cloud statistics (Open3D: mean over avg > 0 divided by n, Bessel std): exact integer sums.
    //      Warps retire independently (no CTA barrier): the last WARP of the job finalises and resets the job's counters.
    flush_stats(J.gs, acc_sum, acc_sq, acc_pos);
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(&J.gs->ticket, 1u) == gridDim.x * kHeavyWarps - 1) {
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
            gs->ticket = 0; gs->work = 0; gs->qn = 0; gs->qhead = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 3b. statistical filter applied to the sorted copies of every level: dead points get x = +inf, so every
//     later distance test against them fails without a second lookup (the radius kernel reads x anyway)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGridThreads)
sor_mark_kernel(const KnnJob* __restrict__ jobs) {
    const KnnJob J = jobs[blockIdx.y];
    const GridState g = *J.gs;
    const int n = g.n;
    const bool sor = J.use_sor != 0;
    const double thr = sor ? J.stats[2] : 0.0;
    const float finf = __int_as_float(0x7f800000);
    int alive_local = 0;
    // levels the radius search will read: 0, 1 (whole-cell shortcut) and the coarser ones only when it counts there
    int nlev = 2;
#pragma unroll
    for (int l = 2; l < kLevels; ++l) if (g.cell * (double)(1 << (2 * l)) <= 0.25 * J.radius) nlev = l + 1;
    for (int p0 = blockIdx.x * kGridThreads; p0 < nlev * n; p0 += gridDim.x * kGridThreads) {   // warp-uniform trip count
        const int p = p0 + threadIdx.x;
        bool alive = false; float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < nlev * n) {
            q = J.sp[p];
            alive = true;
            if (sor) {
                const int orig = __float_as_int(q.w);
                const double a = J.avg[orig];
                alive = (a > 0.0 && a < thr);
                if (!alive) J.sp[p].x = finf;
                if (p < n) { if (!alive) J.cnt[orig] = 0; alive_local += alive ? 1 : 0; }
            }
        }
        // per level-1 cell: number of alive points and their extent along the collapsed axis (entries are sorted by cell:
        // the lanes of a cell reduce among themselves, one lane issues the three atomics)
        const bool l1 = alive && p >= n && p < 2 * n;
        int cell = -1; uint32_t key = 0u;
        if (l1) {
            const int c0 = cell_coord((double)pick_axis(g.a0, q.x, q.y, q.z), g.o0, g.inv_cell, g.d0) >> 2;
            const int c1 = cell_coord((double)pick_axis(g.a1, q.x, q.y, q.z), g.o1, g.inv_cell, g.d1) >> 2;
            cell = c1 * g.ld0[1] + c0;
            key = f2key(pick_axis(g.a2, q.x, q.y, q.z));
        }
        const unsigned grp = __match_any_sync(SD_FULL, cell);
        const uint32_t kmin = __reduce_min_sync(grp, l1 ? key : 0xffffffffu);
        const uint32_t kmax = __reduce_max_sync(grp, l1 ? key : 0u);
        if (l1 && lane_id() == __ffs(grp) - 1) {
            uint4* box = J.cell_box + cell;
            atomicAdd(&box->x, (unsigned)__popc(grp)); atomicMin(&box->y, kmin); atomicMax(&box->z, kmax);
        }
    }
    alive_local = warp_sum(alive_local);
    if (lane_id() == 0 && J.n_alive && alive_local) atomicAdd(J.n_alive, alive_local);
}

// ---------------------------------------------------------------------------------------------
// 4. radius count with early exit (Open3D RemoveRadiusOutliers, per point)
// ---------------------------------------------------------------------------------------------
// With the statistical filter on, sor_mark_kernel has already moved the dead points to x = +inf.
// Level: the finest one whose 3x3 cells around the query hold more than `cap` points (dense regions: the
// walk stops after ~cap tests right around the query); sparse regions count at the coarsest level whose cells
// are at most a quarter of the radius.
#ifndef SD_RADIUS_GATHER
#define SD_RADIUS_GATHER 1   // 1: queries that need the candidate walk are gathered into full warps (0: every lane walks its own query)
#endif
// Most queries of a road cloud are decided by the whole-cell shortcut (their ball holds a few dense level-1 cells), the
// rest walk candidates.  Lane by lane that left the walk loop -- 60 % of the kernel's instructions -- running with 5.6 of
// 32 lanes.  So a warp works in two phases: it claims chunks of 32 queries and runs the shortcut on them, pushing the
// queries that still need a walk into a small per-warp queue (claim order is kept: neighbours stay neighbours), and as soon
// as 32 of them wait -- or the job's queries are used up -- it walks 32 at a time, one per lane.
#ifndef SD_RADIUS_MINB
#define SD_RADIUS_MINB 8     // resident CTAs per SM (64 registers, 38 bytes of spill; the kernel needs next to no shared memory).  Measured with
                             // the k-NN kernel at 7 CTAs: 5 -> 3 070, 6 -> 3 010, 8 -> 3 107, 10 -> 3 087 frames/s
#endif
__global__ void __launch_bounds__(kKnnThreads, SD_RADIUS_MINB)
radius_kernel(const KnnJob* __restrict__ jobs) {
    __shared__ int s_wq[kKnnThreads / 32][64];
    const KnnJob J = jobs[blockIdx.y];
#if SD_KNN_GRID_SMEM
    __shared__ GridRt s_grid;                                    // CTA-uniform geometry: an LDS where it is used instead of ~20 registers
    if (threadIdx.x == 0) s_grid = load_grid(J.gs);
    __syncthreads();
    const GridRt& g = s_grid;
#else
    const GridRt g = load_grid(J.gs);
#endif
    const double r = J.radius, r2 = r * r;
    const float r2_in = __double2float_rd(r2 * (1.0 - 3e-6)), r2_out = __double2float_ru(r2 * (1.0 + 3e-6));
    const int cap = J.count_cap;
    const float finf = __int_as_float(0x7f800000);
    const float r2_sure = __double2float_rd(r2 * (1.0 - 1e-4));
    const int ga2 = J.gs->a2;
    const int lane = lane_id();
    int* const wq = s_wq[warp_id()];
    int Lmax = 0;
#pragma unroll
    for (int l = 1; l < kLevels; ++l) if (g.cell * (double)(1 << (2 * l)) <= 0.25 * r) Lmax = l;   // measured: ~8 rows of small cells beat 3 rows of big ones
    int qhead = 0, qcount = 0;                                   // ring of waiting walkers (warp-uniform)
    bool exhausted = false;
    while (true) {
      // ---- phase A: claim chunks and decide what the shortcut can decide
      while (!exhausted && qcount < (SD_RADIUS_GATHER ? 32 : 1)) {
        int cbase = 0;
        if (lane == 0) cbase = atomicAdd(&J.gs->work, 32);
        cbase = __shfl_sync(SD_FULL, cbase, 0);
        if (cbase >= g.n) { exhausted = true; break; }
        const int i = cbase + lane;
        bool walk = false;
        if (i < g.n) {
            const float4 qp = __ldg(J.sp + i);
            const float qx = qp.x, qy = qp.y, qz = qp.z;
            if (qx != finf) {                                    // else: removed by the statistical filter (count already 0)
                walk = true;
                if (cap >= 0) {
                    // whole level-1 cells inside the ball: their alive points count without being looked at
                    const double q0 = (double)pick_axis(g.a0, qx, qy, qz), q1 = (double)pick_axis(g.a1, qx, qy, qz);
                    const int c0 = cell_coord(q0, g.o0, g.inv_cell, g.d0[0]), c1 = cell_coord(q1, g.o1, g.inv_cell, g.d1[0]);
                    const float cell1 = (float)(g.cell * 4.0), qa = pick_axis(ga2, qx, qy, qz);
                    const float f0 = (float)(q0 - g.o0), f1 = (float)(q1 - g.o1);        // query relative to the grid origin
                    const int k0 = c0 >> 2, k1 = c1 >> 2;
                    int sure = 0;
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int e0 = k0 + dx, e1 = k1 + dy;
                            if (e0 < 0 || e0 >= g.d0[1] || e1 < 0 || e1 >= g.d1[1]) continue;
                            const uint4 bx = __ldg(J.cell_box + e1 * g.d0[1] + e0);
                            const int na = (int)bx.x;
                            if (na <= 0) continue;
                            const float blo = key2f(bx.y), bhi = key2f(bx.z);
                            const float lo0 = (float)e0 * cell1, lo1 = (float)e1 * cell1;
                            const float far0 = fmaxf(fabsf(f0 - lo0), fabsf(lo0 + cell1 - f0)) + 1e-4f;
                            const float far1 = fmaxf(fabsf(f1 - lo1), fabsf(lo1 + cell1 - f1)) + 1e-4f;
                            const float far2 = fmaxf(fabsf(qa - blo), fabsf(bhi - qa)) + 1e-4f;
                            if ((far0 * far0 + far1 * far1) + far2 * far2 <= r2_sure) sure += na;
                        }
                    }
                    if (sure > cap) { J.cnt[__float_as_int(qp.w)] = cap + 1; walk = false; }
                }
            }
        }
        const unsigned wm = __ballot_sync(SD_FULL, walk);
        if (walk) wq[(qhead + qcount + __popc(wm & ((1u << lane) - 1u))) & 63] = i;
        qcount += __popc(wm);
        __syncwarp();
      }
      if (qcount == 0) break;                                    // nothing waits and nothing is left to claim
      // ---- phase B: up to 32 waiting queries, one per lane
      const int take = min(qcount, 32);
      if (lane < take) {
        const int i = wq[(qhead + lane) & 63];
        const float4 qp = __ldg(J.sp + i);
        const float qx = qp.x, qy = qp.y, qz = qp.z;
        const double q0 = (double)pick_axis(g.a0, qx, qy, qz), q1 = (double)pick_axis(g.a1, qx, qy, qz);
        const int c0 = cell_coord(q0, g.o0, g.inv_cell, g.d0[0]), c1 = cell_coord(q1, g.o1, g.inv_cell, g.d1[0]);
        int L = Lmax;
        if (cap >= 0) {
            int s3[3], e3[3];
#pragma unroll
            for (int l = kLevels - 2; l >= 0; --l) if (l < Lmax) { if (block3_runs(J, g, l, c0, c1, s3, e3) > cap) L = l; }
        }
        const LevelRt lv = level_of(J, g, L);
        const int k0 = c0 >> lv.shift, k1 = c1 >> lv.shift;
        const int rows = (int)ceil(r * lv.inv_cell) + 1;
        int count = 0;
        bool done = false;
        for (int t = 0; t <= 2 * rows && !done; ++t) {
            const int dr = (t == 0) ? 0 : ((t & 1) ? (t + 1) / 2 : -(t / 2));
            const int row = k1 + dr;
            if (row < 0 || row >= lv.d1) continue;
            double gap = 0.0;                                // distance from q to the row's slab along a1
            if (dr > 0) gap = (g.o1 + (double)row * lv.cell) - q1;
            else if (dr < 0) gap = q1 - (g.o1 + (double)(row + 1) * lv.cell);
            gap -= g.slack;
            if (gap > r) continue;
            if (gap < 0.0) gap = 0.0;
            const double half = (double)sqrtf(__double2float_ru(r2 - gap * gap)) * (1.0 + 2e-7) + g.slack;
            const int ca = cell_coord(q0 - half, g.o0, g.inv_cell, g.d0[0]) >> lv.shift,
                      cb = cell_coord(q0 + half, g.o0, g.inv_cell, g.d0[0]) >> lv.shift;
            const int rb = row * lv.d0;
            const int s = __ldg(lv.cs + rb + ca), e = __ldg(lv.cs + rb + cb + 1);
            if (s >= e) continue;
            // walk outwards from the query's own position (its own index in its row at level 0, the start of
            // the cell straight above / below it otherwise), four candidates per side and step: near
            // candidates first, so dense queries stop after ~cap tests, and eight independent loads are in flight
            int mid = (dr == 0 && L == 0) ? i : __ldg(lv.cs + rb + k0);
            mid = min(max(mid, s), e);
            int jl = mid - 1, jr = mid;
            while (!done && (jl >= s || jr < e)) {
                float4 c[8];
                unsigned okm = 0xffu;
                if (jr + 4 <= e && jl - 3 >= s) {                            // both sides full: no bound checks
                    const float4* __restrict__ pr = J.sp + jr;
                    const float4* __restrict__ pl = J.sp + (jl - 3);
#pragma unroll
                    for (int u = 0; u < 4; ++u) { c[u] = __ldg(pr + u); c[4 + u] = __ldg(pl + u); }
                } else {
                    okm = 0u;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int j = (u < 4) ? jr + u : jl - (u - 4);
                        const bool ok = (u < 4) ? (j < e) : (j >= s);
                        okm |= ok ? (1u << u) : 0u;
                        c[u] = __ldg(J.sp + (ok ? j : i));
                    }
                }
                unsigned inm = 0u, amb = 0u;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float d2f = key_of(c[u], qx, qy, qz);             // fp32 pre-test, relative error < 1.5e-6
                    inm |= (d2f < r2_in) ? (1u << u) : 0u;
                    amb |= (d2f >= r2_in && d2f <= r2_out) ? (1u << u) : 0u;
                }
                count += __popc(inm & okm);
                amb &= okm;
                if (amb) {                                                  // inside the error band (rare): decide in fp64
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if ((amb >> u) & 1u) count += (dist2_of(c[u], qx, qy, qz) <= r2) ? 1 : 0;
                }
                jr += 4; jl -= 4;
                done = (cap >= 0 && count > cap);
            }
        }
        J.cnt[__float_as_int(qp.w)] = (cap >= 0 && count > cap) ? cap + 1 : count;
      }
      __syncwarp();
      qhead = (qhead + take) & 63; qcount -= take;
    }
    if (lane == 0) {
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

template <int KS>
static int launch_knn_t(const sd::KnnJob* d_jobs, dim3 grid, cudaStream_t st) {
    using namespace sd;
    constexpr size_t smem = KnnCfg<(KS > 1 ? KS - 1 : 1)>::smem_bytes;
    static bool configured = false;       // per instantiation; the attribute is per function
    if (!configured) {
        SD_CUDA_TRY(cudaFuncSetAttribute(knn_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // shared-memory carve-out: just what the resident CTAs need, the rest of the 256 KB stays L1 (the sweeps wait on
        // scattered reads of the sorted copies); SD_KNN_CARVEOUT (percent) overrides for experiments
        const char* cv = getenv("SD_KNN_CARVEOUT");
        const int blocks = (KS <= 11 ? SD_KNN_MINB : 1);
        int pct = cv ? atoi(cv) : (int)(((smem + 1024 + 3584) * blocks * 100 + 228 * 1024 - 1) / (228 * 1024)) + 1;
        if (pct > 100) pct = 100;
        if (pct >= 0) SD_CUDA_TRY(cudaFuncSetAttribute(knn_kernel<KS>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        configured = true;
    }
    knn_kernel<KS><<<grid, kKnnThreads, smem, st>>>(d_jobs);
#ifndef SD_KNN_HEAVY_WAVES
#define SD_KNN_HEAVY_WAVES 6
#endif
    knn_heavy_kernel<KS><<<dim3(max(1u, min(grid.x, (148u * SD_KNN_HEAVY_WAVES) / grid.y)), grid.y), kHeavyWarps * 32, 0, st>>>(d_jobs);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

int sd_launch_knn(const sd::KnnJob* d_jobs, int njobs, int cap, int k, cudaStream_t st) {
    using namespace sd;
    if (njobs <= 0) return SD_OK;
    if (k < 1 || k > kMaxKnnK) return SD_ERR_INVALID;
    dim3 grid = grid_for(cap, kKnnThreads, 1, njobs, SD_KNN_WAVES);   // <= kKnnMaxBlocks CTAs per job
    if (k <= 3) return launch_knn_t<4>(d_jobs, grid, st);
    if (k <= 7) return launch_knn_t<8>(d_jobs, grid, st);
    if (k <= 10) return launch_knn_t<11>(d_jobs, grid, st);
    if (k <= 16) return launch_knn_t<17>(d_jobs, grid, st);
    if (k <= 20) return launch_knn_t<21>(d_jobs, grid, st);
    if (k <= 32) return launch_knn_t<33>(d_jobs, grid, st);
    return launch_knn_t<65>(d_jobs, grid, st);
}


namespace sd {
__global__ void cell_box_init_kernel(uint4* box, size_t count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        box[i] = make_uint4(0u, 0xffffffffu, 0u, 0u);
}
}  // namespace sd

namespace sd {
__global__ void ybox_init_kernel(uint2* box, size_t count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        box[i] = make_uint2(0xffffffffu, 0u);
}
}  // namespace sd

int sd_launch_ybox_init(uint2* d_box, size_t count, cudaStream_t st) {
    sd::ybox_init_kernel<<<148 * 4, 256, 0, st>>>(d_box, count);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

int sd_launch_cell_box_init(uint4* d_box, size_t count, cudaStream_t st) {
    sd::cell_box_init_kernel<<<148 * 4, 256, 0, st>>>(d_box, count);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

int sd_launch_radius(const sd::KnnJob* d_jobs, int njobs, int cap, cudaStream_t st) {
    using namespace sd;
    if (njobs <= 0) return SD_OK;
    // statistical filter applied to the sorted copies (if any) + per-cell statistics of what is left
    sor_mark_kernel<<<grid_for(3 * cap, kGridThreads, 4, njobs, 8), kGridThreads, 0, st>>>(d_jobs);
#ifndef SD_RADIUS_WAVES
#define SD_RADIUS_WAVES SD_RADIUS_MINB   // the resident set, like the k-NN kernel: surplus CTAs of a claim-until-empty kernel only hold slots
                               // (measured: 16 -> 5 waves, pipelined step +1.8 %)
#endif
    radius_kernel<<<grid_for(cap, kKnnThreads, 1, njobs, SD_RADIUS_WAVES), kKnnThreads, 0, st>>>(d_jobs);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
