// Internal job descriptors, workspace layout and launcher prototypes.  Every cloud kernel is
// "batched over jobs": blockIdx.y selects a job descriptor (one cloud of one frame), all sizes are
// read from device memory, so a whole batch of frames runs without any host round trip.
#pragma once
#include "sd_common.cuh"

namespace sd {

constexpr int kSelBits0 = 11, kSelBits1 = 11, kSelBits2 = 10;   // radix-select digit widths (32 bits)
constexpr int kSelBins = 2048;
constexpr int kCompactThreads = 256;
#ifndef SD_COMPACT_ITEMS
#define SD_COMPACT_ITEMS 8
#endif
#ifndef SD_COMPACT_MINB
#define SD_COMPACT_MINB 4
#endif
constexpr int kCompactItems = SD_COMPACT_ITEMS;
constexpr int kCompactTile = kCompactThreads * kCompactItems;   // 2048 points per tile
constexpr int kScanTile = 4096;                                 // cell-count scan tile
#ifndef SD_PLANE_BLOCKS
#define SD_PLANE_BLOCKS 64
#endif
constexpr int kPlaneBlocks = SD_PLANE_BLOCKS;                   // partial-sum blocks per plane job
constexpr int kPlaneSums = 9;
constexpr int kMaxKnnK = 64;
#ifndef SD_KNN_LEVELS
#define SD_KNN_LEVELS 3
#endif
constexpr int kLevels = SD_KNN_LEVELS;                                      // grid resolutions of the neighbour search (cell edge x4 per level)

// ---- device-side per-frame scalars --------------------------------------------------------------
struct FrameState {
    int32_t n[SD_NUM_COUNTS];   // stage counts; also the n_in/n_out cells the jobs point to
    int32_t n_sor_alive;
    int32_t slab_count;
    uint32_t slab_keys[2];      // ordered keys of min / max x in the slab
    uint32_t status;
    float med[5], mad[5];       // road y, road x, fence y, left x, right x
    float fence_mean;
    float pad_;
    double coeff[3][3];         // road / left / right plane: C0, C1, C2 (regression form)
    double sor_stats[3];        // mean, std, threshold
    int32_t ransac_best[3];
    int32_t road_first_alive;   // lowest row of the road cloud that survives MAD y (shift of the plane moments)
};

// ---- radix select ----------------------------------------------------------------------------------
struct SelState {
    uint32_t prefix[2];         // key bits resolved so far for the lower / upper middle rank
    uint32_t rank[2];           // remaining rank inside the prefix bucket
    uint32_t nan_count;
    uint32_t ticket;
    uint32_t alive;             // marking jobs: rows that passed the mark predicate in pass 0 (self-cleaned after pass 2)
    uint32_t first_inv;         // marking jobs: 0xffffffff - (lowest alive row), 0 = none (atomicMax; self-cleaned)
    uint32_t hist[2][kSelBins];
};
// A MAD filter that is not materialised by a compaction: the NEXT kernel over the cloud evaluates it on the fly, writes one
// alive byte per row and counts the survivors (pcl.py:63-67: fl32(fl32(0.6745f * |c - med|) / mad) < thr).
struct MadMark {
    const float* col;           // nullptr: no marking
    const float* med; const float* mad;
    float thr; int32_t pad_;
};
struct SelJob {
    const float* col;
    const int32_t* n;           // number of keys that take part (alive rows).  Marking jobs: written by pass 0 (= n_mark_out)
    const int32_t* n_loop;      // physical rows of the column when rows are masked (nullptr: *n)
    uint8_t* flag;              // alive byte per row (nullptr: every row takes part); marking jobs write it in pass 0
    MadMark mark;               // pass 0 evaluates this filter first, rows that fail never enter the histogram
    int32_t* n_mark_out;        // marking jobs: where the survivor count goes
    int32_t* first_alive_out;   // marking jobs: lowest alive row (a finite representative of the surviving cloud), 0 if none
    const float* center;        // nullptr: keys are col[i]; else keys are |col[i] - *center|
    SelState* st;
    float* out;                 // the median
    uint32_t* status;           // optional status word
    uint32_t zero_bit;          // OR'ed into *status when n > 0 and !(median > 0)  (MAD == 0 / NaN)
    uint32_t pad_;
};

// ---- stable compaction -------------------------------------------------------------------------------
struct PredDev {
    int32_t kind, axis, ia, use_f32;
    float fa, f0, f1, pad_;
    double da, d0, d1, d2;
    const void* aux;            // SOR: const double* avg ; ROR: const int32_t* counts
    const void* aux2;           // SOR+ROR fused: const int32_t* counts
    const float* p_f0;          // device-resident overrides (fused path): median / mean
    const float* p_f1;          // mad
    const double* p_d;          // plane C0,C1,C2 ; SOR: threshold at p_d[0]
};
constexpr int SD_PRED_SORROR = 100;   // internal: 0 < avg < thr (if p_d) && cnt > ia (if aux2)

struct CompactJob {
    const float* x; const float* y; const float* z; const int32_t* src; const int32_t* n_in;
    float* ox; float* oy; float* oz; int32_t* osrc; int32_t* n_out;
    PredDev pred;
    const uint8_t* flag;        // optional alive byte per input row (rows marked dead by earlier, unmaterialised filters)
    PredDev pred2;              // optional second filter applied to the survivors of `pred` (has_pred2): two reference
    int32_t has_pred2;          //   calls in one pass, e.g. remove_noise_by_mad + threshold_complete (semantic_depth.py:279,283)
    int32_t pad2_;
    int32_t* n_mid;             // has_pred2: survivors of flag && pred (the first call's count stays observable)
    uint8_t* mid_alive;         // has_pred2, optional: one byte per input row, 1 = survived flag && pred (keeps that stage inspectable)
    unsigned long long* status; // look-back words, max_tiles long
    ScanCtl* ctl;
    uint32_t* frame_status;     // optional: OR `empty_bit` when the output is empty
    uint32_t empty_bit;
    int32_t max_tiles;
};

// ---- plane fit -----------------------------------------------------------------------------------------
struct PlaneJob {
    const float* x; const float* y; const float* z; const int32_t* n;
    const int32_t* n_loop;      // physical rows when rows are masked (nullptr: *n)
    const uint8_t* flag;        // alive byte per row from earlier unmaterialised filters (nullptr: all rows alive)
    uint8_t* flag_out;          // with `mark`: alive byte per row after this filter (may alias `flag`: one thread per row)
    MadMark mark;               // evaluate this MAD filter first (remove_noise_by_mad before the plane fit)
    int32_t* n_mark_out;        // mark: survivor count
    const int32_t* shift_row;   // row whose coordinates shift the moments (nullptr: row 0); must be a finite row of the cloud
    int32_t axis;               // regressed coordinate (pcl.py axis argument)
    int32_t use_inliers;        // RANSAC refit: only points with |res(hyp)| < thr contribute
    const double* hyp;          // C0,C1,C2 of the best hypothesis (device)
    double thr;
    double* partials;           // [kPlaneBlocks][kPlaneSums]
    uint32_t* ticket;
    double* coeff;              // out: C0, C1, C2
    uint32_t* status;           // optional
    uint32_t empty_bit;         // OR'ed when n == 0
    uint32_t pad_;
};

// ---- NumPy pairwise fp32 mean ----------------------------------------------------------------------------
struct MeanJob {
    const float* col; const int32_t* n; float* out;
    float* leaf;                // [cap/64 + 1] scratch: pairwise leaf sums, indexed by leaf_start / 64
};

// ---- slab min/max ---------------------------------------------------------------------------------------
struct SlabJob {
    const float* x; const float* z; const int32_t* n;
    double lo, hi; float lo32, hi32; int32_t use_f32; int32_t pad_;
    uint32_t* keys;             // [2] min / max ordered keys (reset to 0xffffffff / 0 by the consumer)
    int32_t* count;
};

// ---- uniform grid / kNN ------------------------------------------------------------------------------------
struct GridState {
    uint32_t bbox[6];           // ordered keys: min x,y,z then max x,y,z  (reset by the last bbox block)
    uint32_t ticket;
    int32_t ncells;             // over all levels
    int32_t a0, a1, a2;         // a0 = fast grid axis, a1 = slow grid axis, a2 = collapsed axis
    int32_t d0, d1;             // level-0 cells along a0 / a1
    int32_t ld0[kLevels], ld1[kLevels], loff[kLevels];   // per level: cells along a0 / a1, offset into the cell arrays
    int32_t n;                  // snapshot of the cloud size
    double o0, o1;              // grid origin along a0 / a1
    double cell, inv_cell;
    double ext2;                // extent of the collapsed axis (for the 2D-inside shortcut)
    unsigned long long acc[3][2];  // exact 128-bit fixed-point sums of avg, avg^2 (lo, hi) and the count of avg > 0
    int32_t work;                  // dynamic work counter of the search kernels (next unclaimed sorted index)
    int32_t qn, qhead;             // queue of the heavy k-NN queries: entries written / entries claimed
    int32_t pad_;
};
struct KnnJob {
    const float* x; const float* y; const float* z; const int32_t* n;
    GridState* gs;
    int32_t* cell_count;        // [cell_cap + 1] all levels concatenated, all zero between launches
    int32_t* cell_start;        // [cell_cap + 1] absolute positions into the sorted copies
    int32_t* cell_of;           // [cap] level-0 cell of each point (input order)
    float4* sp;                 // [kLevels * cap] cell-sorted copies (x, y, z, bits(original index)), level L at [L*n, (L+1)*n)
    uint2* ybox;                // [cell_cap / 8] per level-1 / level-2 cell: ordered keys of min / max along the collapsed axis, all points
    int32_t* queue; float* queue_band;   // [cap] heavy k-NN queries: level-0 sorted index, bound of the k-th squared distance
    uint4* cell_box;            // [cell_cap / 8] per level-1 cell: (alive count, ordered keys of min / max along the collapsed axis, -)
    double* avg;                // [cap] mean kNN distance, by ORIGINAL index
    int32_t* cnt;               // [cap] radius counts, by original index
    unsigned long long* scan_status; ScanCtl* scan_ctl;
    double* stats;              // mean, std, thr
    int32_t* n_alive;
    int32_t cell_cap;
    int32_t k;
    double std_ratio;
    double radius;
    int32_t nb_points;
    int32_t use_sor;            // 0: every point is alive for the radius count
    int32_t count_cap;          // saturate radius counts at count_cap + 1 (-1: exact counts)
    int32_t pad_;
    double cell_scale;          // cell edge = cell_scale * sqrt(area / n)
};
constexpr int kKnnMaxBlocks = 148 * 16;

// ---- RANSAC ----------------------------------------------------------------------------------------------
struct RansacJob {
    const float* x; const float* y; const float* z; const int32_t* n;
    const int32_t* triplets;    // [n_hyp][3]
    double* hyp_coeff;          // [n_hyp][4]: C0, C1, C2, valid
    int32_t* hyp_counts;        // [n_hyp]
    int32_t* best;              // out
    double* best_coeff;         // out: C0,C1,C2
    int32_t axis; int32_t n_hyp;
    double thr;
};

// ---- finalize ----------------------------------------------------------------------------------------------
struct FinalJob {
    FrameState* fs;
    SdFrameResult* out;
};

// ---- exact, order-independent accumulation of the cloud statistics ------------------------------------
// The per-point means are bit-exact, but their summation order would depend on the (atomic) order of
// points inside a cell.  Summing 2^-70 fixed-point images of the values in 128-bit integers is exact
// and associative, so the cloud mean / std are deterministic run to run (and closer to the real sum
// than any fp64 summation order).
// RGBA bytes (little-endian packed) of the pasted segmentation layers after scipy's bytescale, for a mask that covers
// part of / all of the frame (sd_overlay.cu)
struct OverlayLayers { uint32_t road_partial, road_full, fence_partial, fence_full; };

struct U128 { unsigned long long lo, hi; };
__device__ __forceinline__ U128 to_fixed70(double v) {      // floor(v * 2^70) for finite v > 0, else 0
    U128 r{0ull, 0ull};
    if (!(v > 0.0)) return r;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const int e = (int)((bits >> 52) & 0x7ffull);
    if (e == 0 || e == 0x7ff) return r;
    const unsigned long long m = (bits & 0xfffffffffffffull) | (1ull << 52);
    int sh = e - 1075 + 70;                                  // v = m * 2^(e-1075)
    if (sh > 74) sh = 74;                                    // saturate (|v| >= 2^57 never happens for metres)
    if (sh >= 64) { r.hi = m << (sh - 64); }
    else if (sh > 0) { r.lo = m << sh; r.hi = m >> (64 - sh); }
    else if (sh == 0) { r.lo = m; }
    else if (sh > -64) { r.lo = m >> (-sh); }
    return r;
}
__device__ __forceinline__ U128 add128(U128 a, U128 b) {
    U128 r; r.lo = a.lo + b.lo; r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull); return r;
}
__device__ __forceinline__ U128 warp_sum128(U128 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        U128 t; t.lo = __shfl_xor_sync(SD_FULL, v.lo, o); t.hi = __shfl_xor_sync(SD_FULL, v.hi, o);
        v = add128(v, t);
    }
    return v;
}
__device__ __forceinline__ void atomic_add128(unsigned long long* acc, U128 v) {
    const unsigned long long old = atomicAdd(&acc[0], v.lo);
    const unsigned long long carry = (old + v.lo < old) ? 1ull : 0ull;
    if (v.hi | carry) atomicAdd(&acc[1], v.hi + carry);
}
__device__ __forceinline__ double fixed70_to_double(unsigned long long lo, unsigned long long hi) {
    return ((double)hi * 18446744073709551616.0 + (double)lo) * 8.470329472543003e-22;   // 2^-70
}


}  // namespace sd

// ---- host-side workspace ------------------------------------------------------------------------------------
struct SdCloudBuf { float* x; float* y; float* z; int32_t* src; };

struct SdWorkspace {
    char* base; size_t bytes;
    int max_frames, height, width, cap, max_tiles, cell_cap, max_hyp;
    // per-frame arrays (index = frame)
    sd::FrameState* fs;                  // [F]
    SdCloudBuf road[2], fence[2], left[2], right[2];   // each buffer: F * cap elements per array
    // select / scan / plane scratch per (frame, chain) ; chain: 0 road, 1 fence, 2 left, 3 right
    sd::SelState* sel;                   // [F][4]
    unsigned long long* cstatus;         // [F][4][max_tiles]
    sd::ScanCtl* cctl;                   // [F][4]
    float* mean_leaf;                    // [F][cap/64 + 1]
    double* partials;                    // [F][4][kPlaneBlocks*kPlaneSums]
    uint32_t* ptick;                     // [F][4]
    // pixel pass
    uint8_t* pflags;                     // [F][H*W] per-pixel label / keep flags
    uint8_t* cflags;                     // [F][4][cap] alive bytes of unmaterialised MAD filters: road, left, right, fence chain
    int32_t* ptcounts;                   // [F][pix_tiles][4]
    int32_t* ptoffs;                     // [F][pix_tiles][2]
    int pix_tiles;
    double* lmask; double* rmask;        // [width] blend ramps (uploaded at create)
    // grid / knn per frame
    sd::GridState* gs;                   // [F]
    int32_t* cell_count; int32_t* cell_start; int32_t* cell_of;
    float4* sp; uint4* cell_box; uint2* ybox; int32_t* queue; float* queue_band;
    double* avg; int32_t* cnt;
    unsigned long long* gstatus; sd::ScanCtl* gctl; int grid_tiles;
    // ransac per (frame, chain 0..2)
    double* hyp_coeff; int32_t* hyp_counts; double* best_coeff;
    // job descriptor arenas (device) + small host-visible scratch
    char* jobs; size_t jobs_bytes; size_t jobs_used;
    char* scratch; size_t scratch_bytes;     // device scratch for per-call ops (results, descriptors)
    void* h_pinned; size_t h_pinned_bytes;   // pinned host mirror for small readbacks
    // prebuilt descriptor arrays for the fused path (device pointers)
    bool fused_ready;
    int fused_batch;
    SdParams fused_params;
};

// launchers (each returns SD_OK / SD_ERR_*); jobs are DEVICE pointers
int sd_launch_select_median(const sd::SelJob* d_jobs, int njobs, int cap, cudaStream_t st);
int sd_launch_compact(const sd::CompactJob* d_jobs, int njobs, int cap, cudaStream_t st);
int sd_launch_plane(const sd::PlaneJob* d_jobs, int njobs, int cap, cudaStream_t st);
int sd_launch_mean(const sd::MeanJob* d_jobs, int njobs, int cap, cudaStream_t st);
int sd_launch_slab(const sd::SlabJob* d_jobs, int njobs, int cap, cudaStream_t st);
int sd_launch_grid_build(const sd::KnnJob* d_jobs, int njobs, int cap, cudaStream_t st);
int sd_launch_knn(const sd::KnnJob* d_jobs, int njobs, int cap, int k, cudaStream_t st);
int sd_launch_radius(const sd::KnnJob* d_jobs, int njobs, int cap, cudaStream_t st);
int sd_launch_cell_box_init(uint4* d_box, size_t count, cudaStream_t st);
int sd_launch_ybox_init(uint2* d_box, size_t count, cudaStream_t st);
int sd_launch_ply_rows(const float* d_x, const float* d_y, const float* d_z, const uint8_t* d_rgb, int n, char* d_out,
                       unsigned long long capacity, uint32_t* d_tile_scratch, unsigned long long* d_total, cudaStream_t st);
int sd_launch_ply_rows_f64(const double* d_x, const double* d_y, const double* d_z, const uint8_t* d_rgb, int n, char* d_out,
                           unsigned long long capacity, uint32_t* d_tile_scratch, unsigned long long* d_total, cudaStream_t st);
int sd_launch_resize_cubic_u8(const uint8_t* d_src, int batch, int src_h, int src_w, int channels, uint8_t* d_dst, int dst_h, int dst_w,
                              cudaStream_t st);
int sd_launch_overlay(const uint8_t* d_frame, const uint8_t* d_labels, int batch, int hw, const sd::OverlayLayers& layers,
                      int* d_counts, uint8_t* d_out, cudaStream_t st);
int sd_launch_banner(uint8_t* d_frames, int batch, int h, int w, const int32_t* d_rects, int nrects,
                     const uint32_t* d_bits, int nbitmaps, int cell_h, int words, const int32_t* d_places, int nplaces,
                     cudaStream_t st);
int sd_launch_ransac(const sd::RansacJob* d_jobs, int njobs, int cap, int n_hyp, cudaStream_t st);
int sd_launch_finalize(const sd::FinalJob* d_jobs, int njobs, const SdParams* params, cudaStream_t st);
int sd_launch_pixel(const float* d_logits, const float* d_disp, const double* d_lmask, const double* d_rmask,
                    int batch, int height, int width, SdCamera cam, double prob_thr, float road_z, int flags,
                    SdCloudBuf road, SdCloudBuf fence, int cap_stride,
                    int32_t* d_cnt_road_gather, int32_t* d_cnt_road_z, int32_t* d_cnt_fence, int cnt_stride,
                    uint8_t* d_labels, float* d_points, float* d_disp_pp,
                    uint8_t* d_flags, int32_t* d_tcounts, int32_t* d_toffs, int pix_tiles, cudaStream_t st,
                    const float* d_scores = nullptr, const float* d_upw = nullptr, const float* d_upb = nullptr,
                    float* d_logits_out = nullptr, int label_mode = 0);
