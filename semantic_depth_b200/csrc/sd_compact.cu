// Stable predicate compaction of SoA clouds + the small column reductions around it.
//
//   sd_launch_compact : every cloud -> cloud filter of pcl.py (remove_from_to :30-43,
//                       remove_noise_by_mad :46-73, remove_noise_by_fitting_plane's residual test
//                       :130-131/163-164/196-197, threshold_complete :240-250, extract_pcls :253-268,
//                       the slab of get_end_points_of_road :283) and the Open3D select_down_sample of
//                       semantic_depth.py:236,241.  Output order == input order (NumPy indexing).
//   sd_launch_mean    : np.mean of an fp32 column in NumPy's pairwise order (pcl.py:258).
//   sd_launch_slab    : min / max x inside the depth slab (pcl.py:283,307-308).
//
// Compaction is a persistent single-pass scan: CTAs draw 4096-point tiles from a per-job ticket,
// evaluate the predicate on 128-bit loads, block-scan the keep flags and chain tiles by decoupled
// look-back; no atomics decide positions, so indices are bit-exact and deterministic.
#include "sd_internal.cuh"

namespace sd {

// ---------------------------------------------------------------------------------------------
// predicates (one IEEE rounding per operator; -fmad=false)
// ---------------------------------------------------------------------------------------------
struct PredRt {   // predicate with device-resident parameters resolved
    int kind, axis, ia, use_f32;
    float fa, f0, f1;
    double da, d0, d1, d2;
    const double* avg; const int32_t* cnt;
    bool has_thr, has_cnt;
};

__device__ __forceinline__ PredRt resolve_pred(const PredDev& p) {
    PredRt r;
    r.kind = p.kind; r.axis = p.axis; r.ia = p.ia; r.use_f32 = p.use_f32;
    r.fa = p.fa; r.f0 = p.p_f0 ? *p.p_f0 : p.f0; r.f1 = p.p_f1 ? *p.p_f1 : p.f1;
    r.da = p.da; r.d0 = p.d0; r.d1 = p.d1; r.d2 = p.d2;
    r.avg = nullptr; r.cnt = nullptr; r.has_thr = false; r.has_cnt = false;
    if (p.kind == SD_PRED_PLANE && p.p_d) { r.d0 = p.p_d[0]; r.d1 = p.p_d[1]; r.d2 = p.p_d[2]; }
    if (p.kind == SD_PRED_GT || p.kind == SD_PRED_LT) { if (p.p_f0) r.fa = *p.p_f0; }
    if (p.kind == SD_PRED_SOR) { r.avg = (const double*)p.aux; if (p.p_d) r.da = p.p_d[0]; }
    if (p.kind == SD_PRED_ROR) { r.cnt = (const int32_t*)p.aux; }
    if (p.kind == SD_PRED_SORROR) {
        r.avg = (const double*)p.aux; r.cnt = (const int32_t*)p.aux2;
        r.has_thr = (p.aux != nullptr); r.has_cnt = (p.aux2 != nullptr);
        if (p.p_d) r.da = p.p_d[0];
    }
    return r;
}

__device__ __forceinline__ bool eval_pred(const PredRt& p, float x, float y, float z, int i) {
    const float c = (p.axis == 0) ? x : (p.axis == 1 ? y : z);
    switch (p.kind) {
        case SD_PRED_LT: return c < p.fa;
        case SD_PRED_GT: return c > p.fa;
        case SD_PRED_ABS_LT: return fabsf(c) < p.fa;
        case SD_PRED_MAD: {
            float ad = fabsf(c - p.f0);              // abs(points1D - median)           pcl.py:79
            float pen = (0.6745f * ad) / p.f1;       // 0.6745 * abs_diffs / mad_axis    pcl.py:63
            return pen < p.fa;                       // NaN / inf compare false          pcl.py:67
        }
        case SD_PRED_PLANE: {
            float u, v;
            if (p.axis == 0) { u = y; v = z; } else if (p.axis == 1) { u = x; v = z; } else { u = x; v = y; }
            double a = ((p.d0 * (double)u + p.d1 * (double)v) - (double)c) + p.d2;   // pcl.py:130/163/196
            return fabs(a) < p.da;
        }
        case SD_PRED_SLAB:
            if (p.use_f32) return (z < p.f1) && (z > p.f0);
            return ((double)z < p.d1) && ((double)z > p.d0);
        case SD_PRED_SOR: { double a = p.avg[i]; return a > 0.0 && a < p.da; }
        case SD_PRED_ROR: return p.cnt[i] > p.ia;
        case SD_PRED_SORROR: {
            bool k = true;
            if (p.has_thr) { double a = p.avg[i]; k = a > 0.0 && a < p.da; }
            if (k && p.has_cnt) k = p.cnt[i] > p.ia;
            return k;
        }
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// compaction
// ---------------------------------------------------------------------------------------------
// One tile = kCompactTile consecutive points.  A thread loads kCompactItems consecutive points with
// 128-bit loads (x, y, z and the source index), evaluates the predicate, the block scans the keep counts,
// one warp chains the tile to its predecessors (decoupled look-back), the survivors are staged in shared
// memory in output order and leave through fully coalesced stores.
__global__ void __launch_bounds__(kCompactThreads, SD_COMPACT_MINB)
compact_kernel(const CompactJob* __restrict__ jobs) {
    __shared__ int s_scan[33];
    __shared__ int s_tile;
    __shared__ unsigned long long s_excl;
    __shared__ float s_x[kCompactTile], s_y[kCompactTile], s_z[kCompactTile];
    __shared__ int s_src[kCompactTile];

#ifndef SD_COMPACT_SMEM
#define SD_COMPACT_SMEM 1
#endif
    const int tid = threadIdx.x;
#if SD_COMPACT_SMEM
    // the job descriptor and the resolved predicates are CTA-uniform: in shared memory they cost an LDS where they are used
    // instead of ~70 registers per thread (16 pointers, two predicates with their fp64 parameters)
    __shared__ CompactJob s_job;
    __shared__ PredRt s_pred[2];
    if (tid == 0) {
        s_job = jobs[blockIdx.y];
        s_pred[0] = resolve_pred(s_job.pred);
        s_pred[1] = s_job.has_pred2 ? resolve_pred(s_job.pred2) : s_pred[0];
    }
    __syncthreads();
    const CompactJob& J = s_job;
    const PredRt& P = s_pred[0];
    const PredRt& P2 = s_pred[1];
    const bool two = J.has_pred2 != 0;
#else
    const CompactJob J = jobs[blockIdx.y];
    const PredRt P = resolve_pred(J.pred);
    const bool two = J.has_pred2 != 0;
    const PredRt P2 = two ? resolve_pred(J.pred2) : P;
#endif
    const int n = *J.n_in;
    const int ntiles = ceil_div(n, kCompactTile);
    const uint8_t* __restrict__ flag = J.flag;
    int mid_local = 0;
    const float* __restrict__ X = J.x; const float* __restrict__ Y = J.y; const float* __restrict__ Z = J.z;
    const bool need_x = (P.kind == SD_PRED_PLANE) || (J.ox != nullptr) || (P.axis == 0 && P.kind <= SD_PRED_GT) || (two && P2.axis == 0);
    const bool need_y = (P.kind == SD_PRED_PLANE) || (J.oy != nullptr) || (P.axis == 1 && P.kind <= SD_PRED_GT) || (two && P2.axis == 1);
    const bool need_z = (P.kind == SD_PRED_PLANE) || (J.oz != nullptr) || (P.axis == 2 && P.kind <= SD_PRED_GT) ||
                        (P.kind == SD_PRED_SLAB) || (two && P2.axis == 2);
    const bool need_s = (J.osrc != nullptr) && (J.src != nullptr);

    const bool vec_ok = ((((uintptr_t)X) | ((uintptr_t)Y) | ((uintptr_t)Z) | ((uintptr_t)J.src)) & 15) == 0;   // 128-bit loads need alignment
    while (true) {
        if (tid == 0) s_tile = (int)atomicAdd(&J.ctl->ticket, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= ntiles) break;
        const int base = tile * kCompactTile + tid * kCompactItems;
        float x[kCompactItems], y[kCompactItems], z[kCompactItems];
        int sidx[kCompactItems];
        unsigned keep = 0u;
        if (vec_ok && base + kCompactItems <= n) {
#pragma unroll
            for (int q = 0; q < kCompactItems / 4; ++q) {
                float4 a = need_x ? __ldg(reinterpret_cast<const float4*>(X + base) + q) : make_float4(0, 0, 0, 0);
                float4 b = need_y ? __ldg(reinterpret_cast<const float4*>(Y + base) + q) : make_float4(0, 0, 0, 0);
                float4 c = need_z ? __ldg(reinterpret_cast<const float4*>(Z + base) + q) : make_float4(0, 0, 0, 0);
                int4 d = need_s ? __ldg(reinterpret_cast<const int4*>(J.src + base) + q)
                                : make_int4(base + 4 * q, base + 4 * q + 1, base + 4 * q + 2, base + 4 * q + 3);
                x[4 * q] = a.x; x[4 * q + 1] = a.y; x[4 * q + 2] = a.z; x[4 * q + 3] = a.w;
                y[4 * q] = b.x; y[4 * q + 1] = b.y; y[4 * q + 2] = b.z; y[4 * q + 3] = b.w;
                z[4 * q] = c.x; z[4 * q + 1] = c.y; z[4 * q + 2] = c.z; z[4 * q + 3] = c.w;
                sidx[4 * q] = d.x; sidx[4 * q + 1] = d.y; sidx[4 * q + 2] = d.z; sidx[4 * q + 3] = d.w;
            }
#pragma unroll
            for (int k = 0; k < kCompactItems; ++k) keep |= eval_pred(P, x[k], y[k], z[k], base + k) ? (1u << k) : 0u;
            if (flag) {
                unsigned fm = 0u;
#pragma unroll
                for (int k = 0; k < kCompactItems; ++k) fm |= flag[base + k] ? (1u << k) : 0u;
                keep &= fm;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kCompactItems; ++k) {
                const int i = base + k;
                x[k] = y[k] = z[k] = 0.f; sidx[k] = i;
                if (i < n) {
                    if (need_x) x[k] = __ldg(X + i);
                    if (need_y) y[k] = __ldg(Y + i);
                    if (need_z) z[k] = __ldg(Z + i);
                    if (need_s) sidx[k] = __ldg(J.src + i);
                    keep |= ((!flag || flag[i]) && eval_pred(P, x[k], y[k], z[k], i)) ? (1u << k) : 0u;
                }
            }
        }
        if (two) {                                                  // the second reference call works on the first one's survivors
            mid_local += __popc(keep);
            if (J.mid_alive) {
#pragma unroll
                for (int k = 0; k < kCompactItems; ++k) if (base + k < n) J.mid_alive[base + k] = (uint8_t)((keep >> k) & 1u);
            }
#pragma unroll
            for (int k = 0; k < kCompactItems; ++k)
                if ((keep >> k) & 1u) { if (!eval_pred(P2, x[k], y[k], z[k], base + k)) keep &= ~(1u << k); }
        }
        int total;
        const int excl = block_excl_scan(__popc(keep), s_scan, &total);
        if (warp_id() == 0) {
            unsigned long long e = lookback_exclusive(J.status, tile, (unsigned long long)total);
            if (lane_id() == 0) s_excl = e;
        }
        // stage the survivors in output order (block_excl_scan ended with a barrier: s_x.. are free)
        int lpos = excl;
#pragma unroll
        for (int k = 0; k < kCompactItems; ++k) {
            if (keep & (1u << k)) { s_x[lpos] = x[k]; s_y[lpos] = y[k]; s_z[lpos] = z[k]; s_src[lpos] = sidx[k]; ++lpos; }
        }
        __syncthreads();
        const int out0 = (int)s_excl;
        if (tile == ntiles - 1 && tid == 0 && J.n_out) *J.n_out = out0 + total;
        for (int i = tid; i < total; i += kCompactThreads) {
            if (J.ox) J.ox[out0 + i] = s_x[i];
            if (J.oy) J.oy[out0 + i] = s_y[i];
            if (J.oz) J.oz[out0 + i] = s_z[i];
            if (J.osrc) J.osrc[out0 + i] = s_src[i];
        }
        __syncthreads();   // s_tile / s_excl / staging are rewritten by the next iteration
    }
    if (two) {                                                      // survivors of the first filter: exact integer sum
        mid_local = warp_sum(mid_local);
        if (lane_id() == 0 && mid_local) atomicAdd(&J.ctl->aux0, (unsigned)mid_local);
    }
    if (scan_finish(J.ctl, J.status, max(ntiles, 0), gridDim.x)) {
        if (tid == 0 && two) { if (J.n_mid) *J.n_mid = (int)__ldcg(&J.ctl->aux0); J.ctl->aux0 = 0u; }
        if (tid == 0) {
            int nout = (n > 0 && J.n_out) ? __ldcg(J.n_out) : 0;
            if (n <= 0 && J.n_out) *J.n_out = 0;
            if (J.frame_status && J.empty_bit && nout == 0) atomicOr(J.frame_status, J.empty_bit);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// np.mean(fp32 column): NumPy's pairwise summation tree (verified bit-for-bit):
//   n < 8    : r = 0; r += a[i] sequentially
//   n <= 128 : 8 strided accumulators, ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)), then the n%8 tail
//   n > 128  : n2 = n/2; n2 -= n2 % 8; sum(a[0:n2]) + sum(a[n2:n])
//   mean = fl32(sum / fl32(n))
// One CTA of 1024 threads per job: thread t owns the depth-10 subtree reached by following the bits
// of t from the root, sums it serially in exact tree order, then the upper ten levels are combined
// in the same order (32 lanes do levels 5..9, one lane levels 0..4).
// ---------------------------------------------------------------------------------------------
constexpr int kMeanDepth = 10;
constexpr int kMeanThreads = 1 << kMeanDepth;

// leaf of the pairwise tree that contains element e
__device__ __forceinline__ void find_leaf(int n, int e, int& off, int& len) {
    off = 0; len = n;
    while (len > 128) {
        int n2 = len / 2; n2 -= n2 % 8;
        if (e < off + n2) len = n2; else { off += n2; len -= n2; }
    }
}

// sum of the subtree (off, len) from the stored leaf sums, in NumPy's order (explicit stack, depth <= 24)
__device__ float subtree_from_leaves(const float* leaf, int off0, int len0) {
    struct Fr { int off, n, state; float left; };
    Fr st[26];
    int sp = 0;
    st[0] = {off0, len0, 0, 0.f};
    float ret = 0.f;
    while (sp >= 0) {
        Fr& f = st[sp];
        if (f.n <= 128) { ret = leaf[f.off >> 6]; --sp; continue; }
        int n2 = f.n / 2; n2 -= n2 % 8;
        if (f.state == 0) { f.state = 1; st[sp + 1] = {f.off, n2, 0, 0.f}; ++sp; }
        else if (f.state == 1) { f.left = ret; f.state = 2; st[sp + 1] = {f.off + n2, f.n - n2, 0, 0.f}; ++sp; }
        else { ret = f.left + ret; --sp; }
    }
    return ret;
}

// Combine `levels` levels of the tree above stored subtree sums.  (off, n) is the node; `path` its
// index among the nodes of its depth (bit string from the root); leaves that end early store their
// value at the slot of their left-most descendant.
__device__ float combine_levels(const float* vals, int n, int path, int levels, int stride_shift) {
    struct Fr { int n, q, d, state; float left; };
    Fr st[12];
    int sp = 0;
    st[0] = {n, 0, 0, 0, 0.f};
    float ret = 0.f;
    while (sp >= 0) {
        Fr& f = st[sp];
        if (f.d == levels || f.n <= 128) {
            ret = vals[(((path << levels) | (f.q << (levels - f.d)))) << stride_shift];
            --sp; continue;
        }
        int n2 = f.n / 2; n2 -= n2 % 8;
        if (f.state == 0) { f.state = 1; st[sp + 1] = {n2, f.q << 1, f.d + 1, 0, 0.f}; ++sp; }
        else if (f.state == 1) { f.left = ret; f.state = 2; st[sp + 1] = {f.n - n2, (f.q << 1) | 1, f.d + 1, 0, 0.f}; ++sp; }
        else { ret = f.left + ret; --sp; }
    }
    return ret;
}

// Leaf sums across the whole GPU, then one CTA of 1024 threads per job for the tree above them.
//   A. (mean_leaf_kernel, many CTAs per job) leaf sums: 8 lanes per leaf play NumPy's 8 strided accumulators
//      (coalesced loads), combined as ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)) by three xor-shuffles, then the n%8
//      tail; a leaf (>= 64 elements whenever n > 128) is owned by the 64-element slot its start falls in
//   B. (mean_kernel) thread t sums the leaf sums of the depth-10 subtree reached by the bits of t, in tree order
//   C. 32 lanes combine levels 5..9, one lane levels 0..4; mean = fl32(sum / fl32(n))
__global__ void __launch_bounds__(kMeanThreads)
mean_leaf_kernel(const MeanJob* __restrict__ jobs) {
    const MeanJob J = jobs[blockIdx.y];
    const int n = *J.n;
    const int t = threadIdx.x;
    const float* __restrict__ a = J.col;
    // the slot loop has the same trip count for every thread of a CTA
    const int grp = t >> 3, j = t & 7, ngroups = kMeanThreads >> 3;
    const int nslots = (n + 63) >> 6;
    if (n == 0 && t == 0 && blockIdx.x == 0) J.leaf[0] = 0.f;
    for (int s0 = blockIdx.x * ngroups; s0 < nslots; s0 += gridDim.x * ngroups) {
        const int slot = s0 + grp;
        int off = 0, len = 0;
        bool owner = false;
        if (slot < nslots) {
            find_leaf(n, min(slot * 64 + 63, n - 1), off, len);
            owner = (off >> 6) == slot;
        }
        float r = 0.f;
        if (owner) {
            if (len < 8) {
                if (j == 0) for (int i = 0; i < len; ++i) r = r + a[off + i];
            } else {
                r = a[off + j];
                const int body = len - (len % 8);
                for (int i = 8; i < body; i += 8) r = r + a[off + i + j];
            }
        }
        const float r1 = r + __shfl_xor_sync(SD_FULL, r, 1);
        const float r2 = r1 + __shfl_xor_sync(SD_FULL, r1, 2);
        const float r3 = r2 + __shfl_xor_sync(SD_FULL, r2, 4);
        if (owner && j == 0) {
            float res = (len < 8) ? r : r3;
            if (len >= 8) for (int i = len - (len % 8); i < len; ++i) res = res + a[off + i];
            J.leaf[slot] = res;
        }
    }
}

__global__ void __launch_bounds__(kMeanThreads)
mean_kernel(const MeanJob* __restrict__ jobs) {
    __shared__ float s_sub[kMeanThreads];
    __shared__ float s_mid[32];
    const MeanJob J = jobs[blockIdx.x];
    const int n = *J.n;
    const int t = threadIdx.x;
    // ---- B: descend kMeanDepth levels following the bits of t (MSB first)
    int off = 0, len = n; bool owner = true; int d = 0;
    for (; d < kMeanDepth; ++d) {
        if (len <= 128) break;
        int n2 = len / 2; n2 -= n2 % 8;
        if ((t >> (kMeanDepth - 1 - d)) & 1) { off += n2; len -= n2; } else { len = n2; }
    }
    if (d < kMeanDepth) owner = ((t & ((1 << (kMeanDepth - d)) - 1)) == 0);
    s_sub[t] = owner ? subtree_from_leaves(J.leaf, off, len) : 0.f;
    __syncthreads();
    // ---- C
    if (t < 32) {
        int len5 = n; bool own5 = true; int d5 = 0;
        for (; d5 < 5; ++d5) {
            if (len5 <= 128) break;
            int n2 = len5 / 2; n2 -= n2 % 8;
            if ((t >> (4 - d5)) & 1) { len5 -= n2; } else { len5 = n2; }
        }
        if (d5 < 5) own5 = ((t & ((1 << (5 - d5)) - 1)) == 0);
        float v = 0.f;
        if (own5) {
            if (d5 < 5) v = s_sub[t << 5];   // early leaf: its value sits at the left-most slot
            else v = combine_levels(s_sub, len5, t, 5, 0);
        }
        s_mid[t] = v;
    }
    __syncthreads();
    if (t == 0) {
        float sum = combine_levels(s_mid, n, 0, 5, 0);
        float mean = sum / (float)n;                 // 0/0 -> NaN for an empty column, like np.mean
        *J.out = mean;
    }
}

// ---------------------------------------------------------------------------------------------
// slab min/max of x (get_end_points_of_road / _segment, pcl.py:271-313)
// ---------------------------------------------------------------------------------------------
constexpr int kSlabThreads = 256;

__global__ void __launch_bounds__(kSlabThreads)
slab_kernel(const SlabJob* __restrict__ jobs) {
    const SlabJob J = jobs[blockIdx.y];
    const int n = *J.n;
    uint32_t kmin = 0xffffffffu, kmax = 0u; int cnt = 0;
    for (int i = blockIdx.x * kSlabThreads + threadIdx.x; i < n; i += gridDim.x * kSlabThreads) {
        const float z = __ldg(J.z + i);
        // use_f32 == 2: no slab test at all (min / max over every row, +-inf and NaN depths included)
        bool in = (J.use_f32 == 2) ? true : (J.use_f32 ? ((z < J.hi32) && (z > J.lo32)) : (((double)z < J.hi) && ((double)z > J.lo)));
        if (in) {
            uint32_t k = f2key(__ldg(J.x + i));
            kmin = min(kmin, k); kmax = max(kmax, k); ++cnt;
        }
    }
    kmin = warp_min(kmin); kmax = warp_max(kmax); cnt = warp_sum(cnt);
    if (lane_id() == 0 && cnt > 0) {
        atomicMin(&J.keys[0], kmin); atomicMax(&J.keys[1], kmax); atomicAdd(J.count, cnt);
    }
}

}  // namespace sd

int sd_launch_compact(const sd::CompactJob* d_jobs, int njobs, int cap, cudaStream_t st) {
    using namespace sd;
    if (njobs <= 0) return SD_OK;
    int tiles = max(1, ceil_div(cap, kCompactTile));
    int target = max(1, (148 * SD_COMPACT_MINB) / njobs);           // every job's CTAs are resident together
    dim3 grid(min(tiles, target), njobs);
    compact_kernel<<<grid, kCompactThreads, 0, st>>>(d_jobs);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

int sd_launch_mean(const sd::MeanJob* d_jobs, int njobs, int cap, cudaStream_t st) {
    using namespace sd;
    if (njobs <= 0) return SD_OK;
    // leaf sums: 64-element slots, 128 per CTA and trip; enough CTAs for one trip of a full-size column, at most a wave
    mean_leaf_kernel<<<dim3(max(1, min(ceil_div(cap, 64 * (kMeanThreads >> 3)), (148 * 2) / njobs)), njobs), kMeanThreads, 0, st>>>(d_jobs);
    mean_kernel<<<njobs, kMeanThreads, 0, st>>>(d_jobs);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

int sd_launch_slab(const sd::SlabJob* d_jobs, int njobs, int cap, cudaStream_t st) {
    using namespace sd;
    if (njobs <= 0) return SD_OK;
    int per = max(1, min(ceil_div(cap, kSlabThreads * 8), max(1, (148 * 4) / njobs)));
    dim3 grid(per, njobs);
    slab_kernel<<<grid, kSlabThreads, 0, st>>>(d_jobs);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
