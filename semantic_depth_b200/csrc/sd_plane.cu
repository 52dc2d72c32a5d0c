// Least-squares plane fit (pcl.remove_noise_by_fitting_plane, pcl.py:84-209) and the last step of
// the frame: rw (semantic_depth.py:254-259) and f2f (:317-324, pcl.py:212-237,316-318).
//
// The reference solves  w ~ C0*u + C1*v + C2  with scipy.linalg.lstsq (LAPACK gelsd, fp64) over all
// points.  Here: one streaming pass accumulates the nine fp64 moments of the cloud *shifted by its
// first point* (keeps the normal equations well conditioned), per-CTA partials are merged in a fixed
// order by the last CTA of the job (deterministic), which also solves the 3x3 system.  Coefficients
// agree with gelsd to ~1e-13 relative, so inlier sets match unless a residual lies within ~1e-12 of
// the threshold (documented tie class).  A rank-deficient cloud gets gelsd's minimum-norm
// solution (closed form on the centred moments), so the chain continues exactly like the reference's.
#include "sd_internal.cuh"

namespace sd {

constexpr int kPlaneThreads = 256;

__device__ __forceinline__ void pick_uvw(int axis, float x, float y, float z, float& u, float& v, float& w) {
    if (axis == 0) { u = y; v = z; w = x; }         // pcl.py:118-119
    else if (axis == 1) { u = x; v = z; w = y; }    // pcl.py:152-153
    else { u = x; v = y; w = z; }                   // pcl.py:184-185
}

__global__ void __launch_bounds__(kPlaneThreads)
plane_moments_kernel(const PlaneJob* __restrict__ jobs) {
    __shared__ double s_part[kPlaneThreads / 32][kPlaneSums];
    __shared__ int s_last;
    const PlaneJob J = jobs[blockIdx.y];
    // masked clouds: rows [0, n) exist, the alive ones (flag) take part; with `mark` the MAD filter that precedes the fit in
    // the reference (remove_noise_by_mad, then remove_noise_by_fitting_plane) is evaluated here and the bytes rewritten
    const bool marking = J.mark.col != nullptr;
    const int n = J.n_loop ? *J.n_loop : *J.n;
    const int tid = threadIdx.x;
    const float mk_med = marking ? *J.mark.med : 0.f, mk_mad = marking ? *J.mark.mad : 1.f;
    const uint8_t* flag = J.flag;

    float u0 = 0.f, v0 = 0.f, w0 = 0.f;
    if (n > 0) {                                      // any finite point of the surviving cloud serves as the shift
        const int r0 = J.shift_row ? min(max(*J.shift_row, 0), n - 1) : 0;
        pick_uvw(J.axis, __ldg(J.x + r0), __ldg(J.y + r0), __ldg(J.z + r0), u0, v0, w0);
    }
    double h0 = 0, h1 = 0, h2 = 0;
    if (J.use_inliers) { h0 = J.hyp[0]; h1 = J.hyp[1]; h2 = J.hyp[2]; }

    double s[kPlaneSums];
#pragma unroll
    for (int k = 0; k < kPlaneSums; ++k) s[k] = 0.0;
#ifndef SD_PLANE_ILP
#define SD_PLANE_ILP 1
#endif
    // SD_PLANE_ILP rows per trip, their loads issued together (the loop is bound by load latency, not by its fp64 work); the rows
    // of a trip are accumulated in ascending order, so the sums do not depend on the unroll factor
    constexpr int ILP = SD_PLANE_ILP;
    const int stride = gridDim.x * kPlaneThreads;
    for (int i0 = blockIdx.x * kPlaneThreads + tid; i0 < n; i0 += stride * ILP) {
        bool alive[ILP]; float px[ILP], py[ILP], pz[ILP], mc[ILP];
#pragma unroll
        for (int r = 0; r < ILP; ++r) {
            const int i = i0 + r * stride;
            const bool in = i < n;
            const int ic = in ? i : i0;
            alive[r] = in && (flag ? (flag[ic] != 0) : true);
            mc[r] = marking ? __ldg(J.mark.col + ic) : 0.f;
            px[r] = __ldg(J.x + ic); py[r] = __ldg(J.y + ic); pz[r] = __ldg(J.z + ic);
        }
#pragma unroll
        for (int r = 0; r < ILP; ++r) {
            const int i = i0 + r * stride;
            if (i >= n) continue;
            if (marking) {
                const float ad = fabsf(mc[r] - mk_med);                            // pcl.py:79
                const float pen = (0.6745f * ad) / mk_mad;                          // pcl.py:63
                alive[r] = alive[r] && (pen < J.mark.thr);                          // pcl.py:67
                J.flag_out[i] = alive[r] ? 1 : 0;
            }
            if (!alive[r]) continue;
            float u, v, w;
            pick_uvw(J.axis, px[r], py[r], pz[r], u, v, w);
            if (J.use_inliers) {
                double a = ((h0 * (double)u + h1 * (double)v) - (double)w) + h2;
                if (!(fabs(a) < J.thr)) continue;
            }
            const double du = (double)u - (double)u0, dv = (double)v - (double)v0, dw = (double)w - (double)w0;
            s[0] += du * du; s[1] += du * dv; s[2] += du;
            s[3] += dv * dv; s[4] += dv;      s[5] += du * dw;
            s[6] += dv * dw; s[7] += dw;      s[8] += 1.0;
        }
    }
#pragma unroll
    for (int k = 0; k < kPlaneSums; ++k) s[k] = warp_sum(s[k]);
    if (lane_id() == 0) {
#pragma unroll
        for (int k = 0; k < kPlaneSums; ++k) s_part[warp_id()][k] = s[k];
    }
    __syncthreads();
    if (tid < kPlaneSums) {
        double t = 0.0;
        for (int w = 0; w < kPlaneThreads / 32; ++w) t += s_part[w][tid];
        J.partials[blockIdx.x * kPlaneSums + tid] = t;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(J.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    // ---- last CTA: fixed-order merge of the partials, then the 3x3 solve
    __shared__ double s_tot[kPlaneSums];
    if (tid < kPlaneSums) {
        double t = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(&J.partials[b * kPlaneSums + tid]);
        s_tot[tid] = t;
    }
    __syncthreads();
    if (tid == 0) {
        *J.ticket = 0;
        const double Suu = s_tot[0], Suv = s_tot[1], Su = s_tot[2], Svv = s_tot[3], Sv = s_tot[4];
        const double Suw = s_tot[5], Svw = s_tot[6], Sw = s_tot[7], N = s_tot[8];
        if (marking && J.n_mark_out) *J.n_mark_out = (int)N;                    // survivors of the MAD filter
        double c0, c1, c2;
        const double qnan = __longlong_as_double(0x7ff8000000000000ull);
        if (N < 1.0) {
            c0 = c1 = c2 = qnan;
            if (J.status && J.empty_bit) atomicOr(J.status, J.empty_bit);
        } else {
            // eliminate the intercept: centred second moments
            const double cuu = Suu - Su * Su / N, cuv = Suv - Su * Sv / N, cvv = Svv - Sv * Sv / N;
            const double cuw = Suw - Su * Sw / N, cvw = Svw - Sv * Sw / N;
            const double det = cuu * cvv - cuv * cuv;
            const double scale = cuu * cvv;
            if (!(det > 1e-14 * scale) || !(scale > 0.0)) {
                // Rank-deficient design matrix [u v 1]: scipy.linalg.lstsq (gelsd, pcl.py:120,154,186) returns the
                // MINIMUM-NORM least-squares solution in the original coordinates and the chain continues.  The ones
                // column is always independent, so the rank is 1 + rank of the centred 2x2 moment matrix:
                //   rank 0 (all (u, v) equal): one equation c0*um + c1*vm + c2 = wm  ->  C = wm * (um, vm, 1) / (um^2 + vm^2 + 1)
                //   rank 1 (collinear (u, v), e.g. a constant column): (c0, c1) = p + t * e_perp along the null direction,
                //           c2 = wm - c0*um - c1*vm; t minimises |C|^2.
                // (u, v) collinear only up to rounding sit on gelsd's own rank cut-off (rcond = eps): documented tie class.
                const double um = (double)u0 + Su / N, vm = (double)v0 + Sv / N, wm = (double)w0 + Sw / N;
                const double tr = cuu + cvv;
                if (!(tr > 0.0)) {
                    const double q = 1.0 / (um * um + vm * vm + 1.0);
                    c0 = wm * um * q; c1 = wm * vm * q; c2 = wm * q;
                } else {
                    double e0, e1;                                         // unit eigenvector of the non-zero eigenvalue (= tr)
                    if (cuu >= cvv) { e0 = cuu; e1 = cuv; } else { e0 = cuv; e1 = cvv; }
                    const double en = sqrt(e0 * e0 + e1 * e1);
                    e0 /= en; e1 /= en;
                    const double pr = (e0 * cuw + e1 * cvw) / tr;          // particular solution p = e (e . r) / lambda
                    const double p0 = e0 * pr, p1 = e1 * pr;
                    const double a2 = wm - p0 * um - p1 * vm;              // C(t) = a + t d
                    const double d0 = -e1, d1 = e0, d2 = e1 * um - e0 * vm;
                    const double t = -(p0 * d0 + p1 * d1 + a2 * d2) / (d0 * d0 + d1 * d1 + d2 * d2);
                    c0 = p0 + t * d0; c1 = p1 + t * d1; c2 = a2 + t * d2;
                }
            } else {
                c0 = (cuw * cvv - cvw * cuv) / det;
                c1 = (cvw * cuu - cuw * cuv) / det;
                const double c2s = (Sw - c0 * Su - c1 * Sv) / N;        // intercept in shifted coordinates
                c2 = ((double)w0 + c2s) - c0 * (double)u0 - c1 * (double)v0;
            }
            // non-finite moments (inf / NaN coordinates): no plane; scipy's lstsq refuses such input
            if (!(isfinite(c0) && isfinite(c1) && isfinite(c2))) {
                c0 = c1 = c2 = qnan;
                if (J.status) atomicOr(J.status, (uint32_t)SD_ST_SINGULAR_FIT);
            }
        }
        J.coeff[0] = c0; J.coeff[1] = c1; J.coeff[2] = c2;
    }
}

// ---------------------------------------------------------------------------------------------
// finalize: one thread per frame
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void coeff4(int axis, const double* C, double* o) {
    // pcl.py:135 / 168 / 201: Cx*x + Cy*y + Cz*z + C = 0 with the regressed axis' coefficient = -1
    if (axis == 0) { o[0] = -1.0; o[1] = C[0]; o[2] = C[1]; }
    else if (axis == 1) { o[0] = C[0]; o[1] = -1.0; o[2] = C[1]; }
    else { o[0] = C[0]; o[1] = C[1]; o[2] = -1.0; }
    o[3] = C[2];
}

// pcl.planes_intersection_at_certain_depth (pcl.py:212-237): [x y]^T = inv(A) B at z = -depth
__device__ __forceinline__ bool intersect(const double* p1, const double* p2, double depth, double* out) {
    const double z = -depth;
    const double a = p1[0], b = p1[1], c = p2[0], d = p2[1];
    const double b0 = -(p1[2] * z + p1[3]), b1 = -(p2[2] * z + p2[3]);
    const double det = a * d - b * c;
    if (det == 0.0 || !isfinite(det)) return false;
    const double i00 = d / det, i01 = -b / det, i10 = -c / det, i11 = a / det;
    out[0] = i00 * b0 + i01 * b1;
    out[1] = i10 * b0 + i11 * b1;
    out[2] = z;
    return isfinite(out[0]) && isfinite(out[1]);
}

__global__ void finalize_kernel(const FinalJob* __restrict__ jobs, int njobs, SdParams P) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= njobs) return;
    FrameState* fs = jobs[f].fs;
    SdFrameResult* R = jobs[f].out;
    const double qnan = __longlong_as_double(0x7ff8000000000000ull);
    uint32_t status = fs->status;
    for (int k = 0; k < SD_NUM_COUNTS; ++k) R->counts[k] = fs->n[k];
    for (int k = 0; k < 5; ++k) { R->median[k] = fs->med[k]; R->mad[k] = fs->mad[k]; }
    R->fence_mean_x = fs->fence_mean;
    R->sor_mean = fs->sor_stats[0]; R->sor_std = fs->sor_stats[1]; R->sor_thr = fs->sor_stats[2];
    for (int k = 0; k < 3; ++k) R->ransac_best[k] = fs->ransac_best[k];
    // SOR survivors: counted by the radius kernel when both filters run, else implied
    R->counts[SD_CNT_ROAD_SOR] = P.use_sor ? (P.use_ror ? fs->n_sor_alive : fs->n[SD_CNT_ROAD_ROR]) : fs->n[SD_CNT_ROAD_PLANE];
    fs->n_sor_alive = 0;
    // ---- rw (semantic_depth.py:254-259)
    R->counts[SD_CNT_ROAD_SLAB] = fs->slab_count;
    if (fs->n[SD_CNT_ROAD_ROR] == 0) status |= SD_ST_EMPTY_ROAD;
    if (fs->slab_count > 0) {
        R->xl = (double)key2f(fs->slab_keys[0]);
        R->xr = (double)key2f(fs->slab_keys[1]);
        R->rw = fabs(R->xl - R->xr);
    } else {
        R->xl = R->xr = R->rw = qnan;
        status |= SD_ST_NO_SLAB_POINTS;
    }
    fs->slab_keys[0] = 0xffffffffu; fs->slab_keys[1] = 0u; fs->slab_count = 0;
    // ---- f2f (semantic_depth.py:317-324)
    coeff4(1, fs->coeff[0], R->road_coeff);
    coeff4(0, fs->coeff[1], R->left_coeff);
    coeff4(0, fs->coeff[2], R->right_coeff);
    R->f2f = qnan;
    for (int k = 0; k < 3; ++k) { R->left_pt[k] = qnan; R->right_pt[k] = qnan; }
    if (P.approach_both) {
        if (fs->n[SD_CNT_FENCE_ABS_Z] == 0) status |= SD_ST_EMPTY_FENCE;
        bool fits_ok = !(status & SD_ST_SINGULAR_FIT) && fs->n[SD_CNT_ROAD_MAD_X] > 0 &&
                       fs->n[SD_CNT_LEFT_MAD_X] > 0 && fs->n[SD_CNT_RIGHT_MAD_X] > 0;
        if (fits_ok) {
            double pl[3], pr[3];
            bool ok = intersect(R->road_coeff, R->left_coeff, P.depth, pl) &&
                      intersect(R->road_coeff, R->right_coeff, P.depth, pr);
            if (ok) {
                for (int k = 0; k < 3; ++k) { R->left_pt[k] = pl[k]; R->right_pt[k] = pr[k]; }
                const double dx = pl[0] - pr[0], dy = pl[1] - pr[1], dz = pl[2] - pr[2];
                R->f2f = sqrt((dx * dx + dy * dy) + dz * dz);      // np.linalg.norm, pcl.py:318
            } else {
                status |= SD_ST_SINGULAR_PLANES;
            }
        }
    }
    R->status = status;
    fs->status = 0;
}

}  // namespace sd

int sd_launch_plane(const sd::PlaneJob* d_jobs, int njobs, int cap, cudaStream_t st) {
    using namespace sd;
    if (njobs <= 0) return SD_OK;
    int per = max(1, min(kPlaneBlocks, ceil_div(cap, kPlaneThreads * 8)));
    per = max(1, min(per, max(1, (148 * 4) / njobs)));
    dim3 grid(per, njobs);
    plane_moments_kernel<<<grid, kPlaneThreads, 0, st>>>(d_jobs);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

int sd_launch_finalize(const sd::FinalJob* d_jobs, int njobs, const SdParams* params, cudaStream_t st) {
    using namespace sd;
    if (njobs <= 0) return SD_OK;
    finalize_kernel<<<ceil_div(njobs, 64), 64, 0, st>>>(d_jobs, njobs, *params);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
