/* sd_fusion.h -- C ABI of the B200-native SemanticDepth fusion path (libsd_fusion.so).
 *
 * The reference has no native code and no FFI: its fusion stage is Python that calls module-level
 * functions of semantic_depth_lib/pcl.py (NumPy in, NumPy out) plus two Open3D calls.  The entry
 * points below are what a binding for that path binds; each cites the reference interface it
 * replaces (paths relative to /root/reference).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void*;
 *   - every pointer named d_* is DEVICE memory owned by the caller (no hidden allocation; scratch
 *     comes from a caller-provided workspace, see sd_ws_*); h_* is HOST memory;
 *   - clouds are structure-of-arrays: x[], y[], z[] fp32 and an optional int32 src[] carrying the
 *     flat source index of each point (pixel index in the fused path, row index in the per-call ops);
 *   - stable (order-preserving) compaction everywhere: output order == input order, exactly what
 *     NumPy boolean / index-array selection produces;
 *   - return value: SD_OK (0) or a negative SD_ERR_* code; sd_last_error() gives a message;
 *   - all work is enqueued on `stream`; functions that return a count to the host synchronise that
 *     stream once, sd_fuse_frames() never synchronises.
 */
#ifndef SD_FUSION_H
#define SD_FUSION_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_ABI_VERSION 2

enum {
    SD_OK = 0,
    SD_ERR_INVALID = -1,      /* bad argument (null pointer, size, axis, k ...) */
    SD_ERR_CUDA = -2,         /* a CUDA runtime call failed; see sd_last_error() */
    SD_ERR_WORKSPACE = -3,    /* workspace too small / not initialised */
    SD_ERR_UNSUPPORTED = -4
};

/* Per-frame status bits of the fused path.  The reference signals these conditions with Python
 * exceptions / None (pcl.py:107,232,303-304; semantic_depth.py:259); the facade maps them back. */
enum {
    SD_ST_EMPTY_ROAD = 1 << 0,
    SD_ST_EMPTY_FENCE_LEFT = 1 << 1,
    SD_ST_EMPTY_FENCE_RIGHT = 1 << 2,
    SD_ST_MAD_ZERO = 1 << 3,
    SD_ST_NO_SLAB_POINTS = 1 << 4,
    SD_ST_SINGULAR_PLANES = 1 << 5,
    SD_ST_EMPTY_FENCE = 1 << 6,
    SD_ST_SINGULAR_FIT = 1 << 7
};

/* Per-stage point counts reported for every frame (the reference's cloud sizes after each call of
 * semantic_depth.py:183-309, in call order). */
enum {
    SD_CNT_ROAD_GATHER = 0,   /* points3D[road_mask]            :183 */
    SD_CNT_FENCE_GATHER,      /* points3D[fence_mask]           :186 */
    SD_CNT_ROAD_Z,            /* remove_from_to                 :206 */
    SD_CNT_ROAD_MAD_Y,        /* remove_noise_by_mad(1, 15)     :209 */
    SD_CNT_ROAD_MAD_X,        /* remove_noise_by_mad(0, 2)      :212 */
    SD_CNT_ROAD_PLANE,        /* remove_noise_by_fitting_plane  :215-219 */
    SD_CNT_ROAD_SOR,          /* statistical_outlier_removal    :234-236 */
    SD_CNT_ROAD_ROR,          /* radius_outlier_removal         :238-241 */
    SD_CNT_ROAD_SLAB,         /* get_end_points_of_road slab    :254-255 */
    SD_CNT_FENCE_MAD_Y,       /* :279 */
    SD_CNT_FENCE_ABS_Z,       /* :283-284 */
    SD_CNT_LEFT_SPLIT,        /* :286-287 */
    SD_CNT_RIGHT_SPLIT,
    SD_CNT_LEFT_MAD_X,        /* :291 */
    SD_CNT_LEFT_PLANE,        /* :294-298 */
    SD_CNT_RIGHT_MAD_X,       /* :302 */
    SD_CNT_RIGHT_PLANE,       /* :305-309 */
    SD_NUM_COUNTS
};

/* Camera of DepthFrame.compute_3D_points (semantic_depth.py:691-696): the four non-trivial entries
 * of the reference's float32 Q matrix, and the fp32 disparity scale of semantic_depth.py:145. */
typedef struct SdCamera {
    float q03;             /* -cx  */
    float q13;             /*  cy  */
    float q23;             /* -f   */
    float q32;             /* 1/b  */
    float disparity_mult;  /* original image width (or 3800 in the sequence driver) */
} SdCamera;

/* The literals of FrameProcessor.process_frame's fusion section (semantic_depth.py:206-309). */
typedef struct SdParams {
    double prob_thr;          /* 0.5   :556,564 */
    float road_z_to_meter;    /* 7.0   :206 */
    float road_mad_y_thr;     /* 15.0  :209 */
    float road_mad_x_thr;     /* 2.0   :212 */
    float fence_mad_y_thr;    /* 5.0   :279 */
    float fence_abs_z_thr;    /* 35.0  :283 */
    float left_mad_x_thr;     /* 5.0   :291 */
    float right_mad_x_thr;    /* 1.0   :302 */
    int32_t sor_nb_neighbors; /* 10    :235 */
    double road_plane_thr;    /* 5.0   :218 */
    double fence_plane_thr;   /* 1.0   :297,308 */
    double sor_std_ratio;     /* 0.5   :235 */
    double ror_radius;        /* 0.5   :239 */
    double slab_lo;           /* -((depth-0.02)+0.05), computed by the host as Python does (pcl.py:283) */
    double slab_hi;           /* -((depth-0.02)-0.05) */
    double depth;             /* 10.0  :736-738, used by the plane intersections :317-323 */
    int32_t ror_nb_points;    /* 80    :239 */
    int32_t use_sor;          /* 1 */
    int32_t use_ror;          /* 1 */
    int32_t approach_both;    /* 1 = rw and f2f, 0 = rw only (:273) */
    int32_t label_mode;       /* 0 = softmax > prob_thr (the reference, :555-556,563-564); 1 = argmax over the classes */
    int32_t pad_;
} SdParams;

/* One frame's answers (what process_frame returns at :460 plus everything observable on the way). */
typedef struct SdFrameResult {
    double rw;                 /* |xL - xR|  :259 ; NaN when the slab is empty */
    double f2f;                /* :324 ; NaN when unavailable */
    double xl, xr;             /* min / max x of the slab (left / right end point, pcl.py:307-308) */
    double left_pt[3];         /* road-plane x left-fence-plane at z = -depth (:317-319) */
    double right_pt[3];        /* :321-323 */
    double road_coeff[4];      /* Cx, Cy, Cz, C  (pcl.py:168) */
    double left_coeff[4];      /* pcl.py:135 */
    double right_coeff[4];
    double sor_mean, sor_std, sor_thr;
    float median[5];           /* medians of the five MAD calls: road y, road x, fence y, left x, right x */
    float mad[5];
    float fence_mean_x;        /* np.mean of extract_pcls (pcl.py:258) */
    uint32_t status;           /* SD_ST_* */
    int32_t counts[SD_NUM_COUNTS];
    int32_t ransac_best[3];    /* best hypothesis index for road / left / right, -1 without RANSAC */
} SdFrameResult;

/* ---- library ------------------------------------------------------------------------------- */
int sd_abi_version(void);
const char* sd_last_error(void);
/* Fills the reference's literals (semantic_depth.py:206-309) for the given depth (default 10). */
void sd_default_params(SdParams* p, double depth);

/* ---- workspace ------------------------------------------------------------------------------ */
typedef struct SdWorkspace SdWorkspace;   /* host-side handle over caller-owned device memory */
/* Bytes of device memory needed for `max_frames` frames of `height` x `width` pixels processed
 * concurrently by sd_fuse_frames (also serves every per-call op on clouds of up to height*width
 * points).  `max_hypotheses` > 0 reserves room for the RANSAC variant. */
size_t sd_ws_bytes(int max_frames, int height, int width, int max_hypotheses);
/* Carves `d_mem` (256-byte aligned, sd_ws_bytes() long), zeroes the control words and uploads the
 * job descriptors.  Synchronises `stream` once. */
int sd_ws_create(SdWorkspace** out, void* d_mem, size_t bytes, int max_frames, int height, int width,
                 int max_hypotheses, void* stream);
void sd_ws_destroy(SdWorkspace* ws);

/* ---- pixel stage ---------------------------------------------------------------------------- */
/* Fused: labels (softmax > thr, semantic_depth.py:555-556,563-564) + post_processing blend
 * (:656-664,676) + disparity scale (:145) + reprojectImageTo3D (:691-696) + raster-ordered gather
 * of the road / fence points (:183-187) + the road z cut (:206).
 *   d_logits [B][H*W][3] fp32, d_disp [B][2][H][W] fp32, d_lmask/d_rmask [W] fp64 ramps (:661-663).
 * Optional outputs (may be NULL): d_labels [B][H*W] uint8 (bit0 road, bit1 fence);
 *   d_points [B][H*W][3] fp32 = the full points3D array of :160; d_disp_pp [B][H*W] fp32 = :676.
 * Cloud outputs live in the workspace (sd_fuse_frames) or in caller arrays (this call):
 *   road_* receives the road points with z < -to_meter, fence_* all fence points, both with src =
 *   flat pixel index; capacity H*W each, per frame stride H*W.  d_counts [B][3] int32 =
 *   {road_gather, road_z, fence_gather}. */
#define SD_PIX_RAW_DISPARITY 1   /* flags: d_disp[b][0] already is the blended, scaled disparity of :145 (skip blend and scale) */
#define SD_PIX_LABEL_ARGMAX 2    /* flags: label by argmax over the three classes instead of softmax > prob_thr */
int sd_pixel_fuse(const float* d_logits, const float* d_disp, const double* d_lmask, const double* d_rmask,
                  int batch, int height, int width, const SdCamera* cam, double prob_thr, float road_z_to_meter, int flags,
                  float* d_road_x, float* d_road_y, float* d_road_z, int32_t* d_road_src,
                  float* d_fence_x, float* d_fence_y, float* d_fence_z, int32_t* d_fence_src,
                  int32_t* d_counts, uint8_t* d_labels, float* d_points, float* d_disp_pp,
                  SdWorkspace* ws, void* stream);

/* Same with the FCN-8s head left unexpanded: the logits are never materialised, the label
 * kernel evaluates fcn8s/fcn.py:207-213 -- conv2d_transpose(second_skip, 3, 16x16, stride 8, 'same') -- itself.
 *   d_scores [B][H/8][W/8][3] fp32 (second_skip), d_up_weights [16][16][3 out][3 in] fp32 (TF kernel layout),
 *   d_up_bias [3] fp32.  Arithmetic contract: fp32, no FMA, taps accumulated from 0.0 in the order (low-res row,
 *   low-res column, input channel) ascending, bias added last.  d_logits_out (optional) [B][H*W][3] receives the
 *   upsampled logits, i.e. the tensor the reference fetches as 'logits:0' (fcn.py:241). */
int sd_pixel_fuse_scores(const float* d_scores, const float* d_up_weights, const float* d_up_bias,
                         const float* d_disp, int batch, int height, int width, const SdCamera* cam,
                         double prob_thr, float road_z_to_meter, int flags,
                         float* d_road_x, float* d_road_y, float* d_road_z, int32_t* d_road_src,
                         float* d_fence_x, float* d_fence_y, float* d_fence_z, int32_t* d_fence_src,
                         int32_t* d_counts, uint8_t* d_labels, float* d_points, float* d_disp_pp, float* d_logits_out,
                         SdWorkspace* ws, void* stream);

/* ---- next to the path -------------------------------------------------- */
/* cv2.resize(frame, (dst_width, dst_height), interpolation=cv2.INTER_CUBIC) on uint8 frames, the first step of
 * process_frame (semantic_depth.py:110-112): OpenCV's fixed-point definition (11-bit weights, integer passes,
 * (v + 2^21) >> 22).  d_src [batch][src_height][src_width][channels], d_dst likewise; channels in {1, 3, 4}. */
int sd_resize_cubic_u8(const uint8_t* d_src, int batch, int src_height, int src_width, int channels,
                       uint8_t* d_dst, int dst_height, int dst_width, void* stream);

/* the vertex rows of PointCloud2Ply.write_ply (semantic_depth_lib/point_cloud_2_ply.py:62-70),
 * np.savetxt(f, hstack([points3D, colors]), '%f %f %f %d %d %d'), byte for byte.  d_rgb [n][3] uint8; d_out receives
 * *h_nbytes bytes of ASCII.  Returns SD_ERR_WORKSPACE with the needed size in *h_nbytes when `capacity` is too small.
 * Synchronises the stream. */
int sd_ply_rows(const float* d_x, const float* d_y, const float* d_z, const uint8_t* d_rgb, int n,
                char* d_out, unsigned long long capacity, unsigned long long* h_nbytes, SdWorkspace* ws, void* stream);
/* The same for float64 clouds (the reference's clouds are float64 after the Open3D round trip, semantic_depth.py:244,
 * and its plane meshes / lines are float64 throughout, pcl.py:107-113,321-331).  Finite |values| must be < 2^128. */
int sd_ply_rows_f64(const double* d_x, const double* d_y, const double* d_z, const uint8_t* d_rgb, int n,
                    char* d_out, unsigned long long capacity, unsigned long long* h_nbytes, SdWorkspace* ws, void* stream);

/* SURVEY.md 8f rank 4 (mask paste): the overlaid frame SegmentFrame.segment_frame returns (semantic_depth.py:547-568):
 * toimage(np.dot(mask, [[r, g, b, a]]), mode="RGBA") pasted with itself as the mask, road first, then fence.
 * d_frame / d_out [B][H][W][3] uint8 (may alias), d_labels [B][H*W] uint8 (bit0 road, bit1 fence: the pixel stage's
 * d_labels), road_rgba / fence_rgba 4 host ints in [0, 255] (reference: {128, 64, 128, 64} and {160, 10, 10, 64}),
 * d_scratch 2*B int32 of device memory owned by the caller.  Never synchronises. */
int sd_overlay_masks(const uint8_t* d_frame, const uint8_t* d_labels, int batch, int height, int width,
                     const int32_t* road_rgba, const int32_t* fence_rgba, uint8_t* d_out, int32_t* d_scratch, void* stream);

/* SURVEY.md 8f rank 4 (banner + putText): the result banner process_frame draws on the segmented frame
 * (semantic_depth.py:339-394: cv2.rectangle(frame, (0,0), (w, int(0.2*h)), (156,157,159), -1) and up to seven
 * cv2.putText(..., fontFace=16, fontScale 2 / 4, thickness 2 / 5) lines; live twin
 * semantic_depth_cityscapes_sequence.py:304-327 with font scales 2 and 2.2), in place on d_frames [B][H][W][3] uint8.
 *   d_rects  [n_rects][6]  int32 {frame, x0, y0, x1, y1, colour}: filled, corners inclusive and ordered, clipped to the frame
 *   d_places [n_places][5] int32 {frame, bitmap, x, y, colour}: bitmap `bitmap` of the atlas pasted with its cell's
 *            top-left pixel at (x, y), clipped to the frame
 *   d_glyph_bits [n_bitmaps][cell_height][cell_words] uint32: bit b of word k of a row = pixel 32*k + b of that row
 *   colour = channel0 | channel1 << 8 | channel2 << 16 (the frame's own channel order, like cv2's Scalar).
 * The atlas (semantic_depth_b200/data/hershey_atlas.npz) is baked from OpenCV's own Hershey rasteriser for the reference's
 * (fontFace, fontScale, thickness) presets; the pen arithmetic that picks bitmap and position (16.16 fixed point, cvRound)
 * is host glue (semantic_depth_b200/frame_ops.py).  Rectangles are drawn before glyphs; glyphs of one call may overlap
 * only if they share a colour (one call per run of equally coloured putText lines keeps cv2's drawing order).
 * Never synchronises. */
int sd_draw_banner(uint8_t* d_frames, int batch, int height, int width, const int32_t* d_rects, int n_rects,
                   const uint32_t* d_glyph_bits, int n_bitmaps, int cell_height, int cell_words,
                   const int32_t* d_places, int n_places, void* stream);

/* ---- per-call cloud ops (the pcl.py call surface; n is known to the host) -------------------- */
/* np.median of a column (pcl.py:78,80): h_out[0] = median(col), h_out[1] = median(|col - median|). */
int sd_median_mad(const float* d_col, int n, float* h_out, SdWorkspace* ws, void* stream);

/* Predicate kinds of sd_filter(). */
enum {
    SD_PRED_LT = 0,        /* col < fa                                  remove_from_to      pcl.py:36 (fa = -to_meter) */
    SD_PRED_ABS_LT = 1,    /* |col| < fa                                threshold_complete  pcl.py:247 */
    SD_PRED_MAD = 2,       /* fl(fl(0.6745f*|col-f0|)/f1) < fa          remove_noise_by_mad pcl.py:63-67 (f0 = median, f1 = mad) */
    SD_PRED_PLANE = 3,     /* |((d0*u + d1*v) - w) + d2| < da  (fp64)   remove_noise_by_fitting_plane pcl.py:130-131 */
    SD_PRED_GT = 4,        /* col > fa                                  extract_pcls right  pcl.py:264 */
    SD_PRED_SLAB = 5,      /* d0 < z < d1 (fp64) or fl32 bounds f0 < z < f1 when use_f32 */
    SD_PRED_SOR = 6,       /* 0 < avg[i] < da                           Open3D RemoveStatisticalOutliers */
    SD_PRED_ROR = 7        /* cnt[i] > ia                               Open3D RemoveRadiusOutliers */
};
typedef struct SdPredicate {
    int32_t kind;
    int32_t axis;          /* column the predicate reads (plane: the regressed axis) */
    int32_t ia;
    int32_t use_f32;
    float fa, f0, f1;
    float pad_;
    double da, d0, d1, d2;
    const void* d_aux;     /* SOR: const double* avg ; ROR: const int32_t* counts */
} SdPredicate;

/* Stable filter of one cloud.  Writes the surviving points to out_* (any may be NULL), out_src[i] =
 * in_src[j] (or j itself when d_in_src is NULL, i.e. the kept row indices) and returns the number
 * kept through h_n_out.  Synchronises the stream. */
int sd_filter(const float* d_x, const float* d_y, const float* d_z, const int32_t* d_in_src, int n,
              const SdPredicate* pred,
              float* d_out_x, float* d_out_y, float* d_out_z, int32_t* d_out_src, int32_t* h_n_out,
              SdWorkspace* ws, void* stream);

/* Least-squares plane w = C0*u + C1*v + C2 of pcl.py:118-120 / 152-154 / 184-186 (axis = regressed
 * coordinate).  h_coeff[3] = C0, C1, C2 (fp64).  Returns SD_OK and sets *h_singular when the normal
 * equations are singular. */
int sd_plane_fit(const float* d_x, const float* d_y, const float* d_z, int n, int axis,
                 double* h_coeff, int32_t* h_singular, SdWorkspace* ws, void* stream);

/* np.mean of an fp32 column with NumPy's pairwise fp32 summation order (pcl.py:258). */
int sd_mean_f32(const float* d_col, int n, float* h_mean, SdWorkspace* ws, void* stream);

/* min / max of x over the points with lo < z < hi and their count (pcl.py:283,307-308).  use_f32: 0 = fp64 bounds,
 * 1 = bounds rounded to fp32 (NumPy's rule for an fp32 cloud), 2 = no slab test: every row counts, whatever its z
 * (np.amin / np.amax over a whole column, pcl.py:307-308; point_cloud_2_ply.py:87). */
int sd_slab_minmax(const float* d_x, const float* d_z, int n, double lo, double hi, int use_f32,
                   float* h_xmin, float* h_xmax, int32_t* h_count, SdWorkspace* ws, void* stream);

/* Open3D RemoveStatisticalOutliers' per-point quantity (semantic_depth.py:234-235): mean fp64
 * distance to the k nearest neighbours, self included.  d_avg [n] fp64.  Also returns the cloud
 * mean / Bessel std / threshold through h_stats[3] when not NULL. */
int sd_knn_mean_distance(const float* d_x, const float* d_y, const float* d_z, int n, int k, double std_ratio,
                         double* d_avg, double* h_stats, SdWorkspace* ws, void* stream);

/* Open3D RemoveRadiusOutliers' per-point quantity (semantic_depth.py:238-239): number of points
 * within `radius` (self included), saturated at `cap`+1 when cap >= 0 (only count > cap matters). */
int sd_radius_count(const float* d_x, const float* d_y, const float* d_z, int n, double radius, int cap,
                    int32_t* d_counts, SdWorkspace* ws, void* stream);

/* RANSAC scoring (north_star row 8-R): inlier count of each hypothesis plane through the points
 * d_triplets[k][3]; d_hyp_counts [K] int32; h_best = arg max (lowest index on ties). */
int sd_ransac_score(const float* d_x, const float* d_y, const float* d_z, int n, int axis, double threshold,
                    const int32_t* d_triplets, int n_hyp, int32_t* d_hyp_counts, int32_t* h_best,
                    double* h_best_coeff, SdWorkspace* ws, void* stream);

/* ---- the fused per-frame path ---------------------------------------------------------------- */
/* Everything between the two networks' raw outputs and (rw, f2f) for `batch` frames, enqueued on
 * `stream` without any host synchronisation (CUDA-graph capturable).  d_results [batch].
 * d_hyp_* : optional RANSAC triplets per chain ([batch][n_hyp][3] each, may be NULL = reference
 * behaviour, all-points least squares). */
int sd_fuse_frames(const float* d_logits, const float* d_disp, int batch, int height, int width,
                   const SdCamera* cam, const SdParams* params,
                   const int32_t* d_hyp_road, const int32_t* d_hyp_left, const int32_t* d_hyp_right, int n_hyp,
                   SdFrameResult* d_results, SdWorkspace* ws, void* stream);

/* The fused path fed by the unexpanded FCN-8s head (see sd_pixel_fuse_scores): 0.19 B/pixel of scores instead of
 * 12 B/pixel of logits. */
int sd_fuse_frames_scores(const float* d_scores, const float* d_up_weights, const float* d_up_bias,
                          const float* d_disp, int batch, int height, int width,
                          const SdCamera* cam, const SdParams* params, SdFrameResult* d_results,
                          SdWorkspace* ws, void* stream);

/* Number of kernels one sd_fuse_frames call launches for these parameters (independent of batch). */
int sd_fuse_kernel_count(const SdParams* params, int with_ransac);

/* Same through HOST buffers (the reference-facing call: NumPy arrays in, results out): copies the
 * inputs host->device, runs sd_fuse_frames and copies the results back, all on `stream`, then
 * synchronises.  d_logits_stage / d_disp_stage are caller-owned device staging buffers. */
int sd_fuse_frames_host(const float* h_logits, const float* h_disp, int batch, int height, int width,
                        const SdCamera* cam, const SdParams* params,
                        float* d_logits_stage, float* d_disp_stage, SdFrameResult* d_results,
                        SdFrameResult* h_results, SdWorkspace* ws, void* stream);

/* ---- FCN-8s decoder head (the producer side of the score-map mode) ------------------------------ */
/* fcn8s/fcn.py:159-205 (`layers` up to second_skip): three 1x1 convolutions of the VGG feature maps to 3 classes, two
 * 4x4 / stride-2 transposed convolutions and the two skip additions.  Its output is the `d_scores` input of
 * sd_fuse_frames_scores / sd_pixel_fuse_scores, whose label kernel evaluates the last layer (fcn.py:207-213) itself.
 *   d_layer3 [B][h8][w8][c3], d_layer4 [B][h8/2][w8/2][c4], d_layer7 [B][h8/4][w8/4][c7]  fp32 NHWC (vgg_layer3/4/7_out,
 *   fcn.py:95-103; 256 / 512 / 4096 channels in VGG16);  weights in TF layouts: conv*_w [c][3] (kernel [1][1][in][out]),
 *   deconv*_w [4][4][3 out][3 in]; biases [3];  d_scores [B][h8][w8][3] receives second_skip.
 * Arithmetic contract (TensorFlow's summation order is not reproducible): fp32, no FMA; a 1x1 convolution sums the
 * channels l, l+32, l+64, ... per lane l from 0.0, combines the 32 lanes by the xor tree 16, 8, 4, 2, 1 and adds the bias
 * last; a transposed convolution accumulates (input row, input column, input channel) ascending from 0.0, then bias,
 * then the skip tensor.  d_scratch: sd_fcn8s_head_scratch_bytes(batch, h8, w8) bytes of device memory. */
typedef struct SdFcnHeadWeights {
    const float* conv3_w; const float* conv3_b;
    const float* conv4_w; const float* conv4_b;
    const float* conv7_w; const float* conv7_b;
    const float* deconv1_w; const float* deconv1_b;
    const float* deconv2_w; const float* deconv2_b;
} SdFcnHeadWeights;
size_t sd_fcn8s_head_scratch_bytes(int batch, int h8, int w8);
int sd_fcn8s_head(const float* d_layer3, const float* d_layer4, const float* d_layer7, int batch, int h8, int w8,
                  int c3, int c4, int c7, const SdFcnHeadWeights* weights, float* d_scratch, size_t scratch_bytes,
                  float* d_scores, void* stream);

/* Optional device-side timing of the fused call: when enabled, sd_fuse_frames records CUDA events on
 * `stream` before / after the pixel-stage kernel and after the last kernel (skipped while the stream is
 * being captured into a CUDA graph).  sd_ws_stage_elapsed_ms: which = 0 pixel-stage kernel,
 * 1 = whole fused call; the stream must have been synchronised past the call. */
int sd_ws_enable_timing(SdWorkspace* ws, int enable);
/* Restrict the next sd_fuse_frames calls to a subset of the path (default 15 = everything):
 *   bit 0 = pixel stage; bit 1 = cloud stages up to the neighbour-search grid, and the whole fence chain;
 *   bit 2 = the k-NN kernel of the statistical filter; bit 3 = radius search, final road compaction,
 *   slab and answers.  Lets a caller capture the segments into separate CUDA graphs and bracket a
 *   kernel with its own events (events recorded inside a captured graph cannot be timed). */
int sd_ws_set_stage_mask(SdWorkspace* ws, int mask);
int sd_ws_stage_elapsed_ms(SdWorkspace* ws, int which, float* h_ms);

/* Per-stage device times of the last fused call, the counterpart of the reference's tic / toc pairs around every stage
 * of process_frame (semantic_depth.py:157-190 segmentation / disparity / 3-D, :203-221 road denoising, :227-245 the two
 * Open3D filters, :253-268 rw, :274-311 fences, :316-332 f2f; dumped at :445-454).  Needs sd_ws_enable_timing(ws, 1)
 * and a full call (stage mask 15) outside stream capture; such a call keeps the fence chain on `stream` so that the
 * stages are disjoint and add up.  h_ms [SD_NUM_STAGES], milliseconds, after the stream has been synchronised. */
enum SdStage {
    SD_STAGE_PIXEL = 0,        /* labels, disparity blend, x mult, reprojection, mask gather, remove_from_to */
    SD_STAGE_ROAD_MAD = 1,     /* remove_noise_by_mad on y and on x */
    SD_STAGE_ROAD_PLANE = 2,   /* remove_noise_by_fitting_plane (and the RANSAC scoring when hypotheses are given) */
    SD_STAGE_ROAD_GRID = 3,    /* search grid of the two Open3D filters */
    SD_STAGE_ROAD_KNN = 4,     /* statistical_outlier_removal: k-NN mean distances + cloud statistics */
    SD_STAGE_ROAD_ROR = 5,     /* radius_outlier_removal + the compaction that applies both filters */
    SD_STAGE_RW = 6,           /* get_end_points_of_road slab scan */
    SD_STAGE_FENCES = 7,       /* the whole fence chain */
    SD_STAGE_ANSWERS = 8,      /* rw, plane intersections, f2f */
    SD_NUM_STAGES = 9
};
int sd_ws_stage_times(SdWorkspace* ws, float* h_ms);

/* Device pointers of a frame's final clouds inside the workspace (valid until the next fuse call):
 * which = 0 road (after ROR), 1 left fence (after plane filter), 2 right fence. */
int sd_ws_cloud(SdWorkspace* ws, int frame, int which, const float** d_x, const float** d_y, const float** d_z,
                const int32_t** d_src, const int32_t** d_n);
/* Device pointers of a frame's intermediate per-stage source-index lists for parity tests:
 * stage = SD_CNT_* ; only stages that materialise a cloud are available (returns SD_ERR_UNSUPPORTED otherwise). */
int sd_ws_stage_src(SdWorkspace* ws, int frame, int stage, const int32_t** d_src);
/* Stages whose filter is NOT materialised by a compaction (the MAD filters in front of a plane fit, when no RANSAC
 * hypotheses are given): the stage's cloud is the rows i < *d_rows of the chain's input with d_alive[i] != 0, in input
 * order; d_src[i] is the source pixel of row i.  stage = SD_CNT_FENCE_MAD_Y, SD_CNT_LEFT_MAD_X or SD_CNT_RIGHT_MAD_X (the road chain's
 * input buffer is reused by its final compaction, so the road MAD stages are observable through their counts only);
 * SD_ERR_UNSUPPORTED when the stage is materialised (use sd_ws_stage_src) or not retained. */
int sd_ws_stage_alive(SdWorkspace* ws, int frame, int stage, const int32_t** d_src, const uint8_t** d_alive, const int32_t** d_rows);

#ifdef __cplusplus
}
#endif
#endif /* SD_FUSION_H */
