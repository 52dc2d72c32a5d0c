// C ABI of libsd_fusion.so (see include/sd_fusion.h): workspace carving, job-descriptor tables, the
// per-call cloud ops behind the semantic_depth_lib.pcl facade, and the fused per-frame path.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>
#include <new>
#include "sd_internal.cuh"

using namespace sd;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
void sd_set_last_cuda_error(int code, const char* what) {
    g_last_error = std::string("CUDA error ") + std::to_string(code) + " (" +
                   cudaGetErrorString((cudaError_t)code) + ") at " + what;
}
static int fail(int code, const char* msg) { g_last_error = msg; return code; }

extern "C" int sd_abi_version(void) { return SD_ABI_VERSION; }
extern "C" const char* sd_last_error(void) { return g_last_error.c_str(); }

extern "C" void sd_default_params(SdParams* p, double depth) {
    memset(p, 0, sizeof(*p));
    p->prob_thr = 0.5;
    p->road_z_to_meter = 7.0f;
    p->road_mad_y_thr = 15.0f; p->road_mad_x_thr = 2.0f;
    p->fence_mad_y_thr = 5.0f; p->fence_abs_z_thr = 35.0f;
    p->left_mad_x_thr = 5.0f; p->right_mad_x_thr = 1.0f;
    p->sor_nb_neighbors = 10;
    p->road_plane_thr = 5.0; p->fence_plane_thr = 1.0;
    p->sor_std_ratio = 0.5; p->ror_radius = 0.5;
    const double d = depth - 0.02;                  // semantic_depth.py:254-255
    p->slab_lo = -(d + 0.05); p->slab_hi = -(d - 0.05);   // pcl.py:283
    p->depth = depth;
    p->ror_nb_points = 80; p->use_sor = 1; p->use_ror = 1; p->approach_both = 1; p->label_mode = 0;
}

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
namespace {

constexpr size_t kAlign = 256;
constexpr int kChains = 4;                     // road, fence, left, right
constexpr size_t kJobsPerFrameBytes = 12 * 1024;
constexpr size_t kScratchBytes = 64 * 1024;
constexpr size_t kPinnedBytes = 16 * 1024;

struct Carver {
    char* base; size_t off; size_t limit; bool dry;
    template <typename T> T* take(size_t count) {
        off = (off + kAlign - 1) / kAlign * kAlign;
        T* p = dry ? nullptr : reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return p;
    }
};

int round_up(int v, int m) { return (v + m - 1) / m * m; }

// device scratch of the per-call ops
struct CallScratch {
    int32_t n_in, n_out, count, singular;
    uint32_t keys[2]; uint32_t status; uint32_t pad;
    float f[8];
    double d[8];
    int32_t best; int32_t pad2;
    SelState sel;
    ScanCtl ctl;
    uint32_t ptick; uint32_t pad3;
    GridState gs;
    double partials[kPlaneBlocks * kPlaneSums];
};

// everything the fused path needs per workspace: device job tables (pointers into ws->jobs)
struct FusedTables {
    SelJob* sel_road_y_med; SelJob* sel_road_y_mad; SelJob* sel_road_x_med; SelJob* sel_road_x_mad;
    SelJob* sel_fence_y_med; SelJob* sel_fence_y_mad; SelJob* sel_side_x_med; SelJob* sel_side_x_mad;   // side: [2F]
    CompactJob* c_road_y; CompactJob* c_road_x; CompactJob* c_road_plane; CompactJob* c_road_final;
    CompactJob* c_fence_y; CompactJob* c_split; CompactJob* c_side_x; CompactJob* c_side_plane;
    PlaneJob* p_road; PlaneJob* p_side;
    MeanJob* m_fence; SlabJob* s_road; KnnJob* k_road; FinalJob* fin;
    RansacJob* r_road; RansacJob* r_side;
};

struct WsPriv {
    FusedTables t;
    cudaStream_t side_stream; cudaEvent_t ev_fork, ev_join;
    bool single_stream;
    const int32_t* hyp_road; const int32_t* hyp_left; const int32_t* hyp_right; int n_hyp;
    SdFrameResult* results;
    SdCamera cam;
    double cell_scale;
    cudaEvent_t ev_t[3]; bool timing; int stage_mask;
    bool lazy_road, lazy_side;        // the current tables keep the MAD filters of the road / side chains as alive bytes
    cudaEvent_t ev_s[SD_NUM_STAGES + 1]; bool stages_valid;     // stage boundaries of the last timed call (profiling mode)
};

size_t carve_all(SdWorkspace* ws, Carver& c) {
    const int F = ws->max_frames;
    const size_t cap = (size_t)ws->cap;
    ws->fs = c.take<FrameState>(F);
    SdCloudBuf* bufs[8] = {&ws->road[0], &ws->road[1], &ws->fence[0], &ws->fence[1],
                           &ws->left[0], &ws->left[1], &ws->right[0], &ws->right[1]};
    for (SdCloudBuf* b : bufs) {
        b->x = c.take<float>(F * cap); b->y = c.take<float>(F * cap); b->z = c.take<float>(F * cap);
        b->src = c.take<int32_t>(F * cap);
    }
    ws->sel = c.take<SelState>((size_t)F * kChains);
    ws->cstatus = c.take<unsigned long long>((size_t)F * kChains * ws->max_tiles);
    ws->cctl = c.take<ScanCtl>((size_t)F * kChains);
    ws->mean_leaf = c.take<float>((size_t)F * (cap / 64 + 1));
    ws->partials = c.take<double>((size_t)F * kChains * kPlaneBlocks * kPlaneSums);
    ws->ptick = c.take<uint32_t>((size_t)F * kChains);
    ws->pflags = c.take<uint8_t>((size_t)F * ws->height * ws->width);
    ws->cflags = c.take<uint8_t>((size_t)F * 4 * cap);          // alive bytes of the unmaterialised filters: road, left, right, fence
    ws->ptcounts = c.take<int32_t>((size_t)F * ws->pix_tiles * 4);
    ws->ptoffs = c.take<int32_t>((size_t)F * ws->pix_tiles * 2);
    ws->lmask = c.take<double>(ws->width); ws->rmask = c.take<double>(ws->width);
    ws->gs = c.take<GridState>(F);
    ws->cell_count = c.take<int32_t>(F * ((size_t)ws->cell_cap + 8));
    ws->cell_start = c.take<int32_t>(F * ((size_t)ws->cell_cap + 8));
    ws->cell_of = c.take<int32_t>(F * cap);
    ws->sp = c.take<float4>(F * cap * kLevels);
    ws->cell_box = c.take<uint4>(F * ((size_t)ws->cell_cap / 8 + 64));
    ws->ybox = c.take<uint2>(F * ((size_t)ws->cell_cap / 8 + 64));
    ws->queue = c.take<int32_t>(F * cap); ws->queue_band = c.take<float>(F * cap);
    ws->avg = c.take<double>(F * cap);
    ws->cnt = c.take<int32_t>(F * cap);
    ws->gstatus = c.take<unsigned long long>((size_t)F * ws->grid_tiles);
    ws->gctl = c.take<ScanCtl>(F);
    const size_t mh = (size_t)(ws->max_hyp > 0 ? ws->max_hyp : 1);
    ws->hyp_coeff = c.take<double>((size_t)F * 3 * mh * 4);
    ws->hyp_counts = c.take<int32_t>((size_t)F * 3 * mh);
    ws->best_coeff = c.take<double>((size_t)F * 3 * 4);
    ws->jobs_bytes = (size_t)F * kJobsPerFrameBytes + 16 * 1024;
    ws->jobs = c.take<char>(ws->jobs_bytes);
    ws->scratch_bytes = kScratchBytes;
    ws->scratch = c.take<char>(kScratchBytes);
    return c.off;
}

void set_dims(SdWorkspace* ws, int max_frames, int height, int width, int max_hyp) {
    ws->max_frames = max_frames; ws->height = height; ws->width = width; ws->max_hyp = max_hyp;
    ws->cap = round_up(height * width, kCompactTile);
    ws->max_tiles = ws->cap / kCompactTile;
    ws->pix_tiles = (height * width + 1023) / 1024;
    ws->cell_cap = 2 * ws->cap + 4096;          // all grid levels together; the cell edge grows until they fit
    ws->grid_tiles = (ws->cell_cap + kScanTile - 1) / kScanTile + 1;
}

inline WsPriv* priv(SdWorkspace* ws) { return reinterpret_cast<WsPriv*>(ws->h_pinned ? (char*)ws->h_pinned + kPinnedBytes : nullptr); }

// bump allocator inside the device job arena, mirrored by a host staging vector
struct JobBuilder {
    SdWorkspace* ws; std::vector<char> host; size_t off;
    explicit JobBuilder(SdWorkspace* w) : ws(w), host(w->jobs_bytes, 0), off(0) {}
    template <typename T> T* alloc(size_t count, T** dev) {
        off = (off + 63) / 64 * 64;
        if (off + count * sizeof(T) > host.size()) return nullptr;
        T* h = reinterpret_cast<T*>(host.data() + off);
        *dev = reinterpret_cast<T*>(ws->jobs + off);
        off += count * sizeof(T);
        return h;
    }
};

SdCloudBuf frame_buf(const SdCloudBuf& b, int f, int cap) {
    SdCloudBuf r; const size_t o = (size_t)f * cap;
    r.x = b.x + o; r.y = b.y + o; r.z = b.z + o; r.src = b.src + o;
    return r;
}

PredDev make_pred(int kind, int axis) {
    PredDev p; memset(&p, 0, sizeof(p));
    p.kind = kind; p.axis = axis;
    return p;
}

void fill_compact(CompactJob& j, const SdCloudBuf& in, const int32_t* n_in, const SdCloudBuf& out, int32_t* n_out,
                  const PredDev& pred, SdWorkspace* ws, int f, int chain) {
    memset(&j, 0, sizeof(j));
    j.x = in.x; j.y = in.y; j.z = in.z; j.src = in.src; j.n_in = n_in;
    j.ox = out.x; j.oy = out.y; j.oz = out.z; j.osrc = out.src; j.n_out = n_out;
    j.pred = pred;
    j.status = ws->cstatus + ((size_t)f * kChains + chain) * ws->max_tiles;
    j.ctl = ws->cctl + (size_t)f * kChains + chain;
    j.max_tiles = ws->max_tiles;
}

void fill_sel(SelJob& j, const float* col, const int32_t* n, const float* center, float* out, SdWorkspace* ws, int f,
              int chain, uint32_t* status, uint32_t zero_bit) {
    memset(&j, 0, sizeof(j));
    j.col = col; j.n = n; j.center = center; j.out = out;
    j.st = ws->sel + (size_t)f * kChains + chain;
    j.status = status; j.zero_bit = zero_bit;
}

void fill_plane(PlaneJob& j, const SdCloudBuf& in, const int32_t* n, int axis, double* coeff, SdWorkspace* ws, int f,
                int chain, uint32_t* status, uint32_t empty_bit) {
    memset(&j, 0, sizeof(j));
    j.x = in.x; j.y = in.y; j.z = in.z; j.n = n; j.axis = axis;
    j.partials = ws->partials + ((size_t)f * kChains + chain) * kPlaneBlocks * kPlaneSums;
    j.ticket = ws->ptick + (size_t)f * kChains + chain;
    j.coeff = coeff; j.status = status; j.empty_bit = empty_bit;
}

void fill_knn(KnnJob& j, const float* x, const float* y, const float* z, const int32_t* n, SdWorkspace* ws, int f,
              double cell_scale) {
    memset(&j, 0, sizeof(j));
    const size_t o = (size_t)f * ws->cap, oc = (size_t)f * ((size_t)ws->cell_cap + 8);
    j.x = x; j.y = y; j.z = z; j.n = n;
    j.gs = ws->gs + f;
    j.cell_count = ws->cell_count + oc; j.cell_start = ws->cell_start + oc; j.cell_of = ws->cell_of + o;
    j.sp = ws->sp + o * kLevels;
    j.cell_box = ws->cell_box + (size_t)f * ((size_t)ws->cell_cap / 8 + 64);
    j.ybox = ws->ybox + (size_t)f * ((size_t)ws->cell_cap / 8 + 64);
    j.queue = ws->queue + o; j.queue_band = ws->queue_band + o;
    j.avg = ws->avg + o; j.cnt = ws->cnt + o;
    j.scan_status = ws->gstatus + (size_t)f * ws->grid_tiles; j.scan_ctl = ws->gctl + f;
    j.cell_cap = ws->cell_cap;
    j.cell_scale = cell_scale;
    j.count_cap = -1;
}

}  // namespace

extern "C" size_t sd_ws_bytes(int max_frames, int height, int width, int max_hypotheses) {
    if (max_frames < 1 || height < 1 || width < 4 || width % 4 != 0) return 0;
    SdWorkspace tmp; memset(&tmp, 0, sizeof(tmp));
    set_dims(&tmp, max_frames, height, width, max_hypotheses);
    Carver c{nullptr, 0, 0, true};
    return carve_all(&tmp, c) + kAlign;
}

extern "C" int sd_ws_create(SdWorkspace** out, void* d_mem, size_t bytes, int max_frames, int height, int width,
                            int max_hypotheses, void* stream) {
    if (!out || !d_mem) return fail(SD_ERR_INVALID, "sd_ws_create: null argument");
    if (width % 4 != 0) return fail(SD_ERR_INVALID, "sd_ws_create: width must be a multiple of 4");
    if (((uintptr_t)d_mem) % kAlign != 0) return fail(SD_ERR_INVALID, "sd_ws_create: memory must be 256-byte aligned");
    const size_t need = sd_ws_bytes(max_frames, height, width, max_hypotheses);
    if (need == 0 || bytes < need) return fail(SD_ERR_WORKSPACE, "sd_ws_create: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    SdWorkspace* ws = new (std::nothrow) SdWorkspace;
    if (!ws) return fail(SD_ERR_INVALID, "sd_ws_create: out of host memory");
    memset(ws, 0, sizeof(*ws));
    ws->base = (char*)d_mem; ws->bytes = bytes;
    set_dims(ws, max_frames, height, width, max_hypotheses);
    Carver c{ws->base, 0, bytes, false};
    carve_all(ws, c);
    // zero all control structures (clouds need no initialisation)
    SD_CUDA_TRY(cudaMemsetAsync(ws->fs, 0, sizeof(FrameState) * max_frames, st));
    SD_CUDA_TRY(cudaMemsetAsync(ws->sel, 0, sizeof(SelState) * max_frames * kChains, st));
    SD_CUDA_TRY(cudaMemsetAsync(ws->cstatus, 0, sizeof(unsigned long long) * max_frames * kChains * ws->max_tiles, st));
    SD_CUDA_TRY(cudaMemsetAsync(ws->cctl, 0, sizeof(ScanCtl) * max_frames * kChains, st));
    SD_CUDA_TRY(cudaMemsetAsync(ws->ptick, 0, sizeof(uint32_t) * max_frames * kChains, st));
    SD_CUDA_TRY(cudaMemsetAsync(ws->cell_count, 0, sizeof(int32_t) * max_frames * ((size_t)ws->cell_cap + 8), st));
    SD_CUDA_TRY(cudaMemsetAsync(ws->gstatus, 0, sizeof(unsigned long long) * max_frames * ws->grid_tiles, st));
    SD_CUDA_TRY(cudaMemsetAsync(ws->gctl, 0, sizeof(ScanCtl) * max_frames, st));
    SD_CUDA_TRY(cudaMemsetAsync(ws->scratch, 0, ws->scratch_bytes, st));
    { int rc_yb = sd_launch_ybox_init(ws->ybox, (size_t)max_frames * ((size_t)ws->cell_cap / 8 + 64), st); if (rc_yb) return rc_yb; }
    { int rc_box = sd_launch_cell_box_init(ws->cell_box, (size_t)max_frames * ((size_t)ws->cell_cap / 8 + 64), st); if (rc_box) return rc_box; }
    // GridState: bbox keys start at (+max, 0); FrameState slab keys likewise
    {
        std::vector<GridState> g(max_frames);
        memset(g.data(), 0, sizeof(GridState) * max_frames);
        for (auto& s : g) { for (int a = 0; a < 3; ++a) { s.bbox[a] = 0xffffffffu; s.bbox[3 + a] = 0u; } }
        SD_CUDA_TRY(cudaMemcpyAsync(ws->gs, g.data(), sizeof(GridState) * max_frames, cudaMemcpyHostToDevice, st));
        std::vector<FrameState> fsv(max_frames);
        memset(fsv.data(), 0, sizeof(FrameState) * max_frames);
        for (auto& s : fsv) { s.slab_keys[0] = 0xffffffffu; s.slab_keys[1] = 0u; s.ransac_best[0] = s.ransac_best[1] = s.ransac_best[2] = -1; }
        SD_CUDA_TRY(cudaMemcpyAsync(ws->fs, fsv.data(), sizeof(FrameState) * max_frames, cudaMemcpyHostToDevice, st));
        // blend ramps of DepthFrame.post_processing (semantic_depth.py:661-663), np.linspace semantics:
        // linspace(0,1,w)[i] = i * (1/(w-1)) for i < w-1, last element exactly 1
        std::vector<double> lm(width), rm(width);
        const double step = (width > 1) ? 1.0 / (double)(width - 1) : 0.0;
        for (int i = 0; i < width; ++i) {
            double l = (i == width - 1 && width > 1) ? 1.0 : (double)i * step;
            double v = 20.0 * (l - 0.05);
            v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
            lm[i] = 1.0 - v;
        }
        for (int i = 0; i < width; ++i) rm[i] = lm[width - 1 - i];
        SD_CUDA_TRY(cudaMemcpyAsync(ws->lmask, lm.data(), sizeof(double) * width, cudaMemcpyHostToDevice, st));
        SD_CUDA_TRY(cudaMemcpyAsync(ws->rmask, rm.data(), sizeof(double) * width, cudaMemcpyHostToDevice, st));
        SD_CUDA_TRY(cudaStreamSynchronize(st));
    }
    ws->h_pinned_bytes = kPinnedBytes + sizeof(WsPriv);
    SD_CUDA_TRY(cudaMallocHost(&ws->h_pinned, ws->h_pinned_bytes));
    memset(ws->h_pinned, 0, ws->h_pinned_bytes);
    WsPriv* pv = priv(ws);
    SD_CUDA_TRY(cudaStreamCreateWithFlags(&pv->side_stream, cudaStreamNonBlocking));
    SD_CUDA_TRY(cudaEventCreateWithFlags(&pv->ev_fork, cudaEventDisableTiming));
    SD_CUDA_TRY(cudaEventCreateWithFlags(&pv->ev_join, cudaEventDisableTiming));
    for (int i = 0; i < 3; ++i) SD_CUDA_TRY(cudaEventCreate(&pv->ev_t[i]));
    for (int i = 0; i <= SD_NUM_STAGES; ++i) SD_CUDA_TRY(cudaEventCreate(&pv->ev_s[i]));
    pv->timing = false; pv->stage_mask = 15; pv->stages_valid = false;
    const char* e = getenv("SD_FUSE_SINGLE_STREAM");
    pv->single_stream = (e && e[0] == '1');
    const char* cs = getenv("SD_KNN_CELL_SCALE");
    pv->cell_scale = cs ? atof(cs) : 0.8;
    if (!(pv->cell_scale > 0.0)) pv->cell_scale = 0.8;
    ws->fused_ready = false;
    *out = ws;
    return SD_OK;
}

extern "C" void sd_ws_destroy(SdWorkspace* ws) {
    if (!ws) return;
    if (ws->h_pinned) {
        WsPriv* pv = priv(ws);
        if (pv->side_stream) cudaStreamDestroy(pv->side_stream);
        if (pv->ev_fork) cudaEventDestroy(pv->ev_fork);
        if (pv->ev_join) cudaEventDestroy(pv->ev_join);
        for (int i = 0; i < 3; ++i) if (pv->ev_t[i]) cudaEventDestroy(pv->ev_t[i]);
        for (int i = 0; i <= SD_NUM_STAGES; ++i) if (pv->ev_s[i]) cudaEventDestroy(pv->ev_s[i]);
        cudaFreeHost(ws->h_pinned);
    }
    delete ws;
}

// ------------------------------------------------------------------------------------------------
// per-call ops
// ------------------------------------------------------------------------------------------------
namespace {

CallScratch* call_scratch(SdWorkspace* ws) { return reinterpret_cast<CallScratch*>(ws->scratch); }
char* call_jobs(SdWorkspace* ws) { return ws->scratch + ((sizeof(CallScratch) + 255) / 256 * 256); }

template <typename T>
int upload(T* dev, const T& host, cudaStream_t st) {
    SD_CUDA_TRY(cudaMemcpyAsync(dev, &host, sizeof(T), cudaMemcpyHostToDevice, st));
    return SD_OK;
}
template <typename T>
int download_sync(T* host, const T* dev, size_t count, SdWorkspace* ws, cudaStream_t st) {
    if (count * sizeof(T) > kPinnedBytes) return SD_ERR_INVALID;
    SD_CUDA_TRY(cudaMemcpyAsync(ws->h_pinned, dev, count * sizeof(T), cudaMemcpyDeviceToHost, st));
    SD_CUDA_TRY(cudaStreamSynchronize(st));
    memcpy(host, ws->h_pinned, count * sizeof(T));
    return SD_OK;
}

int check_n(SdWorkspace* ws, int n) {
    if (!ws) return fail(SD_ERR_WORKSPACE, "null workspace");
    if (n < 0) return fail(SD_ERR_INVALID, "negative point count");
    if (n > ws->cap) return fail(SD_ERR_WORKSPACE, "cloud larger than the workspace capacity (height*width)");
    return SD_OK;
}

}  // namespace

static int pixel_fuse_impl(const float* d_logits, const float* d_scores, const float* d_upw, const float* d_upb,
                           const float* d_disp, const double* d_lmask, const double* d_rmask,
                           int batch, int height, int width, const SdCamera* cam, double prob_thr, float road_z_to_meter, int flags,
                           float* d_road_x, float* d_road_y, float* d_road_z, int32_t* d_road_src,
                           float* d_fence_x, float* d_fence_y, float* d_fence_z, int32_t* d_fence_src,
                           int32_t* d_counts, uint8_t* d_labels, float* d_points, float* d_disp_pp, float* d_logits_out,
                           SdWorkspace* ws, void* stream) {
    if (!ws || (!d_logits && !d_scores) || !d_disp || !cam || !d_counts) return fail(SD_ERR_INVALID, "sd_pixel_fuse: null argument");
    if (height != ws->height || width != ws->width || batch > ws->max_frames || batch < 1)
        return fail(SD_ERR_WORKSPACE, "sd_pixel_fuse: shape does not match the workspace");
    if (d_scores && (height % 8 != 0 || width % 8 != 0 || !d_upw || !d_upb))
        return fail(SD_ERR_INVALID, "sd_pixel_fuse_scores: frame size must be a multiple of 8 and weights / bias must be given");
    if (!d_road_x || !d_road_y || !d_road_z || !d_road_src || !d_fence_x || !d_fence_y || !d_fence_z || !d_fence_src)
        return fail(SD_ERR_INVALID, "sd_pixel_fuse: null cloud output");
    SdCloudBuf road{d_road_x, d_road_y, d_road_z, d_road_src}, fence{d_fence_x, d_fence_y, d_fence_z, d_fence_src};
    return sd_launch_pixel(d_logits, d_disp, d_lmask ? d_lmask : ws->lmask, d_rmask ? d_rmask : ws->rmask,
                           batch, height, width, *cam, prob_thr, road_z_to_meter, flags & SD_PIX_RAW_DISPARITY, road, fence, height * width,
                           d_counts + 0, d_counts + 1, d_counts + 2, 3, d_labels, d_points, d_disp_pp,
                           ws->pflags, ws->ptcounts, ws->ptoffs, ws->pix_tiles, (cudaStream_t)stream,
                           d_scores, d_upw, d_upb, d_logits_out, (flags & SD_PIX_LABEL_ARGMAX) ? 1 : 0);
}

extern "C" int sd_pixel_fuse(const float* d_logits, const float* d_disp, const double* d_lmask, const double* d_rmask,
                             int batch, int height, int width, const SdCamera* cam, double prob_thr, float road_z_to_meter, int flags,
                             float* d_road_x, float* d_road_y, float* d_road_z, int32_t* d_road_src,
                             float* d_fence_x, float* d_fence_y, float* d_fence_z, int32_t* d_fence_src,
                             int32_t* d_counts, uint8_t* d_labels, float* d_points, float* d_disp_pp,
                             SdWorkspace* ws, void* stream) {
    if (!d_logits) return fail(SD_ERR_INVALID, "sd_pixel_fuse: null argument");
    return pixel_fuse_impl(d_logits, nullptr, nullptr, nullptr, d_disp, d_lmask, d_rmask, batch, height, width, cam, prob_thr,
                           road_z_to_meter, flags, d_road_x, d_road_y, d_road_z, d_road_src, d_fence_x, d_fence_y, d_fence_z,
                           d_fence_src, d_counts, d_labels, d_points, d_disp_pp, nullptr, ws, stream);
}

extern "C" int sd_pixel_fuse_scores(const float* d_scores, const float* d_up_weights, const float* d_up_bias,
                                    const float* d_disp, int batch, int height, int width, const SdCamera* cam,
                                    double prob_thr, float road_z_to_meter, int flags,
                                    float* d_road_x, float* d_road_y, float* d_road_z, int32_t* d_road_src,
                                    float* d_fence_x, float* d_fence_y, float* d_fence_z, int32_t* d_fence_src,
                                    int32_t* d_counts, uint8_t* d_labels, float* d_points, float* d_disp_pp, float* d_logits_out,
                                    SdWorkspace* ws, void* stream) {
    if (!d_scores) return fail(SD_ERR_INVALID, "sd_pixel_fuse_scores: null argument");
    return pixel_fuse_impl(nullptr, d_scores, d_up_weights, d_up_bias, d_disp, nullptr, nullptr, batch, height, width, cam, prob_thr,
                           road_z_to_meter, flags, d_road_x, d_road_y, d_road_z, d_road_src, d_fence_x, d_fence_y, d_fence_z,
                           d_fence_src, d_counts, d_labels, d_points, d_disp_pp, d_logits_out, ws, stream);
}

template <typename T, typename Launch>
static int ply_rows_impl(const T* d_x, const T* d_y, const T* d_z, const uint8_t* d_rgb, int n, char* d_out,
                         unsigned long long capacity, unsigned long long* h_nbytes, SdWorkspace* ws, void* stream, Launch launch) {
    int rc = check_n(ws, n); if (rc) return rc;
    if (!h_nbytes || (n > 0 && (!d_x || !d_y || !d_z || !d_rgb || !d_out))) return fail(SD_ERR_INVALID, "sd_ply_rows: null argument");
    *h_nbytes = 0ull;
    if (n == 0) return SD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CallScratch* cs = call_scratch(ws);
    // tile totals / offsets live in the neighbour search's cell_start array (2 * ceil(n / 256) words; free between calls)
    uint32_t* tiles = reinterpret_cast<uint32_t*>(ws->cell_start);
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(cs->d);
    rc = launch(d_x, d_y, d_z, d_rgb, n, d_out, capacity, tiles, d_total, st); if (rc) return rc;
    unsigned long long total = 0ull;
    rc = download_sync(&total, d_total, 1, ws, st); if (rc) return rc;
    *h_nbytes = total;
    if (total > capacity) return fail(SD_ERR_WORKSPACE, "sd_ply_rows: output buffer too small (h_nbytes holds the size needed)");
    return SD_OK;
}

extern "C" int sd_ply_rows(const float* d_x, const float* d_y, const float* d_z, const uint8_t* d_rgb, int n,
                           char* d_out, unsigned long long capacity, unsigned long long* h_nbytes, SdWorkspace* ws, void* stream) {
    return ply_rows_impl(d_x, d_y, d_z, d_rgb, n, d_out, capacity, h_nbytes, ws, stream, sd_launch_ply_rows);
}

extern "C" int sd_ply_rows_f64(const double* d_x, const double* d_y, const double* d_z, const uint8_t* d_rgb, int n,
                               char* d_out, unsigned long long capacity, unsigned long long* h_nbytes, SdWorkspace* ws, void* stream) {
    return ply_rows_impl(d_x, d_y, d_z, d_rgb, n, d_out, capacity, h_nbytes, ws, stream, sd_launch_ply_rows_f64);
}

extern "C" int sd_resize_cubic_u8(const uint8_t* d_src, int batch, int src_height, int src_width, int channels,
                                  uint8_t* d_dst, int dst_height, int dst_width, void* stream) {
    if (!d_src || !d_dst || batch < 1 || src_height < 1 || src_width < 1 || dst_height < 1 || dst_width < 1)
        return fail(SD_ERR_INVALID, "sd_resize_cubic_u8: bad argument");
    if (channels != 1 && channels != 3 && channels != 4) return fail(SD_ERR_UNSUPPORTED, "sd_resize_cubic_u8: channels must be 1, 3 or 4");
    return sd_launch_resize_cubic_u8(d_src, batch, src_height, src_width, channels, d_dst, dst_height, dst_width, (cudaStream_t)stream);
}

// scipy 1.2.1 bytescale of the int64 layer `mask x rgba` (values {0} U rgba when the mask is partial, rgba when it is
// full), evaluated like NumPy: int64 difference, fp64 scale, clip, + 0.5, truncation to uint8
static uint32_t overlay_layer(const int32_t* rgba, bool full) {
    long long cmin = full ? rgba[0] : 0, cmax = full ? rgba[0] : 0;
    for (int k = 0; k < 4; ++k) { cmin = std::min<long long>(cmin, rgba[k]); cmax = std::max<long long>(cmax, rgba[k]); }
    long long cscale = cmax - cmin;
    if (cscale == 0) cscale = 1;
    const double scale = 255.0 / (double)cscale;
    uint32_t packed = 0;
    for (int k = 0; k < 4; ++k) {
        double v = (double)((long long)rgba[k] - cmin) * scale + 0.0;
        v = std::min(std::max(v, 0.0), 255.0) + 0.5;
        packed |= ((uint32_t)(long long)v & 255u) << (8 * k);
    }
    return packed;
}

extern "C" int sd_overlay_masks(const uint8_t* d_frame, const uint8_t* d_labels, int batch, int height, int width,
                                const int32_t* road_rgba, const int32_t* fence_rgba, uint8_t* d_out, int32_t* d_scratch,
                                void* stream) {
    if (!d_frame || !d_labels || !d_out || !d_scratch || !road_rgba || !fence_rgba || batch < 1 || height < 1 || width < 1)
        return fail(SD_ERR_INVALID, "sd_overlay_masks: bad argument");
    if ((long long)height * width > 0x7fffffffll / 4) return fail(SD_ERR_INVALID, "sd_overlay_masks: frame too large");
    for (int k = 0; k < 4; ++k)
        if (road_rgba[k] < 0 || road_rgba[k] > 255 || fence_rgba[k] < 0 || fence_rgba[k] > 255)
            return fail(SD_ERR_INVALID, "sd_overlay_masks: colour components must be in [0, 255]");
    sd::OverlayLayers L;
    L.road_partial = overlay_layer(road_rgba, false);  L.road_full = overlay_layer(road_rgba, true);
    L.fence_partial = overlay_layer(fence_rgba, false); L.fence_full = overlay_layer(fence_rgba, true);
    return sd_launch_overlay(d_frame, d_labels, batch, height * width, L, d_scratch, d_out, (cudaStream_t)stream);
}

extern "C" int sd_draw_banner(uint8_t* d_frames, int batch, int height, int width, const int32_t* d_rects, int n_rects,
                              const uint32_t* d_glyph_bits, int n_bitmaps, int cell_height, int cell_words,
                              const int32_t* d_places, int n_places, void* stream) {
    if (!d_frames || batch < 1 || height < 1 || width < 1 || n_rects < 0 || n_places < 0)
        return fail(SD_ERR_INVALID, "sd_draw_banner: bad argument");
    if ((long long)height * width > 0x7fffffffll / 4) return fail(SD_ERR_INVALID, "sd_draw_banner: frame too large");
    if (n_rects > 65535) return fail(SD_ERR_INVALID, "sd_draw_banner: at most 65535 rectangles per call");
    if (n_rects > 0 && !d_rects) return fail(SD_ERR_INVALID, "sd_draw_banner: null rectangles");
    if (n_places > 0 && (!d_places || !d_glyph_bits || n_bitmaps < 1 || cell_height < 1 || cell_words < 1))
        return fail(SD_ERR_INVALID, "sd_draw_banner: glyph placements need an atlas");
    return sd_launch_banner(d_frames, batch, height, width, d_rects, n_rects, d_glyph_bits, n_bitmaps, cell_height, cell_words,
                            d_places, n_places, (cudaStream_t)stream);
}

extern "C" int sd_median_mad(const float* d_col, int n, float* h_out, SdWorkspace* ws, void* stream) {
    int rc = check_n(ws, n); if (rc) return rc;
    if (!d_col || !h_out) return fail(SD_ERR_INVALID, "sd_median_mad: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CallScratch* cs = call_scratch(ws);
    SelJob* dj = reinterpret_cast<SelJob*>(call_jobs(ws));
    SD_CUDA_TRY(cudaMemcpyAsync(&cs->n_in, &n, sizeof(int), cudaMemcpyHostToDevice, st));
    SelJob j[2]; memset(j, 0, sizeof(j));
    j[0].col = d_col; j[0].n = &cs->n_in; j[0].center = nullptr; j[0].st = &cs->sel; j[0].out = &cs->f[0];
    j[1] = j[0]; j[1].center = &cs->f[0]; j[1].out = &cs->f[1];
    SD_CUDA_TRY(cudaMemcpyAsync(dj, j, sizeof(j), cudaMemcpyHostToDevice, st));
    rc = sd_launch_select_median(dj, 1, n > 0 ? n : 1, st); if (rc) return rc;
    rc = sd_launch_select_median(dj + 1, 1, n > 0 ? n : 1, st); if (rc) return rc;
    return download_sync(h_out, cs->f, 2, ws, st);
}

extern "C" int sd_filter(const float* d_x, const float* d_y, const float* d_z, const int32_t* d_in_src, int n,
                         const SdPredicate* pred,
                         float* d_out_x, float* d_out_y, float* d_out_z, int32_t* d_out_src, int32_t* h_n_out,
                         SdWorkspace* ws, void* stream) {
    int rc = check_n(ws, n); if (rc) return rc;
    if (!pred || !h_n_out) return fail(SD_ERR_INVALID, "sd_filter: null argument");
    if (pred->axis < 0 || pred->axis > 2) return fail(SD_ERR_INVALID, "sd_filter: axis must be 0, 1 or 2");
    if (pred->kind < SD_PRED_LT || pred->kind > SD_PRED_ROR) return fail(SD_ERR_INVALID, "sd_filter: unknown predicate");
    cudaStream_t st = (cudaStream_t)stream;
    CallScratch* cs = call_scratch(ws);
    CompactJob* dj = reinterpret_cast<CompactJob*>(call_jobs(ws));
    SD_CUDA_TRY(cudaMemcpyAsync(&cs->n_in, &n, sizeof(int), cudaMemcpyHostToDevice, st));
    CompactJob j; memset(&j, 0, sizeof(j));
    j.x = d_x; j.y = d_y; j.z = d_z; j.src = d_in_src; j.n_in = &cs->n_in;
    j.ox = d_out_x; j.oy = d_out_y; j.oz = d_out_z; j.osrc = d_out_src; j.n_out = &cs->n_out;
    PredDev p = make_pred(pred->kind, pred->axis);
    p.ia = pred->ia; p.use_f32 = pred->use_f32; p.fa = pred->fa; p.f0 = pred->f0; p.f1 = pred->f1;
    p.da = pred->da; p.d0 = pred->d0; p.d1 = pred->d1; p.d2 = pred->d2; p.aux = pred->d_aux;
    j.pred = p;
    j.status = ws->cstatus; j.ctl = &cs->ctl; j.max_tiles = ws->max_tiles;
    SD_CUDA_TRY(cudaMemcpyAsync(dj, &j, sizeof(j), cudaMemcpyHostToDevice, st));
    rc = sd_launch_compact(dj, 1, n > 0 ? n : 1, st); if (rc) return rc;
    return download_sync(h_n_out, &cs->n_out, 1, ws, st);
}

extern "C" int sd_plane_fit(const float* d_x, const float* d_y, const float* d_z, int n, int axis,
                            double* h_coeff, int32_t* h_singular, SdWorkspace* ws, void* stream) {
    int rc = check_n(ws, n); if (rc) return rc;
    if (axis < 0 || axis > 2 || !h_coeff) return fail(SD_ERR_INVALID, "sd_plane_fit: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CallScratch* cs = call_scratch(ws);
    PlaneJob* dj = reinterpret_cast<PlaneJob*>(call_jobs(ws));
    SD_CUDA_TRY(cudaMemcpyAsync(&cs->n_in, &n, sizeof(int), cudaMemcpyHostToDevice, st));
    SD_CUDA_TRY(cudaMemsetAsync(&cs->status, 0, sizeof(uint32_t), st));
    PlaneJob j; memset(&j, 0, sizeof(j));
    j.x = d_x; j.y = d_y; j.z = d_z; j.n = &cs->n_in; j.axis = axis;
    j.partials = cs->partials; j.ticket = &cs->ptick; j.coeff = cs->d; j.status = &cs->status; j.empty_bit = SD_ST_EMPTY_ROAD;
    SD_CUDA_TRY(cudaMemcpyAsync(dj, &j, sizeof(j), cudaMemcpyHostToDevice, st));
    rc = sd_launch_plane(dj, 1, n > 0 ? n : 1, st); if (rc) return rc;
    rc = download_sync(h_coeff, cs->d, 3, ws, st); if (rc) return rc;
    uint32_t status = 0;
    rc = download_sync(&status, &cs->status, 1, ws, st); if (rc) return rc;
    if (h_singular) *h_singular = (status & (SD_ST_SINGULAR_FIT | SD_ST_EMPTY_ROAD)) ? 1 : 0;
    return SD_OK;
}

extern "C" int sd_mean_f32(const float* d_col, int n, float* h_mean, SdWorkspace* ws, void* stream) {
    int rc = check_n(ws, n); if (rc) return rc;
    if (!h_mean) return fail(SD_ERR_INVALID, "sd_mean_f32: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CallScratch* cs = call_scratch(ws);
    MeanJob* dj = reinterpret_cast<MeanJob*>(call_jobs(ws));
    SD_CUDA_TRY(cudaMemcpyAsync(&cs->n_in, &n, sizeof(int), cudaMemcpyHostToDevice, st));
    MeanJob j{d_col, &cs->n_in, &cs->f[2], ws->mean_leaf};
    SD_CUDA_TRY(cudaMemcpyAsync(dj, &j, sizeof(j), cudaMemcpyHostToDevice, st));
    rc = sd_launch_mean(dj, 1, ws->cap, st); if (rc) return rc;
    return download_sync(h_mean, &cs->f[2], 1, ws, st);
}

extern "C" int sd_slab_minmax(const float* d_x, const float* d_z, int n, double lo, double hi, int use_f32,
                              float* h_xmin, float* h_xmax, int32_t* h_count, SdWorkspace* ws, void* stream) {
    int rc = check_n(ws, n); if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CallScratch* cs = call_scratch(ws);
    SlabJob* dj = reinterpret_cast<SlabJob*>(call_jobs(ws));
    uint32_t init[2] = {0xffffffffu, 0u}; int zero = 0;
    SD_CUDA_TRY(cudaMemcpyAsync(&cs->n_in, &n, sizeof(int), cudaMemcpyHostToDevice, st));
    SD_CUDA_TRY(cudaMemcpyAsync(cs->keys, init, sizeof(init), cudaMemcpyHostToDevice, st));
    SD_CUDA_TRY(cudaMemcpyAsync(&cs->count, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
    SlabJob j; memset(&j, 0, sizeof(j));
    j.x = d_x; j.z = d_z; j.n = &cs->n_in; j.lo = lo; j.hi = hi; j.lo32 = (float)lo; j.hi32 = (float)hi; j.use_f32 = use_f32;
    j.keys = cs->keys; j.count = &cs->count;
    SD_CUDA_TRY(cudaMemcpyAsync(dj, &j, sizeof(j), cudaMemcpyHostToDevice, st));
    rc = sd_launch_slab(dj, 1, n > 0 ? n : 1, st); if (rc) return rc;
    uint32_t keys[2]; int cnt = 0;
    rc = download_sync(keys, cs->keys, 2, ws, st); if (rc) return rc;
    rc = download_sync(&cnt, &cs->count, 1, ws, st); if (rc) return rc;
    if (h_count) *h_count = cnt;
    if (h_xmin) *h_xmin = key2f(keys[0]);
    if (h_xmax) *h_xmax = key2f(keys[1]);
    return SD_OK;
}

static int knn_common(const float* d_x, const float* d_y, const float* d_z, int n, SdWorkspace* ws, KnnJob** dj_out,
                      KnnJob* hj, cudaStream_t st) {
    CallScratch* cs = call_scratch(ws);
    KnnJob* dj = reinterpret_cast<KnnJob*>(call_jobs(ws));
    SD_CUDA_TRY(cudaMemcpyAsync(&cs->n_in, &n, sizeof(int), cudaMemcpyHostToDevice, st));
    fill_knn(*hj, d_x, d_y, d_z, &cs->n_in, ws, 0, priv(ws)->cell_scale);
    hj->stats = cs->d; hj->n_alive = nullptr;
    *dj_out = dj;
    return SD_OK;
}

extern "C" int sd_knn_mean_distance(const float* d_x, const float* d_y, const float* d_z, int n, int k, double std_ratio,
                                    double* d_avg, double* h_stats, SdWorkspace* ws, void* stream) {
    int rc = check_n(ws, n); if (rc) return rc;
    if (k < 1 || k > kMaxKnnK) return fail(SD_ERR_INVALID, "sd_knn_mean_distance: k must be in [1, 64]");
    if (!d_avg) return fail(SD_ERR_INVALID, "sd_knn_mean_distance: null output");
    cudaStream_t st = (cudaStream_t)stream;
    KnnJob j, *dj;
    rc = knn_common(d_x, d_y, d_z, n, ws, &dj, &j, st); if (rc) return rc;
    j.k = k; j.std_ratio = std_ratio; j.avg = d_avg;
    SD_CUDA_TRY(cudaMemcpyAsync(dj, &j, sizeof(j), cudaMemcpyHostToDevice, st));
    rc = sd_launch_grid_build(dj, 1, n > 0 ? n : 1, st); if (rc) return rc;
    rc = sd_launch_knn(dj, 1, n > 0 ? n : 1, k, st); if (rc) return rc;
    double stats[3];
    rc = download_sync(stats, call_scratch(ws)->d, 3, ws, st); if (rc) return rc;
    if (h_stats) memcpy(h_stats, stats, sizeof(stats));
    return SD_OK;
}

extern "C" int sd_radius_count(const float* d_x, const float* d_y, const float* d_z, int n, double radius, int cap,
                               int32_t* d_counts, SdWorkspace* ws, void* stream) {
    int rc = check_n(ws, n); if (rc) return rc;
    if (!(radius > 0.0) || !d_counts) return fail(SD_ERR_INVALID, "sd_radius_count: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    KnnJob j, *dj;
    rc = knn_common(d_x, d_y, d_z, n, ws, &dj, &j, st); if (rc) return rc;
    j.radius = radius; j.use_sor = 0; j.count_cap = cap; j.cnt = d_counts;
    SD_CUDA_TRY(cudaMemcpyAsync(dj, &j, sizeof(j), cudaMemcpyHostToDevice, st));
    rc = sd_launch_grid_build(dj, 1, n > 0 ? n : 1, st); if (rc) return rc;
    rc = sd_launch_radius(dj, 1, n > 0 ? n : 1, st); if (rc) return rc;
    SD_CUDA_TRY(cudaStreamSynchronize(st));
    return SD_OK;
}

extern "C" int sd_ransac_score(const float* d_x, const float* d_y, const float* d_z, int n, int axis, double threshold,
                               const int32_t* d_triplets, int n_hyp, int32_t* d_hyp_counts, int32_t* h_best,
                               double* h_best_coeff, SdWorkspace* ws, void* stream) {
    int rc = check_n(ws, n); if (rc) return rc;
    if (axis < 0 || axis > 2 || !d_triplets || !d_hyp_counts || n_hyp < 1) return fail(SD_ERR_INVALID, "sd_ransac_score: bad argument");
    if (n_hyp > ws->max_hyp) return fail(SD_ERR_WORKSPACE, "sd_ransac_score: more hypotheses than the workspace reserves");
    cudaStream_t st = (cudaStream_t)stream;
    CallScratch* cs = call_scratch(ws);
    RansacJob* dj = reinterpret_cast<RansacJob*>(call_jobs(ws));
    SD_CUDA_TRY(cudaMemcpyAsync(&cs->n_in, &n, sizeof(int), cudaMemcpyHostToDevice, st));
    RansacJob j; memset(&j, 0, sizeof(j));
    j.x = d_x; j.y = d_y; j.z = d_z; j.n = &cs->n_in; j.triplets = d_triplets;
    j.hyp_coeff = ws->hyp_coeff; j.hyp_counts = d_hyp_counts; j.best = &cs->best; j.best_coeff = cs->d + 4;
    j.axis = axis; j.n_hyp = n_hyp; j.thr = threshold;
    SD_CUDA_TRY(cudaMemcpyAsync(dj, &j, sizeof(j), cudaMemcpyHostToDevice, st));
    rc = sd_launch_ransac(dj, 1, n > 0 ? n : 1, n_hyp, st); if (rc) return rc;
    int best = -1; double bc[3];
    rc = download_sync(&best, &cs->best, 1, ws, st); if (rc) return rc;
    rc = download_sync(bc, cs->d + 4, 3, ws, st); if (rc) return rc;
    if (h_best) *h_best = best;
    if (h_best_coeff) memcpy(h_best_coeff, bc, sizeof(bc));
    return SD_OK;
}

// ------------------------------------------------------------------------------------------------
// fused path
// ------------------------------------------------------------------------------------------------
namespace {

bool same_params(const SdParams& a, const SdParams& b) { return memcmp(&a, &b, sizeof(SdParams)) == 0; }

int build_fused_tables(SdWorkspace* ws, int B, const SdParams& P, const SdCamera& cam, const int32_t* hyp_road, const int32_t* hyp_left,
                       const int32_t* hyp_right, int n_hyp, SdFrameResult* d_results, cudaStream_t st) {
    WsPriv* pv = priv(ws);
    pv->cam = cam;
    JobBuilder jb(ws);
    FusedTables& T = pv->t;
    memset(&T, 0, sizeof(T));
    const int cap = ws->cap;
    const bool lazy_road = (hyp_road == nullptr), lazy_side = (hyp_left == nullptr && hyp_right == nullptr);
    pv->lazy_road = lazy_road; pv->lazy_side = lazy_side;
#define SD_ALLOC(field, type, count) type* h_##field = jb.alloc<type>(count, &T.field); if (!h_##field) return fail(SD_ERR_WORKSPACE, "job arena too small")
    SD_ALLOC(sel_road_y_med, SelJob, B); SD_ALLOC(sel_road_y_mad, SelJob, B);
    SD_ALLOC(sel_road_x_med, SelJob, B); SD_ALLOC(sel_road_x_mad, SelJob, B);
    SD_ALLOC(sel_fence_y_med, SelJob, B); SD_ALLOC(sel_fence_y_mad, SelJob, B);
    SD_ALLOC(sel_side_x_med, SelJob, 2 * B); SD_ALLOC(sel_side_x_mad, SelJob, 2 * B);
    SD_ALLOC(c_road_y, CompactJob, B); SD_ALLOC(c_road_x, CompactJob, B);
    SD_ALLOC(c_road_plane, CompactJob, B); SD_ALLOC(c_road_final, CompactJob, B);
    SD_ALLOC(c_fence_y, CompactJob, B);
    SD_ALLOC(c_split, CompactJob, 2 * B); SD_ALLOC(c_side_x, CompactJob, 2 * B); SD_ALLOC(c_side_plane, CompactJob, 2 * B);
    SD_ALLOC(p_road, PlaneJob, B); SD_ALLOC(p_side, PlaneJob, 2 * B);
    SD_ALLOC(m_fence, MeanJob, B); SD_ALLOC(s_road, SlabJob, B); SD_ALLOC(k_road, KnnJob, B); SD_ALLOC(fin, FinalJob, B);
    SD_ALLOC(r_road, RansacJob, B); SD_ALLOC(r_side, RansacJob, 2 * B);
#undef SD_ALLOC
    const size_t mh = (size_t)(ws->max_hyp > 0 ? ws->max_hyp : 1);
    for (int f = 0; f < B; ++f) {
        FrameState* fs = ws->fs + f;
        SdCloudBuf rA = frame_buf(ws->road[0], f, cap), rB = frame_buf(ws->road[1], f, cap);
        SdCloudBuf fA = frame_buf(ws->fence[0], f, cap), fB = frame_buf(ws->fence[1], f, cap);
        SdCloudBuf lA = frame_buf(ws->left[0], f, cap), lB = frame_buf(ws->left[1], f, cap);
        SdCloudBuf gA = frame_buf(ws->right[0], f, cap), gB = frame_buf(ws->right[1], f, cap);
        // ---------------- road chain ----------------
        // Without RANSAC hypotheses (the reference's behaviour) the three filters that precede the Open3D pair are NOT
        // materialised one by one: MAD y is evaluated by the first pass of the x median (alive bytes + count), MAD x by the
        // plane-moment kernel, and ONE compaction applies (alive && plane residual).  Three compactions of a cloud that keeps
        // > 99 % of its points become one; counts and kept indices are what three separate calls give.  With hypotheses the
        // triplets index rows of the filtered cloud, so that cloud has to exist: the classic layout below.
        uint8_t* rflag = ws->cflags + ((size_t)f * 4 + 0) * cap;
        uint8_t* lflag = ws->cflags + ((size_t)f * 4 + 1) * cap;
        uint8_t* gflag = ws->cflags + ((size_t)f * 4 + 2) * cap;
        uint8_t* fflag = ws->cflags + ((size_t)f * 4 + 3) * cap;
        // MAD y (thr 15) on rA (after the z cut)                           semantic_depth.py:209
        fill_sel(h_sel_road_y_med[f], rA.y, &fs->n[SD_CNT_ROAD_Z], nullptr, &fs->med[0], ws, f, 0, nullptr, 0);
        fill_sel(h_sel_road_y_mad[f], rA.y, &fs->n[SD_CNT_ROAD_Z], &fs->med[0], &fs->mad[0], ws, f, 0, &fs->status, SD_ST_MAD_ZERO);
        if (lazy_road) {
            // MAD x (thr 2) on the survivors of MAD y                      :212
            fill_sel(h_sel_road_x_med[f], rA.x, &fs->n[SD_CNT_ROAD_MAD_Y], nullptr, &fs->med[1], ws, f, 0, nullptr, 0);
            h_sel_road_x_med[f].n_loop = &fs->n[SD_CNT_ROAD_Z]; h_sel_road_x_med[f].flag = rflag;
            h_sel_road_x_med[f].mark = MadMark{rA.y, &fs->med[0], &fs->mad[0], P.road_mad_y_thr, 0};
            h_sel_road_x_med[f].n_mark_out = &fs->n[SD_CNT_ROAD_MAD_Y]; h_sel_road_x_med[f].first_alive_out = &fs->road_first_alive;
            fill_sel(h_sel_road_x_mad[f], rA.x, &fs->n[SD_CNT_ROAD_MAD_Y], &fs->med[1], &fs->mad[1], ws, f, 0, &fs->status, SD_ST_MAD_ZERO);
            h_sel_road_x_mad[f].n_loop = &fs->n[SD_CNT_ROAD_Z]; h_sel_road_x_mad[f].flag = rflag;
            // plane (axis 1, thr 5) on the survivors of MAD x: rA -> rB    :215-219
            fill_plane(h_p_road[f], rA, &fs->n[SD_CNT_ROAD_MAD_X], 1, fs->coeff[0], ws, f, 0, &fs->status, SD_ST_EMPTY_ROAD);
            h_p_road[f].n_loop = &fs->n[SD_CNT_ROAD_Z]; h_p_road[f].flag = rflag; h_p_road[f].flag_out = rflag;
            h_p_road[f].mark = MadMark{rA.x, &fs->med[1], &fs->mad[1], P.road_mad_x_thr, 0};
            h_p_road[f].n_mark_out = &fs->n[SD_CNT_ROAD_MAD_X]; h_p_road[f].shift_row = &fs->road_first_alive;
            { PredDev p = make_pred(SD_PRED_PLANE, 1); p.da = P.road_plane_thr; p.p_d = fs->coeff[0];
              fill_compact(h_c_road_plane[f], rA, &fs->n[SD_CNT_ROAD_Z], rB, &fs->n[SD_CNT_ROAD_PLANE], p, ws, f, 0);
              h_c_road_plane[f].flag = rflag; }
        } else {
        { PredDev p = make_pred(SD_PRED_MAD, 1); p.fa = P.road_mad_y_thr; p.p_f0 = &fs->med[0]; p.p_f1 = &fs->mad[0];
          fill_compact(h_c_road_y[f], rA, &fs->n[SD_CNT_ROAD_Z], rB, &fs->n[SD_CNT_ROAD_MAD_Y], p, ws, f, 0); }
        // MAD x (thr 2): rB -> rA                                          :212
        fill_sel(h_sel_road_x_med[f], rB.x, &fs->n[SD_CNT_ROAD_MAD_Y], nullptr, &fs->med[1], ws, f, 0, nullptr, 0);
        fill_sel(h_sel_road_x_mad[f], rB.x, &fs->n[SD_CNT_ROAD_MAD_Y], &fs->med[1], &fs->mad[1], ws, f, 0, &fs->status, SD_ST_MAD_ZERO);
        { PredDev p = make_pred(SD_PRED_MAD, 0); p.fa = P.road_mad_x_thr; p.p_f0 = &fs->med[1]; p.p_f1 = &fs->mad[1];
          fill_compact(h_c_road_x[f], rB, &fs->n[SD_CNT_ROAD_MAD_Y], rA, &fs->n[SD_CNT_ROAD_MAD_X], p, ws, f, 0); }
        // plane (axis 1, thr 5): rA -> rB                                  :215-219
        fill_plane(h_p_road[f], rA, &fs->n[SD_CNT_ROAD_MAD_X], 1, fs->coeff[0], ws, f, 0, &fs->status, SD_ST_EMPTY_ROAD);
        { PredDev p = make_pred(SD_PRED_PLANE, 1); p.da = P.road_plane_thr; p.p_d = fs->coeff[0];
          fill_compact(h_c_road_plane[f], rA, &fs->n[SD_CNT_ROAD_MAD_X], rB, &fs->n[SD_CNT_ROAD_PLANE], p, ws, f, 0); }
        }
        if (hyp_road) {
            RansacJob& r = h_r_road[f]; memset(&r, 0, sizeof(r));
            r.x = rA.x; r.y = rA.y; r.z = rA.z; r.n = &fs->n[SD_CNT_ROAD_MAD_X];
            r.triplets = hyp_road + (size_t)f * n_hyp * 3;
            r.hyp_coeff = ws->hyp_coeff + ((size_t)f * 3 + 0) * mh * 4; r.hyp_counts = ws->hyp_counts + ((size_t)f * 3 + 0) * mh;
            r.best = &fs->ransac_best[0]; r.best_coeff = ws->best_coeff + ((size_t)f * 3 + 0) * 4;
            r.axis = 1; r.n_hyp = n_hyp; r.thr = P.road_plane_thr;
            h_p_road[f].use_inliers = 1; h_p_road[f].hyp = r.best_coeff; h_p_road[f].thr = P.road_plane_thr;
        }
        // SOR + ROR on rB -> rA                                            :227-245
        fill_knn(h_k_road[f], rB.x, rB.y, rB.z, &fs->n[SD_CNT_ROAD_PLANE], ws, f, pv->cell_scale);
        h_k_road[f].k = P.sor_nb_neighbors; h_k_road[f].std_ratio = P.sor_std_ratio; h_k_road[f].radius = P.ror_radius;
        h_k_road[f].nb_points = P.ror_nb_points; h_k_road[f].use_sor = P.use_sor; h_k_road[f].count_cap = P.ror_nb_points;
        h_k_road[f].stats = fs->sor_stats; h_k_road[f].n_alive = &fs->n_sor_alive;
        { PredDev p = make_pred(SD_PRED_SORROR, 0); p.ia = P.ror_nb_points;
          if (P.use_sor) { p.aux = h_k_road[f].avg; p.p_d = fs->sor_stats + 2; }
          if (P.use_ror) { p.aux2 = h_k_road[f].cnt; }
          fill_compact(h_c_road_final[f], rB, &fs->n[SD_CNT_ROAD_PLANE], rA, &fs->n[SD_CNT_ROAD_ROR], p, ws, f, 0); }
        // slab min/max on rA                                               :254-259
        { SlabJob& s = h_s_road[f]; memset(&s, 0, sizeof(s));
          s.x = rA.x; s.z = rA.z; s.n = &fs->n[SD_CNT_ROAD_ROR]; s.lo = P.slab_lo; s.hi = P.slab_hi;
          s.lo32 = (float)P.slab_lo; s.hi32 = (float)P.slab_hi; s.use_f32 = 0;   // the cloud is fp64 after Open3D (:244)
          s.keys = fs->slab_keys; s.count = &fs->slab_count; }
        // ---------------- fence chain ----------------
        // MAD y (thr 5) and |z| < 35 in ONE compaction, fA -> fB: the second call sees the first one's survivors and the
        // first call's count stays observable                              :279, :283-284
        fill_sel(h_sel_fence_y_med[f], fA.y, &fs->n[SD_CNT_FENCE_GATHER], nullptr, &fs->med[2], ws, f, 1, nullptr, 0);
        fill_sel(h_sel_fence_y_mad[f], fA.y, &fs->n[SD_CNT_FENCE_GATHER], &fs->med[2], &fs->mad[2], ws, f, 1, &fs->status, SD_ST_MAD_ZERO);
        { PredDev p = make_pred(SD_PRED_MAD, 1); p.fa = P.fence_mad_y_thr; p.p_f0 = &fs->med[2]; p.p_f1 = &fs->mad[2];
          fill_compact(h_c_fence_y[f], fA, &fs->n[SD_CNT_FENCE_GATHER], fB, &fs->n[SD_CNT_FENCE_ABS_Z], p, ws, f, 1);
          PredDev q = make_pred(SD_PRED_ABS_LT, 2); q.fa = P.fence_abs_z_thr;
          h_c_fence_y[f].pred2 = q; h_c_fence_y[f].has_pred2 = 1; h_c_fence_y[f].n_mid = &fs->n[SD_CNT_FENCE_MAD_Y];
          h_c_fence_y[f].mid_alive = fflag; }
        // mean x and the split: fB -> lA, gA                               :286-287
        h_m_fence[f] = MeanJob{fB.x, &fs->n[SD_CNT_FENCE_ABS_Z], &fs->fence_mean, ws->mean_leaf + (size_t)f * (cap / 64 + 1)};
        { PredDev p = make_pred(SD_PRED_LT, 0); p.p_f0 = &fs->fence_mean;
          fill_compact(h_c_split[2 * f], fB, &fs->n[SD_CNT_FENCE_ABS_Z], lA, &fs->n[SD_CNT_LEFT_SPLIT], p, ws, f, 2); }
        { PredDev p = make_pred(SD_PRED_GT, 0); p.p_f0 = &fs->fence_mean;
          fill_compact(h_c_split[2 * f + 1], fB, &fs->n[SD_CNT_FENCE_ABS_Z], gA, &fs->n[SD_CNT_RIGHT_SPLIT], p, ws, f, 3); }
        // side MAD x (thr 5 / 1) on lA / gA                                :291, :302
        fill_sel(h_sel_side_x_med[2 * f], lA.x, &fs->n[SD_CNT_LEFT_SPLIT], nullptr, &fs->med[3], ws, f, 2, nullptr, 0);
        fill_sel(h_sel_side_x_mad[2 * f], lA.x, &fs->n[SD_CNT_LEFT_SPLIT], &fs->med[3], &fs->mad[3], ws, f, 2, &fs->status, SD_ST_MAD_ZERO);
        fill_sel(h_sel_side_x_med[2 * f + 1], gA.x, &fs->n[SD_CNT_RIGHT_SPLIT], nullptr, &fs->med[4], ws, f, 3, nullptr, 0);
        fill_sel(h_sel_side_x_mad[2 * f + 1], gA.x, &fs->n[SD_CNT_RIGHT_SPLIT], &fs->med[4], &fs->mad[4], ws, f, 3, &fs->status, SD_ST_MAD_ZERO);
        if (lazy_side) {
            // the MAD filter is evaluated by the plane-moment kernel (alive bytes + count), one compaction applies
            // (alive && plane residual): lA -> lB, gA -> gB                :294-298, :305-309
            fill_plane(h_p_side[2 * f], lA, &fs->n[SD_CNT_LEFT_MAD_X], 0, fs->coeff[1], ws, f, 2, &fs->status, SD_ST_EMPTY_FENCE_LEFT);
            h_p_side[2 * f].n_loop = &fs->n[SD_CNT_LEFT_SPLIT]; h_p_side[2 * f].flag_out = lflag;
            h_p_side[2 * f].mark = MadMark{lA.x, &fs->med[3], &fs->mad[3], P.left_mad_x_thr, 0}; h_p_side[2 * f].n_mark_out = &fs->n[SD_CNT_LEFT_MAD_X];
            fill_plane(h_p_side[2 * f + 1], gA, &fs->n[SD_CNT_RIGHT_MAD_X], 0, fs->coeff[2], ws, f, 3, &fs->status, SD_ST_EMPTY_FENCE_RIGHT);
            h_p_side[2 * f + 1].n_loop = &fs->n[SD_CNT_RIGHT_SPLIT]; h_p_side[2 * f + 1].flag_out = gflag;
            h_p_side[2 * f + 1].mark = MadMark{gA.x, &fs->med[4], &fs->mad[4], P.right_mad_x_thr, 0}; h_p_side[2 * f + 1].n_mark_out = &fs->n[SD_CNT_RIGHT_MAD_X];
            { PredDev p = make_pred(SD_PRED_PLANE, 0); p.da = P.fence_plane_thr; p.p_d = fs->coeff[1];
              fill_compact(h_c_side_plane[2 * f], lA, &fs->n[SD_CNT_LEFT_SPLIT], lB, &fs->n[SD_CNT_LEFT_PLANE], p, ws, f, 2);
              h_c_side_plane[2 * f].flag = lflag; }
            { PredDev p = make_pred(SD_PRED_PLANE, 0); p.da = P.fence_plane_thr; p.p_d = fs->coeff[2];
              fill_compact(h_c_side_plane[2 * f + 1], gA, &fs->n[SD_CNT_RIGHT_SPLIT], gB, &fs->n[SD_CNT_RIGHT_PLANE], p, ws, f, 3);
              h_c_side_plane[2 * f + 1].flag = gflag; }
        } else {
        { PredDev p = make_pred(SD_PRED_MAD, 0); p.fa = P.left_mad_x_thr; p.p_f0 = &fs->med[3]; p.p_f1 = &fs->mad[3];
          fill_compact(h_c_side_x[2 * f], lA, &fs->n[SD_CNT_LEFT_SPLIT], lB, &fs->n[SD_CNT_LEFT_MAD_X], p, ws, f, 2); }
        { PredDev p = make_pred(SD_PRED_MAD, 0); p.fa = P.right_mad_x_thr; p.p_f0 = &fs->med[4]; p.p_f1 = &fs->mad[4];
          fill_compact(h_c_side_x[2 * f + 1], gA, &fs->n[SD_CNT_RIGHT_SPLIT], gB, &fs->n[SD_CNT_RIGHT_MAD_X], p, ws, f, 3); }
        // side planes (axis 0, thr 1): lB -> lA, gB -> gA                  :294-298, :305-309
        fill_plane(h_p_side[2 * f], lB, &fs->n[SD_CNT_LEFT_MAD_X], 0, fs->coeff[1], ws, f, 2, &fs->status, SD_ST_EMPTY_FENCE_LEFT);
        fill_plane(h_p_side[2 * f + 1], gB, &fs->n[SD_CNT_RIGHT_MAD_X], 0, fs->coeff[2], ws, f, 3, &fs->status, SD_ST_EMPTY_FENCE_RIGHT);
        { PredDev p = make_pred(SD_PRED_PLANE, 0); p.da = P.fence_plane_thr; p.p_d = fs->coeff[1];
          fill_compact(h_c_side_plane[2 * f], lB, &fs->n[SD_CNT_LEFT_MAD_X], lA, &fs->n[SD_CNT_LEFT_PLANE], p, ws, f, 2); }
        { PredDev p = make_pred(SD_PRED_PLANE, 0); p.da = P.fence_plane_thr; p.p_d = fs->coeff[2];
          fill_compact(h_c_side_plane[2 * f + 1], gB, &fs->n[SD_CNT_RIGHT_MAD_X], gA, &fs->n[SD_CNT_RIGHT_PLANE], p, ws, f, 3); }
        }
        for (int s = 0; s < 2; ++s) {
            const int32_t* hyp = s == 0 ? hyp_left : hyp_right;
            RansacJob& r = h_r_side[2 * f + s]; memset(&r, 0, sizeof(r));
            if (!hyp) continue;
            const SdCloudBuf& in = s == 0 ? lB : gB;
            r.x = in.x; r.y = in.y; r.z = in.z; r.n = &fs->n[s == 0 ? SD_CNT_LEFT_MAD_X : SD_CNT_RIGHT_MAD_X];
            r.triplets = hyp + (size_t)f * n_hyp * 3;
            r.hyp_coeff = ws->hyp_coeff + ((size_t)f * 3 + 1 + s) * mh * 4; r.hyp_counts = ws->hyp_counts + ((size_t)f * 3 + 1 + s) * mh;
            r.best = &fs->ransac_best[1 + s]; r.best_coeff = ws->best_coeff + ((size_t)f * 3 + 1 + s) * 4;
            r.axis = 0; r.n_hyp = n_hyp; r.thr = P.fence_plane_thr;
            h_p_side[2 * f + s].use_inliers = 1; h_p_side[2 * f + s].hyp = r.best_coeff; h_p_side[2 * f + s].thr = P.fence_plane_thr;
        }
        h_fin[f] = FinalJob{fs, d_results + f};
    }
    SD_CUDA_TRY(cudaMemcpyAsync(ws->jobs, jb.host.data(), jb.off, cudaMemcpyHostToDevice, st));
    SD_CUDA_TRY(cudaStreamSynchronize(st));
    ws->fused_ready = true; ws->fused_batch = B; ws->fused_params = P;
    pv->hyp_road = hyp_road; pv->hyp_left = hyp_left; pv->hyp_right = hyp_right; pv->n_hyp = n_hyp; pv->results = d_results;
    return SD_OK;
}

}  // namespace

static int fuse_impl(const float* d_logits, const float* d_scores, const float* d_upw, const float* d_upb,
                     const float* d_disp, int batch, int height, int width,
                     const SdCamera* cam, const SdParams* params,
                     const int32_t* d_hyp_road, const int32_t* d_hyp_left, const int32_t* d_hyp_right, int n_hyp,
                     SdFrameResult* d_results, SdWorkspace* ws, void* stream) {
    if (!ws || (!d_logits && !d_scores) || !d_disp || !cam || !params || !d_results) return fail(SD_ERR_INVALID, "sd_fuse_frames: null argument");
    if (d_scores && (height % 8 != 0 || width % 8 != 0 || !d_upw || !d_upb))
        return fail(SD_ERR_INVALID, "sd_fuse_frames_scores: frame size must be a multiple of 8 and weights / bias must be given");
    if (height != ws->height || width != ws->width || batch < 1 || batch > ws->max_frames)
        return fail(SD_ERR_WORKSPACE, "sd_fuse_frames: shape/batch does not match the workspace");
    const bool any_hyp = d_hyp_road || d_hyp_left || d_hyp_right;
    if ((d_hyp_left != nullptr) != (d_hyp_right != nullptr)) return fail(SD_ERR_INVALID, "sd_fuse_frames: give both fence hypothesis sets or neither");
    if (any_hyp && (n_hyp < 1 || n_hyp > ws->max_hyp)) return fail(SD_ERR_WORKSPACE, "sd_fuse_frames: n_hyp exceeds the workspace reservation");
    if (params->sor_nb_neighbors < 1 || params->sor_nb_neighbors > kMaxKnnK) return fail(SD_ERR_INVALID, "sd_fuse_frames: sor_nb_neighbors must be in [1, 64]");
    cudaStream_t st = (cudaStream_t)stream;
    WsPriv* pv = priv(ws);
    const SdParams& P = *params;
    int rc;
    if (!ws->fused_ready || ws->fused_batch != batch || !same_params(ws->fused_params, P) || pv->hyp_road != d_hyp_road ||
        pv->hyp_left != d_hyp_left || pv->hyp_right != d_hyp_right || pv->n_hyp != n_hyp || pv->results != d_results ||
        memcmp(&pv->cam, cam, sizeof(SdCamera)) != 0) {
        cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cst);
        if (cst != cudaStreamCaptureStatusNone)
            return fail(SD_ERR_UNSUPPORTED, "sd_fuse_frames: first call with new parameters must happen outside stream capture");
        rc = build_fused_tables(ws, batch, P, *cam, d_hyp_road, d_hyp_left, d_hyp_right, n_hyp, d_results, st);
        if (rc) return rc;
    }
    const FusedTables& T = pv->t;
    const int B = batch, cap = ws->cap;
    const int cnt_stride = (int)(sizeof(FrameState) / sizeof(int32_t));
    cudaStreamCaptureStatus cap_st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap_st);
    const bool timing = pv->timing && cap_st == cudaStreamCaptureStatusNone;
    // stage mask: 1 pixel stage | 2 cloud stages up to the neighbour-search grid + the whole fence chain |
    //             4 the k-NN kernel | 8 radius search, final road compaction, slab, answers
    const int sm_ = pv->stage_mask;
    const bool do_pixel = (sm_ & 1) != 0, do_pre = (sm_ & 2) != 0, do_knn = (sm_ & 4) != 0, do_post = (sm_ & 8) != 0;
    // ---- pixel stage: rA <- road (z cut applied), fA <- fence
    if (timing) SD_CUDA_TRY(cudaEventRecord(pv->ev_t[0], st));
    // profiling mode (sd_ws_enable_timing on a full, uncaptured call): every stage is bracketed and the fence chain stays on
    // `st`, so the stage times add up to the call -- the reference's tic/toc pairs (semantic_depth.py:157-332)
    const bool stages = timing && sm_ == 15;
#define SD_MARK(k) do { if (stages) SD_CUDA_TRY(cudaEventRecord(pv->ev_s[k], st)); } while (0)
    SD_MARK(0);
    rc = !do_pixel ? SD_OK : sd_launch_pixel(d_logits, d_disp, ws->lmask, ws->rmask, B, height, width, *cam, P.prob_thr, P.road_z_to_meter, 0,
                         ws->road[0], ws->fence[0], cap,
                         &ws->fs[0].n[SD_CNT_ROAD_GATHER], &ws->fs[0].n[SD_CNT_ROAD_Z], &ws->fs[0].n[SD_CNT_FENCE_GATHER], cnt_stride,
                         nullptr, nullptr, nullptr, ws->pflags, ws->ptcounts, ws->ptoffs, ws->pix_tiles, st,
                         d_scores, d_upw, d_upb, nullptr, P.label_mode);
    if (rc) return rc;
    if (timing) SD_CUDA_TRY(cudaEventRecord(pv->ev_t[1], st));
    SD_MARK(SD_STAGE_PIXEL + 1);
    if (!do_pre && !do_knn && !do_post) return SD_OK;
    cudaStream_t sf = st;    // fence chain stream
    const bool fork = do_pre && P.approach_both && !pv->single_stream && !stages;
    if (fork) {
        sf = pv->side_stream;
        SD_CUDA_TRY(cudaEventRecord(pv->ev_fork, st));
        SD_CUDA_TRY(cudaStreamWaitEvent(sf, pv->ev_fork, 0));
    }
#define SD_RUN(expr) do { rc = (expr); if (rc) return rc; } while (0)
    // ---- road chain (stream st)
    if (do_pre) {
    SD_RUN(sd_launch_select_median(T.sel_road_y_med, B, cap, st));
    SD_RUN(sd_launch_select_median(T.sel_road_y_mad, B, cap, st));
    if (!pv->lazy_road) SD_RUN(sd_launch_compact(T.c_road_y, B, cap, st));
    SD_RUN(sd_launch_select_median(T.sel_road_x_med, B, cap, st));
    SD_RUN(sd_launch_select_median(T.sel_road_x_mad, B, cap, st));
    if (!pv->lazy_road) SD_RUN(sd_launch_compact(T.c_road_x, B, cap, st));
    SD_MARK(SD_STAGE_ROAD_MAD + 1);
    if (d_hyp_road) SD_RUN(sd_launch_ransac(T.r_road, B, cap, n_hyp, st));
    SD_RUN(sd_launch_plane(T.p_road, B, cap, st));
    SD_RUN(sd_launch_compact(T.c_road_plane, B, cap, st));
    SD_MARK(SD_STAGE_ROAD_PLANE + 1);
    if (P.use_sor || P.use_ror) SD_RUN(sd_launch_grid_build(T.k_road, B, cap, st));
    SD_MARK(SD_STAGE_ROAD_GRID + 1);
    }
    if (do_knn && P.use_sor) SD_RUN(sd_launch_knn(T.k_road, B, cap, P.sor_nb_neighbors, st));
    SD_MARK(SD_STAGE_ROAD_KNN + 1);
    if (do_post && P.use_ror) SD_RUN(sd_launch_radius(T.k_road, B, cap, st));
    if (do_post) {
        SD_RUN(sd_launch_compact(T.c_road_final, B, cap, st));
        SD_MARK(SD_STAGE_ROAD_ROR + 1);
        SD_RUN(sd_launch_slab(T.s_road, B, cap, st));
        SD_MARK(SD_STAGE_RW + 1);
    }
    // ---- fence chain (stream sf)
    if (do_pre && P.approach_both) {
        SD_RUN(sd_launch_select_median(T.sel_fence_y_med, B, cap, sf));
        SD_RUN(sd_launch_select_median(T.sel_fence_y_mad, B, cap, sf));
        SD_RUN(sd_launch_compact(T.c_fence_y, B, cap, sf));            // remove_noise_by_mad + threshold_complete
        SD_RUN(sd_launch_mean(T.m_fence, B, cap, sf));
        SD_RUN(sd_launch_compact(T.c_split, 2 * B, cap, sf));
        SD_RUN(sd_launch_select_median(T.sel_side_x_med, 2 * B, cap, sf));
        SD_RUN(sd_launch_select_median(T.sel_side_x_mad, 2 * B, cap, sf));
        if (!pv->lazy_side) SD_RUN(sd_launch_compact(T.c_side_x, 2 * B, cap, sf));
        if (d_hyp_left && d_hyp_right) SD_RUN(sd_launch_ransac(T.r_side, 2 * B, cap, n_hyp, sf));
        SD_RUN(sd_launch_plane(T.p_side, 2 * B, cap, sf));
        SD_RUN(sd_launch_compact(T.c_side_plane, 2 * B, cap, sf));
    }
    if (fork) {     // joins before the answers (or at the end of this call when the path is split across calls)
        SD_CUDA_TRY(cudaEventRecord(pv->ev_join, sf));
        SD_CUDA_TRY(cudaStreamWaitEvent(st, pv->ev_join, 0));
    }
    SD_MARK(SD_STAGE_FENCES + 1);
    if (do_post) SD_RUN(sd_launch_finalize(T.fin, B, &P, st));
#undef SD_RUN
    SD_MARK(SD_STAGE_ANSWERS + 1);
#undef SD_MARK
    pv->stages_valid = stages;
    if (timing) SD_CUDA_TRY(cudaEventRecord(pv->ev_t[2], st));
    return SD_OK;
}

extern "C" int sd_fuse_frames(const float* d_logits, const float* d_disp, int batch, int height, int width,
                              const SdCamera* cam, const SdParams* params,
                              const int32_t* d_hyp_road, const int32_t* d_hyp_left, const int32_t* d_hyp_right, int n_hyp,
                              SdFrameResult* d_results, SdWorkspace* ws, void* stream) {
    if (!d_logits) return fail(SD_ERR_INVALID, "sd_fuse_frames: null argument");
    return fuse_impl(d_logits, nullptr, nullptr, nullptr, d_disp, batch, height, width, cam, params, d_hyp_road, d_hyp_left,
                     d_hyp_right, n_hyp, d_results, ws, stream);
}

extern "C" int sd_fuse_frames_scores(const float* d_scores, const float* d_up_weights, const float* d_up_bias,
                                     const float* d_disp, int batch, int height, int width,
                                     const SdCamera* cam, const SdParams* params, SdFrameResult* d_results,
                                     SdWorkspace* ws, void* stream) {
    if (!d_scores) return fail(SD_ERR_INVALID, "sd_fuse_frames_scores: null argument");
    return fuse_impl(nullptr, d_scores, d_up_weights, d_up_bias, d_disp, batch, height, width, cam, params, nullptr, nullptr,
                     nullptr, 0, d_results, ws, stream);
}

extern "C" int sd_fuse_kernel_count(const SdParams* P, int with_ransac) {
    if (!P) return 0;
    const int sel = 3, ransac = with_ransac ? 3 : 0;
    int n = 3;                                   // pixel stage: label, scan, scatter
    n += 4 * sel + (with_ransac ? 2 : 0) + ransac + 1 + 1;   // road: 2 MADs (4 medians; their compactions only with RANSAC), plane fit + filter
    if (P->use_sor || P->use_ror) n += 4;        // grid: bbox, count, scan, scatter
    if (P->use_sor) n += 2;                      // k-NN: main kernel + heavy queries
    if (P->use_ror) n += 2;                      // statistical filter applied to the sorted copies + per-cell statistics, radius search
    n += 1 + 1;                                  // final road compaction, slab
    if (P->approach_both) n += 2 * sel + 1 + 2 + 1 + 2 * sel + (with_ransac ? 1 : 0) + ransac + 1 + 1;   // fence chain (MAD y + |z| in one compaction; np.mean = leaf sums + tree)
    n += 1;                                      // finalize
    return n;
}

extern "C" int sd_fuse_frames_host(const float* h_logits, const float* h_disp, int batch, int height, int width,
                                   const SdCamera* cam, const SdParams* params,
                                   float* d_logits_stage, float* d_disp_stage, SdFrameResult* d_results,
                                   SdFrameResult* h_results, SdWorkspace* ws, void* stream) {
    if (!h_logits || !h_disp || !d_logits_stage || !d_disp_stage || !h_results) return fail(SD_ERR_INVALID, "sd_fuse_frames_host: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t hw = (size_t)height * width;
    SD_CUDA_TRY(cudaMemcpyAsync(d_logits_stage, h_logits, sizeof(float) * 3 * hw * batch, cudaMemcpyHostToDevice, st));
    SD_CUDA_TRY(cudaMemcpyAsync(d_disp_stage, h_disp, sizeof(float) * 2 * hw * batch, cudaMemcpyHostToDevice, st));
    int rc = sd_fuse_frames(d_logits_stage, d_disp_stage, batch, height, width, cam, params, nullptr, nullptr, nullptr, 0,
                            d_results, ws, stream);
    if (rc) return rc;
    SD_CUDA_TRY(cudaMemcpyAsync(h_results, d_results, sizeof(SdFrameResult) * batch, cudaMemcpyDeviceToHost, st));
    SD_CUDA_TRY(cudaStreamSynchronize(st));
    return SD_OK;
}

int sd_launch_fcn_head(const float* l3, const float* l4, const float* l7, int batch, int h, int w, int c3, int c4, int c7,
                       const float* w3, const float* b3, const float* w4, const float* b4, const float* w7, const float* b7,
                       const float* wd1, const float* bd1, const float* wd2, const float* bd2,
                       float* s7, float* s4, float* s3, float* first_skip, float* second_skip, cudaStream_t st);

extern "C" size_t sd_fcn8s_head_scratch_bytes(int batch, int h8, int w8) {
    if (batch < 1 || h8 < 4 || w8 < 4 || (h8 % 4) || (w8 % 4)) return 0;
    const size_t n3 = (size_t)batch * h8 * w8;
    return sizeof(float) * 3 * (n3 / 16 + n3 / 4 + n3 + n3 / 4);        // conv_1x1_of_7, _of_4, _of_3, first_skip
}

extern "C" int sd_fcn8s_head(const float* d_layer3, const float* d_layer4, const float* d_layer7, int batch, int h8, int w8,
                             int c3, int c4, int c7, const SdFcnHeadWeights* w, float* d_scratch, size_t scratch_bytes,
                             float* d_scores, void* stream) {
    if (!d_layer3 || !d_layer4 || !d_layer7 || !w || !d_scratch || !d_scores) return fail(SD_ERR_INVALID, "sd_fcn8s_head: null argument");
    if (!w->conv3_w || !w->conv3_b || !w->conv4_w || !w->conv4_b || !w->conv7_w || !w->conv7_b || !w->deconv1_w || !w->deconv1_b ||
        !w->deconv2_w || !w->deconv2_b) return fail(SD_ERR_INVALID, "sd_fcn8s_head: null weight pointer");
    const size_t need = sd_fcn8s_head_scratch_bytes(batch, h8, w8);
    if (need == 0 || c3 < 1 || c4 < 1 || c7 < 1) return fail(SD_ERR_INVALID, "sd_fcn8s_head: the 1/8 map must be a multiple of 4 in both directions");
    if (scratch_bytes < need) return fail(SD_ERR_WORKSPACE, "sd_fcn8s_head: scratch too small (sd_fcn8s_head_scratch_bytes)");
    const size_t n3 = (size_t)batch * h8 * w8;
    float* s7 = d_scratch; float* s4 = s7 + 3 * (n3 / 16); float* s3 = s4 + 3 * (n3 / 4); float* fs = s3 + 3 * n3;
    return sd_launch_fcn_head(d_layer3, d_layer4, d_layer7, batch, h8, w8, c3, c4, c7, w->conv3_w, w->conv3_b, w->conv4_w, w->conv4_b,
                              w->conv7_w, w->conv7_b, w->deconv1_w, w->deconv1_b, w->deconv2_w, w->deconv2_b, s7, s4, s3, fs, d_scores,
                              (cudaStream_t)stream);
}

extern "C" int sd_ws_enable_timing(SdWorkspace* ws, int enable) {
    if (!ws) return fail(SD_ERR_INVALID, "sd_ws_enable_timing: null workspace");
    priv(ws)->timing = enable != 0;
    return SD_OK;
}

extern "C" int sd_ws_set_stage_mask(SdWorkspace* ws, int mask) {
    if (!ws || mask < 1 || mask > 15) return fail(SD_ERR_INVALID, "sd_ws_set_stage_mask: mask must be in [1, 15]");
    priv(ws)->stage_mask = mask;
    return SD_OK;
}

extern "C" int sd_ws_stage_elapsed_ms(SdWorkspace* ws, int which, float* h_ms) {
    if (!ws || !h_ms || which < 0 || which > 1) return fail(SD_ERR_INVALID, "sd_ws_stage_elapsed_ms: bad argument");
    WsPriv* pv = priv(ws);
    SD_CUDA_TRY(cudaEventElapsedTime(h_ms, pv->ev_t[0], pv->ev_t[which == 0 ? 1 : 2]));
    return SD_OK;
}

extern "C" int sd_ws_stage_alive(SdWorkspace* ws, int frame, int stage, const int32_t** d_src, const uint8_t** d_alive,
                                 const int32_t** d_rows) {
    if (!ws || !d_src || !d_alive || !d_rows || frame < 0 || frame >= ws->max_frames) return fail(SD_ERR_INVALID, "sd_ws_stage_alive: bad argument");
    const WsPriv* pv = priv(ws);
    const size_t cap = (size_t)ws->cap;
    const FrameState* fs = ws->fs + frame;
    switch (stage) {
        // (the road chain's input buffer is reused by the final road compaction, so its MAD stages are not retained)
        case SD_CNT_FENCE_MAD_Y:         // survivors of remove_noise_by_mad inside the fused MAD y + |z| compaction
            *d_src = ws->fence[0].src + frame * cap; *d_alive = ws->cflags + ((size_t)frame * 4 + 3) * cap; *d_rows = &fs->n[SD_CNT_FENCE_GATHER];
            return SD_OK;
        case SD_CNT_LEFT_MAD_X:
            if (!pv->lazy_side) break;
            *d_src = ws->left[0].src + frame * cap; *d_alive = ws->cflags + ((size_t)frame * 4 + 1) * cap; *d_rows = &fs->n[SD_CNT_LEFT_SPLIT];
            return SD_OK;
        case SD_CNT_RIGHT_MAD_X:
            if (!pv->lazy_side) break;
            *d_src = ws->right[0].src + frame * cap; *d_alive = ws->cflags + ((size_t)frame * 4 + 2) * cap; *d_rows = &fs->n[SD_CNT_RIGHT_SPLIT];
            return SD_OK;
        default: break;
    }
    return fail(SD_ERR_UNSUPPORTED, "sd_ws_stage_alive: this stage is either materialised (sd_ws_stage_src) or not retained");
}

extern "C" int sd_ws_stage_times(SdWorkspace* ws, float* h_ms) {
    if (!ws || !h_ms) return fail(SD_ERR_INVALID, "sd_ws_stage_times: null argument");
    WsPriv* pv = priv(ws);
    if (!pv->stages_valid) return fail(SD_ERR_UNSUPPORTED, "sd_ws_stage_times: the last fused call was not a full, uncaptured call with timing enabled");
    for (int k = 0; k < SD_NUM_STAGES; ++k) SD_CUDA_TRY(cudaEventElapsedTime(&h_ms[k], pv->ev_s[k], pv->ev_s[k + 1]));
    return SD_OK;
}

extern "C" int sd_ws_cloud(SdWorkspace* ws, int frame, int which, const float** d_x, const float** d_y, const float** d_z,
                           const int32_t** d_src, const int32_t** d_n) {
    if (!ws || frame < 0 || frame >= ws->max_frames || which < 0 || which > 2) return fail(SD_ERR_INVALID, "sd_ws_cloud: bad argument");
    // the side chains end in their second buffer when the MAD filter was not materialised (one compaction instead of two)
    const int side = priv(ws)->lazy_side ? 1 : 0;
    const SdCloudBuf& b = which == 0 ? ws->road[0] : (which == 1 ? ws->left[side] : ws->right[side]);
    SdCloudBuf fb = frame_buf(b, frame, ws->cap);
    if (d_x) *d_x = fb.x; if (d_y) *d_y = fb.y; if (d_z) *d_z = fb.z; if (d_src) *d_src = fb.src;
    const int cnt = which == 0 ? SD_CNT_ROAD_ROR : (which == 1 ? SD_CNT_LEFT_PLANE : SD_CNT_RIGHT_PLANE);
    if (d_n) *d_n = &ws->fs[frame].n[cnt];
    return SD_OK;
}

extern "C" int sd_ws_stage_src(SdWorkspace* ws, int frame, int stage, const int32_t** d_src) {
    if (!ws || !d_src || frame < 0 || frame >= ws->max_frames) return fail(SD_ERR_INVALID, "sd_ws_stage_src: bad argument");
    // buffers that still hold a stage's cloud when the fused call has finished
    const SdCloudBuf* b = nullptr;
    const WsPriv* pv = priv(ws);
    const int side = pv->lazy_side ? 1 : 0;
    switch (stage) {
        case SD_CNT_ROAD_ROR: b = &ws->road[0]; break;       // final road cloud
        case SD_CNT_ROAD_PLANE: b = &ws->road[1]; break;     // input of SOR/ROR
        case SD_CNT_FENCE_ABS_Z: b = &ws->fence[1]; break;   // input of the split
        case SD_CNT_LEFT_PLANE: b = &ws->left[side]; break;
        case SD_CNT_RIGHT_PLANE: b = &ws->right[side]; break;
        case SD_CNT_LEFT_MAD_X: if (!pv->lazy_side) b = &ws->left[1]; break;
        case SD_CNT_RIGHT_MAD_X: if (!pv->lazy_side) b = &ws->right[1]; break;
        default: break;
    }
    if (!b) return fail(SD_ERR_UNSUPPORTED, "sd_ws_stage_src: stage cloud is not materialised (see sd_ws_stage_alive)");
    *d_src = b->src + (size_t)frame * ws->cap;
    return SD_OK;
}
