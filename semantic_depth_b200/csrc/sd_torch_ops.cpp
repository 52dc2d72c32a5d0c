// TORCH_LIBRARY op layer over the C ABI of the fused path ("wrapped by a TORCH_LIBRARY op layer that
// checks device / dtype / contiguity and raises RuntimeError").
//
//   torch.ops.sd_fusion.fuse_frames(logits, disp, camera, params, hyp_road, hyp_left, hyp_right, workspace, results)
//   torch.ops.sd_fusion.fuse_frames_scores(scores, weights, bias, disp, camera, params, workspace, results)
//
// replace the fusion section of FrameProcessor.process_frame (/root/reference/semantic_depth.py:145-324) for a batch of
// frames; they validate every tensor in C++ (TORCH_CHECK -> RuntimeError), take the CUDA stream PyTorch is using on the
// tensors' device and forward plain pointers to sd_fuse_frames / sd_fuse_frames_scores (include/sd_fusion.h).  No kernel
// lives here: this translation unit is host C++ only and links libsd_fusion.so.
//   camera / params : CPU uint8 tensors holding the bytes of SdCamera / SdParams (the C structs of the ABI)
//   workspace       : the SdWorkspace* of sd_ws_create as an integer
//   results         : CUDA uint8 [B, sizeof(SdFrameResult)], written in place
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <string>

#include "sd_fusion.h"

namespace {

// Numbers go into the messages through std::to_string, never through operator<<(long): this library is loaded after the
// CUDA driver stack, and on the GPU boxes the lazily bound std::ostream integer inserter of THIS object resolved into an
// incompatible copy (segmentation fault while formatting a failed check); strings, devices and dtypes stream fine.
std::string num(int64_t v) { return std::to_string(v); }
std::string dims(const at::Tensor& t) {
    std::string s = "[";
    for (int64_t i = 0; i < t.dim(); ++i) s += (i ? ", " : "") + std::to_string(t.size(i));
    return s + "]";
}

void check_cuda(const at::Tensor& t, at::ScalarType dtype, const char* name, const c10::Device& dev) {
    TORCH_CHECK(t.is_cuda(), "sd_fusion: `", name, "` must be a CUDA tensor (got ", t.device(), "); there is no CPU path");
    TORCH_CHECK(t.device() == dev, "sd_fusion: `", name, "` is on ", t.device(), ", expected ", dev);
    TORCH_CHECK(t.scalar_type() == dtype, "sd_fusion: `", name, "` must be ", dtype, " (got ", t.scalar_type(), ")");
    TORCH_CHECK(t.is_contiguous(), "sd_fusion: `", name, "` must be contiguous");
}

template <typename T>
const T* host_struct(const at::Tensor& t, const char* name) {
    TORCH_CHECK(t.device().is_cpu() && t.scalar_type() == at::kByte && t.is_contiguous() && t.numel() == (int64_t)sizeof(T),
                "sd_fusion: `", name, "` must be a contiguous CPU uint8 tensor of ", num((int64_t)sizeof(T)), " bytes (the C struct of sd_fusion.h)");
    return reinterpret_cast<const T*>(t.data_ptr<uint8_t>());
}

SdWorkspace* workspace_of(int64_t handle) {
    TORCH_CHECK(handle != 0, "sd_fusion: null workspace handle");
    return reinterpret_cast<SdWorkspace*>(static_cast<uintptr_t>(handle));
}

SdFrameResult* results_of(const at::Tensor& results, int64_t batch, const c10::Device& dev) {
    check_cuda(results, at::kByte, "results", dev);
    TORCH_CHECK(results.numel() >= batch * (int64_t)sizeof(SdFrameResult), "sd_fusion: `results` holds ", num(results.numel()),
                " bytes, ", num(batch), " frames need ", num(batch * (int64_t)sizeof(SdFrameResult)));
    TORCH_CHECK(reinterpret_cast<uintptr_t>(results.data_ptr()) % alignof(SdFrameResult) == 0, "sd_fusion: `results` is misaligned");
    return reinterpret_cast<SdFrameResult*>(results.data_ptr<uint8_t>());
}

void check_disp(const at::Tensor& disp, int64_t batch, const c10::Device& dev) {
    check_cuda(disp, at::kFloat, "disp", dev);
    TORCH_CHECK(disp.dim() == 4 && disp.size(0) == batch && disp.size(1) == 2, "sd_fusion: `disp` must be [B, 2, H, W] with B = ",
                num(batch), " (got ", dims(disp), ")");
}

const int32_t* hypotheses(const c10::optional<at::Tensor>& h, const char* name, int64_t batch, const c10::Device& dev, int64_t& n_hyp) {
    if (!h.has_value() || !h->defined()) return nullptr;
    check_cuda(*h, at::kInt, name, dev);
    TORCH_CHECK(h->dim() == 3 && h->size(0) == batch && h->size(2) == 3, "sd_fusion: `", name, "` must be int32 [B, K, 3]");
    TORCH_CHECK(n_hyp == 0 || n_hyp == h->size(1), "sd_fusion: the hypothesis tables must hold the same number of triplets");
    n_hyp = h->size(1);
    return h->data_ptr<int32_t>();
}

void fuse_frames(const at::Tensor& logits, const at::Tensor& disp, const at::Tensor& camera, const at::Tensor& params,
                 const c10::optional<at::Tensor>& hyp_road, const c10::optional<at::Tensor>& hyp_left,
                 const c10::optional<at::Tensor>& hyp_right, int64_t workspace, at::Tensor results) {
    TORCH_CHECK(logits.is_cuda(), "sd_fusion::fuse_frames: `logits` must be a CUDA tensor (got ", logits.device(), "); there is no CPU path");
    const c10::Device dev = logits.device();
    check_cuda(logits, at::kFloat, "logits", dev);
    TORCH_CHECK(logits.dim() == 3 && logits.size(2) == 3, "sd_fusion: `logits` must be [B, H*W, 3] (got ", dims(logits), ")");
    const int64_t batch = logits.size(0);
    check_disp(disp, batch, dev);
    const int64_t h = disp.size(2), w = disp.size(3);
    TORCH_CHECK(logits.size(1) == h * w, "sd_fusion: `logits` has ", num(logits.size(1)), " pixels, `disp` ", num(h * w));
    int64_t n_hyp = 0;
    const int32_t* hr = hypotheses(hyp_road, "hyp_road", batch, dev, n_hyp);
    const int32_t* hl = hypotheses(hyp_left, "hyp_left", batch, dev, n_hyp);
    const int32_t* hg = hypotheses(hyp_right, "hyp_right", batch, dev, n_hyp);
    const SdCamera* cam = host_struct<SdCamera>(camera, "camera");
    const SdParams* ps = host_struct<SdParams>(params, "params");
    SdFrameResult* out = results_of(results, batch, dev);
    SdWorkspace* ws = workspace_of(workspace);
    c10::cuda::CUDAGuard guard(dev);
    void* stream = c10::cuda::getCurrentCUDAStream(dev.index()).stream();
    const int rc = sd_fuse_frames(logits.data_ptr<float>(), disp.data_ptr<float>(), (int)batch, (int)h, (int)w, cam, ps, hr, hl, hg,
                                  (int)n_hyp, out, ws, stream);
    TORCH_CHECK(rc == SD_OK, "sd_fuse_frames failed (", num(rc), "): ", sd_last_error());
}

void fuse_frames_scores(const at::Tensor& scores, const at::Tensor& weights, const at::Tensor& bias, const at::Tensor& disp,
                        const at::Tensor& camera, const at::Tensor& params, int64_t workspace, at::Tensor results) {
    TORCH_CHECK(scores.is_cuda(), "sd_fusion::fuse_frames_scores: `scores` must be a CUDA tensor (got ", scores.device(), "); there is no CPU path");
    const c10::Device dev = scores.device();
    check_cuda(scores, at::kFloat, "scores", dev);
    check_cuda(weights, at::kFloat, "weights", dev);
    check_cuda(bias, at::kFloat, "bias", dev);
    TORCH_CHECK(scores.dim() == 4 && scores.size(3) == 3, "sd_fusion: `scores` must be [B, H/8, W/8, 3] (got ", dims(scores), ")");
    TORCH_CHECK(weights.numel() == 16 * 16 * 3 * 3 && bias.numel() == 3, "sd_fusion: `weights` must be [16, 16, 3, 3] and `bias` [3]");
    const int64_t batch = scores.size(0);
    check_disp(disp, batch, dev);
    const int64_t h = disp.size(2), w = disp.size(3);
    TORCH_CHECK(scores.size(1) * 8 == h && scores.size(2) * 8 == w, "sd_fusion: `scores` ", dims(scores), " does not match `disp` ", dims(disp));
    const SdCamera* cam = host_struct<SdCamera>(camera, "camera");
    const SdParams* ps = host_struct<SdParams>(params, "params");
    SdFrameResult* out = results_of(results, batch, dev);
    SdWorkspace* ws = workspace_of(workspace);
    c10::cuda::CUDAGuard guard(dev);
    void* stream = c10::cuda::getCurrentCUDAStream(dev.index()).stream();
    const int rc = sd_fuse_frames_scores(scores.data_ptr<float>(), weights.data_ptr<float>(), bias.data_ptr<float>(), disp.data_ptr<float>(),
                                         (int)batch, (int)h, (int)w, cam, ps, out, ws, stream);
    TORCH_CHECK(rc == SD_OK, "sd_fuse_frames_scores failed (", num(rc), "): ", sd_last_error());
}

int64_t abi_version() { return sd_abi_version(); }
int64_t result_bytes() { return (int64_t)sizeof(SdFrameResult); }

}  // namespace

TORCH_LIBRARY(sd_fusion, m) {
    m.def("fuse_frames(Tensor logits, Tensor disp, Tensor camera, Tensor params, Tensor? hyp_road, Tensor? hyp_left, "
          "Tensor? hyp_right, int workspace, Tensor(a!) results) -> ()");
    m.def("fuse_frames_scores(Tensor scores, Tensor weights, Tensor bias, Tensor disp, Tensor camera, Tensor params, "
          "int workspace, Tensor(a!) results) -> ()");
    m.def("abi_version() -> int", abi_version);
    m.def("result_bytes() -> int", result_bytes);
}

// one implementation for every backend key: the device checks above are the dispatch (a CPU tensor is an error, not a fallback)
TORCH_LIBRARY_IMPL(sd_fusion, CompositeExplicitAutograd, m) {
    m.impl("fuse_frames", fuse_frames);
    m.impl("fuse_frames_scores", fuse_frames_scores);
}
