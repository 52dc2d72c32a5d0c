// Segmentation overlay of SegmentFrame.segment_frame (semantic_depth.py:547-568; SURVEY.md 8f rank 4, the mask-paste
// part): the road mask, then the fence mask, pasted as translucent RGBA layers onto the 8-bit frame.
//
//   road_mask = toimage(np.dot(mask, [[128, 64, 128, 64]]), mode="RGBA");  street_im.paste(road_mask, None, road_mask)
//
// scipy 1.2.1's toimage stretches the int64 layer from [min, max] of the WHOLE array to [0, 255] (bytescale), so the
// layer's bytes depend on whether the mask is empty, partial or covers every pixel: the host evaluates bytescale for
// the partial and the full case (sd_api.cu), a first kernel counts the mask pixels per frame, and the blend kernel
// picks the variant from the counts in device memory (no host round trip).  PIL's paste is
// out = MULDIV255(dst * (255 - a) + src * a), MULDIV255(t) = ((t + 128 >> 8) + t + 128) >> 8.
// Byte work, HBM-bound: 4 B/pixel in (frame + label), 3 B/pixel out, plus 1 B/pixel for the count.
#include "sd_internal.cuh"

namespace sd {

constexpr int kOverlayThreads = 256;

// counts[2 * frame + {0, 1}] += number of road / fence pixels
__global__ void __launch_bounds__(kOverlayThreads)
overlay_count_kernel(const uint8_t* __restrict__ labels, int hw, int* __restrict__ counts) {
    const uint8_t* lab = labels + (size_t)blockIdx.y * hw;
    const int head = min(hw, (int)((4 - (reinterpret_cast<uintptr_t>(lab) & 3)) & 3));
    const int nwords = (hw - head) >> 2;
    const int tail0 = head + (nwords << 2);
    int nr = 0, nf = 0;
    const int tid = blockIdx.x * kOverlayThreads + threadIdx.x, nth = gridDim.x * kOverlayThreads;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(lab + head);
    for (int i = tid; i < nwords; i += nth) {
        const uint32_t v = __ldg(w + i);
        nr += __popc(v & 0x01010101u);
        nf += __popc(v & 0x02020202u);
    }
    if (tid < head) { nr += lab[tid] & 1; nf += (lab[tid] >> 1) & 1; }
    if (tid < hw - tail0) { nr += lab[tail0 + tid] & 1; nf += (lab[tail0 + tid] >> 1) & 1; }
    nr = __reduce_add_sync(SD_FULL, nr);
    nf = __reduce_add_sync(SD_FULL, nf);
    __shared__ int s_r[kOverlayThreads / 32], s_f[kOverlayThreads / 32];
    if ((threadIdx.x & 31) == 0) { s_r[threadIdx.x >> 5] = nr; s_f[threadIdx.x >> 5] = nf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int r = 0, f = 0;
#pragma unroll
        for (int k = 0; k < kOverlayThreads / 32; ++k) { r += s_r[k]; f += s_f[k]; }
        if (r) atomicAdd(counts + 2 * blockIdx.y, r);
        if (f) atomicAdd(counts + 2 * blockIdx.y + 1, f);
    }
}

__device__ __forceinline__ uint32_t blend8(uint32_t dst, uint32_t src, uint32_t a) {
    const uint32_t t = dst * (255u - a) + src * a + 128u;
    return ((t >> 8) + t) >> 8;
}

// one pixel: road layer first, then the fence layer on top of the result (layers are RGBA packed little-endian)
__device__ __forceinline__ void blend_pixel(uint32_t& r, uint32_t& g, uint32_t& b, uint32_t label, uint32_t road, uint32_t fence) {
    if (label & 1u) {
        const uint32_t a = road >> 24;
        r = blend8(r, road & 255u, a); g = blend8(g, (road >> 8) & 255u, a); b = blend8(b, (road >> 16) & 255u, a);
    }
    if (label & 2u) {
        const uint32_t a = fence >> 24;
        r = blend8(r, fence & 255u, a); g = blend8(g, (fence >> 8) & 255u, a); b = blend8(b, (fence >> 16) & 255u, a);
    }
}

template <bool VEC>
__global__ void __launch_bounds__(kOverlayThreads)
overlay_blend_kernel(const uint8_t* __restrict__ frame, const uint8_t* __restrict__ labels, int hw, OverlayLayers L,
                     const int* __restrict__ counts, uint8_t* __restrict__ out) {
    const int f = blockIdx.y;
    const int nr = counts[2 * f], nf = counts[2 * f + 1];
    // an empty mask is an all-zero layer (alpha 0: the paste is the identity); set pixels of a partial / full mask
    const uint32_t road = nr == hw ? L.road_full : L.road_partial;
    const uint32_t fence = nf == hw ? L.fence_full : L.fence_partial;
    const uint8_t* src = frame + (size_t)f * hw * 3;
    const uint8_t* lab = labels + (size_t)f * hw;
    uint8_t* dst = out + (size_t)f * hw * 3;
    const int tid = blockIdx.x * kOverlayThreads + threadIdx.x, nth = gridDim.x * kOverlayThreads;
    if (VEC) {
        // 4 pixels per step: 12 frame bytes as three words, 4 labels as one
        const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
        const uint32_t* l4 = reinterpret_cast<const uint32_t*>(lab);
        uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
        for (int i = tid; i < (hw >> 2); i += nth) {
            const uint32_t lw = __ldg(l4 + i);
            uint32_t w0 = __ldg(s4 + 3 * i), w1 = __ldg(s4 + 3 * i + 1), w2 = __ldg(s4 + 3 * i + 2);
            if (lw & 0x03030303u) {
                uint32_t c[12];
#pragma unroll
                for (int k = 0; k < 4; ++k) { c[k] = (w0 >> (8 * k)) & 255u; c[4 + k] = (w1 >> (8 * k)) & 255u; c[8 + k] = (w2 >> (8 * k)) & 255u; }
#pragma unroll
                for (int p = 0; p < 4; ++p) blend_pixel(c[3 * p], c[3 * p + 1], c[3 * p + 2], (lw >> (8 * p)) & 255u, road, fence);
                w0 = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
                w1 = c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24);
                w2 = c[8] | (c[9] << 8) | (c[10] << 16) | (c[11] << 24);
            }
            d4[3 * i] = w0; d4[3 * i + 1] = w1; d4[3 * i + 2] = w2;
        }
    } else {
        for (int i = tid; i < hw; i += nth) {
            uint32_t r = src[3 * (size_t)i], g = src[3 * (size_t)i + 1], b = src[3 * (size_t)i + 2];
            blend_pixel(r, g, b, lab[i], road, fence);
            dst[3 * (size_t)i] = (uint8_t)r; dst[3 * (size_t)i + 1] = (uint8_t)g; dst[3 * (size_t)i + 2] = (uint8_t)b;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Result banner of process_frame (semantic_depth.py:339-394, sequence:304-327): cv2.rectangle(..., -1) + cv2.putText.
// The glyph bitmaps come from an atlas baked from OpenCV's own Hershey rasteriser (semantic_depth_b200/data/
// make_hershey_atlas.py); the host decides which bitmap goes where (pen arithmetic in 16.16 fixed point, frame_ops.py).
//   rect  = {frame, x0, y0, x1, y1, colour}: inclusive corners, already ordered; clipped here
//   place = {frame, bitmap, x, y, colour}  : top-left pixel of the bitmap's cell; clipped here
// colour = channel 0 | channel 1 << 8 | channel 2 << 16, in the frame's own channel order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kOverlayThreads)
banner_rect_kernel(uint8_t* __restrict__ frames, int batch, int h, int w, const int32_t* __restrict__ rects) {
    const int32_t* r = rects + 6 * blockIdx.y;
    const int f = r[0];
    const int x0 = max(r[1], 0), y0 = max(r[2], 0), x1 = min(r[3], w - 1), y1 = min(r[4], h - 1);
    if (f < 0 || f >= batch || x1 < x0 || y1 < y0) return;
    const uint32_t col = (uint32_t)r[5];
    const int rowbytes = (x1 - x0 + 1) * 3;
    const long long total = (long long)rowbytes * (y1 - y0 + 1);
    uint8_t* base = frames + ((size_t)f * h * w) * 3;
    for (long long i = (long long)blockIdx.x * kOverlayThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kOverlayThreads) {
        const int row = (int)(i / rowbytes), b = (int)(i - (long long)row * rowbytes);
        base[((size_t)(y0 + row) * w + x0) * 3 + b] = (uint8_t)(col >> (8 * (b % 3)));
    }
}

constexpr int kGlyphThreads = 128;
__global__ void __launch_bounds__(kGlyphThreads)
banner_glyph_kernel(uint8_t* __restrict__ frames, int batch, int h, int w, const uint32_t* __restrict__ bits, int nbitmaps,
                    int cell_h, int words, const int32_t* __restrict__ places) {
    const int32_t* p = places + 5 * blockIdx.x;
    const int f = p[0], g = p[1], px = p[2], py = p[3];
    if (f < 0 || f >= batch || g < 0 || g >= nbitmaps) return;
    const uint32_t col = (uint32_t)p[4];
    const uint32_t* bm = bits + (size_t)g * cell_h * words;
    uint8_t* base = frames + ((size_t)f * h * w) * 3;
    for (int i = threadIdx.x; i < cell_h * words; i += kGlyphThreads) {
        uint32_t v = __ldg(bm + i);
        if (!v) continue;
        const int row = i / words, wd = i - row * words;
        const int y = py + row;
        if (y < 0 || y >= h) continue;
        while (v) {
            const int b = __ffs(v) - 1;
            v &= v - 1;
            const int x = px + wd * 32 + b;
            if (x < 0 || x >= w) continue;
            uint8_t* q = base + ((size_t)y * w + x) * 3;
            q[0] = (uint8_t)col; q[1] = (uint8_t)(col >> 8); q[2] = (uint8_t)(col >> 16);
        }
    }
}

}  // namespace sd

int sd_launch_overlay(const uint8_t* d_frame, const uint8_t* d_labels, int batch, int hw, const sd::OverlayLayers& layers,
                      int* d_counts, uint8_t* d_out, cudaStream_t st) {
    using namespace sd;
    SD_CUDA_TRY(cudaMemsetAsync(d_counts, 0, sizeof(int) * 2 * (size_t)batch, st));
    const int per = max(1, min(ceil_div(hw, kOverlayThreads * 16), max(1, (148 * 8) / batch)));
    overlay_count_kernel<<<dim3(per, batch), kOverlayThreads, 0, st>>>(d_labels, hw, d_counts);
    SD_LAUNCH_CHECK();
    const bool vec = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_frame) | reinterpret_cast<uintptr_t>(d_labels) |
                                        reinterpret_cast<uintptr_t>(d_out)) & 3) == 0;
    if (vec) overlay_blend_kernel<true><<<dim3(per, batch), kOverlayThreads, 0, st>>>(d_frame, d_labels, hw, layers, d_counts, d_out);
    else overlay_blend_kernel<false><<<dim3(per, batch), kOverlayThreads, 0, st>>>(d_frame, d_labels, hw, layers, d_counts, d_out);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

int sd_launch_banner(uint8_t* d_frames, int batch, int h, int w, const int32_t* d_rects, int nrects,
                     const uint32_t* d_bits, int nbitmaps, int cell_h, int words, const int32_t* d_places, int nplaces,
                     cudaStream_t st) {
    using namespace sd;
    if (nrects > 0) {
        // rectangles first (the text is drawn over the banner); a rectangle is at most the whole frame
        const int per = max(1, min(ceil_div(h * w * 3, kOverlayThreads * 16), (148 * 8) / nrects + 1));
        banner_rect_kernel<<<dim3(per, nrects), kOverlayThreads, 0, st>>>(d_frames, batch, h, w, d_rects);
    }
    if (nplaces > 0)
        banner_glyph_kernel<<<nplaces, kGlyphThreads, 0, st>>>(d_frames, batch, h, w, d_bits, nbitmaps, cell_h, words, d_places);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
