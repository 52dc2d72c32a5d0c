// cv2.resize(frame, (w, h), interpolation=cv2.INTER_CUBIC) on 8-bit frames: the first thing process_frame does
// with every frame (semantic_depth.py:110-112; SURVEY.md 8f rank 1 -- 54.7 % of the thesis' end-to-end time).
//
// OpenCV's own implementation (imgproc/src/resize.cpp 4.13, what runs with IPP off): bicubic weights (A = -0.75) evaluated in
// fp32 and rounded to 11 fractional bits (saturate_cast<short>(c * 2048)), an integer horizontal pass, and a vertical pass that
// is fp32 on the SIMD part of each row and integer ((v + 2^21) >> 22) on its tail; saturation.  Byte-identical to cv2.
// One thread per output pixel computes its own 4 + 4 weights (the fp32 expressions of interpolateCubic, evaluated in
// the same order; -fmad=false) and the 4 x 4 x channels integer taps.  The frame is read once (each source pixel
// belongs to about one window when shrinking by 4), so the kernel is HBM-bound: src bytes in + dst bytes out.
#include "sd_internal.cuh"

namespace sd {

struct CubicTap { int idx[4]; int w[4]; };

__device__ __forceinline__ CubicTap cubic_tap(int d, int src_n, double scale) {
    CubicTap t;
    const float f = (float)(((double)d + 0.5) * scale - 0.5);
    const float fl = floorf(f);
    const int s = (int)fl;
    const float x = f - fl;
    const float A = -0.75f;
    const float c0 = ((A * (x + 1.f) - 5.f * A) * (x + 1.f) + 8.f * A) * (x + 1.f) - 4.f * A;
    const float c1 = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    const float xm = 1.f - x;
    const float c2 = ((A + 2.f) * xm - (A + 3.f)) * xm * xm + 1.f;
    const float c3 = 1.f - c0 - c1 - c2;
    const float c[4] = {c0, c1, c2, c3};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int v = __float2int_rn(c[k] * 2048.f);                 // cvRound, then saturate_cast<short>
        t.w[k] = max(-32768, min(32767, v));
        t.idx[k] = max(0, min(src_n - 1, s - 1 + k));          // replicated border
    }
    return t;
}

template <int C>
__global__ void __launch_bounds__(256)
resize_cubic_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int src_h, int src_w, int dst_h, int dst_w,
                       double scale_x, double scale_y) {
    const int dx = blockIdx.x * 32 + (threadIdx.x & 31), dy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (dx >= dst_w || dy >= dst_h) return;
    const uint8_t* s = src + (size_t)blockIdx.z * src_h * src_w * C;
    const CubicTap tx = cubic_tap(dx, src_w, scale_x), ty = cubic_tap(dy, src_h, scale_y);
    // horizontal pass: the four row-buffer values of this output pixel (ints), per channel
    int hrow[4][C];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint8_t* row = s + (size_t)ty.idx[r] * src_w * C;
#pragma unroll
        for (int c = 0; c < C; ++c) hrow[r][c] = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint8_t* px = row + (size_t)tx.idx[k] * C;
#pragma unroll
            for (int c = 0; c < C; ++c) hrow[r][c] += (int)__ldg(px + c) * tx.w[k];
        }
    }
    // vertical pass.  OpenCV (resize.cpp, VResizeCubicVec_32s8u at the SSE baseline) evaluates the first 8 * floor(n / 8)
    // elements of a row of n = dst_w * channels ints in fp32 -- S0*b0 + (S1*b1 + (S2*b2 + S3*b3)), b = beta * 2^-22, every
    // product and sum rounded, round-half-even -- and the tail of the row in integers: (sum + 2^21) >> 22.
    const int n_vec = (dst_w * C) & ~7;
    const float b0 = (float)ty.w[0] * 0x1p-22f, b1 = (float)ty.w[1] * 0x1p-22f, b2 = (float)ty.w[2] * 0x1p-22f, b3 = (float)ty.w[3] * 0x1p-22f;
    uint8_t* o = dst + ((size_t)blockIdx.z * dst_h * dst_w + (size_t)dy * dst_w + dx) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        int v;
        if (dx * C + c < n_vec) {
            float acc = __fmul_rn((float)hrow[3][c], b3);
            acc = __fadd_rn(__fmul_rn((float)hrow[2][c], b2), acc);
            acc = __fadd_rn(__fmul_rn((float)hrow[1][c], b1), acc);
            acc = __fadd_rn(__fmul_rn((float)hrow[0][c], b0), acc);
            v = __float2int_rn(acc);
        } else {
            long long a = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r) a += (long long)hrow[r][c] * ty.w[r];
            v = (int)((a + (1ll << 21)) >> 22);
        }
        o[c] = (uint8_t)max(0, min(255, v));
    }
}

}  // namespace sd

int sd_launch_resize_cubic_u8(const uint8_t* d_src, int batch, int src_h, int src_w, int channels, uint8_t* d_dst, int dst_h, int dst_w,
                              cudaStream_t st) {
    using namespace sd;
    dim3 grid(ceil_div(dst_w, 32), ceil_div(dst_h, 8), batch);
    const double sx = (double)src_w / dst_w, sy = (double)src_h / dst_h;
    if (channels == 1) resize_cubic_u8_kernel<1><<<grid, 256, 0, st>>>(d_src, d_dst, src_h, src_w, dst_h, dst_w, sx, sy);
    else if (channels == 3) resize_cubic_u8_kernel<3><<<grid, 256, 0, st>>>(d_src, d_dst, src_h, src_w, dst_h, dst_w, sx, sy);
    else if (channels == 4) resize_cubic_u8_kernel<4><<<grid, 256, 0, st>>>(d_src, d_dst, src_h, src_w, dst_h, dst_w, sx, sy);
    else return SD_ERR_UNSUPPORTED;
    SD_LAUNCH_CHECK();
    return SD_OK;
}
