// FCN-8s decoder head (fcn8s/fcn.py:159-215, `layers`): the part of the segmentation network between the three VGG
// feature maps and `second_skip`, the tensor the label kernel up-samples itself (score-map mode, sd_pixel.cu).
//
//   conv_1x1_of_7 = conv2d(vgg_layer7_out [h/4, w/4, C7], 3, 1x1)            fcn.py:166-170
//   conv_1x1_of_4 = conv2d(vgg_layer4_out [h/2, w/2, C4], 3, 1x1)            fcn.py:172-176
//   conv_1x1_of_3 = conv2d(vgg_layer3_out [h,   w,   C3], 3, 1x1)            fcn.py:178-183
//   first_skip    = conv2d_transpose(conv_1x1_of_7, 3, 4x4, stride 2, 'same') + conv_1x1_of_4      fcn.py:187-194
//   second_skip   = conv2d_transpose(first_skip,    3, 4x4, stride 2, 'same') + conv_1x1_of_3      fcn.py:197-205
//
// (h, w) = the 1/8-resolution map.  Everything is HBM-bound: a 1x1 convolution to 3 classes is three dot products per
// pixel over C channels (1.5 flop per byte read), the two transposed convolutions work on 3-channel maps.  No tensor
// cores.  Arithmetic contract (TF's own summation order is not reproducible, so this is the definition the oracle
// mirrors, oracle/fcn_ref.py): fp32, no FMA;
//   1x1 conv : one warp per pixel, lane l accumulates channels l, l+32, l+64, ... from 0.0 in ascending order, the 32
//              partial sums are combined by the xor-shuffle tree (16, 8, 4, 2, 1), bias added last;
//   deconv   : taps accumulated from 0.0 in the order (input row, input column, input channel) ascending, bias added
//              last, then `+ skip` (tf.add(deconv, conv_1x1), fcn.py:194,205).
#include "sd_internal.cuh"

namespace sd {

constexpr int kConvThreads = 256;      // 8 pixels per CTA

// features [B*npix][C] (NHWC), weights [C][3] (TF kernel layout [1][1][in][out]), bias [3]  ->  out [B*npix][3]
__global__ void __launch_bounds__(kConvThreads)
conv1x1_to3_kernel(const float* __restrict__ feat, const float* __restrict__ wgt, const float* __restrict__ bias,
                   float* __restrict__ out, long long npix, int C) {
    const int lane = lane_id();
    const long long warps = (long long)gridDim.x * (kConvThreads / 32);
    for (long long p = (long long)blockIdx.x * (kConvThreads / 32) + warp_id(); p < npix; p += warps) {
        const float* __restrict__ f = feat + p * C;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        int c = lane;
        for (; c + 96 < C; c += 128) {                      // four independent loads in flight, added in channel order
            const float f0 = __ldg(f + c), f1 = __ldg(f + c + 32), f2 = __ldg(f + c + 64), f3 = __ldg(f + c + 96);
            const float* w0 = wgt + (size_t)c * 3;
            a0 = a0 + f0 * __ldg(w0);       a1 = a1 + f0 * __ldg(w0 + 1);   a2 = a2 + f0 * __ldg(w0 + 2);
            a0 = a0 + f1 * __ldg(w0 + 96);  a1 = a1 + f1 * __ldg(w0 + 97);  a2 = a2 + f1 * __ldg(w0 + 98);
            a0 = a0 + f2 * __ldg(w0 + 192); a1 = a1 + f2 * __ldg(w0 + 193); a2 = a2 + f2 * __ldg(w0 + 194);
            a0 = a0 + f3 * __ldg(w0 + 288); a1 = a1 + f3 * __ldg(w0 + 289); a2 = a2 + f3 * __ldg(w0 + 290);
        }
        for (; c < C; c += 32) {
            const float fv = __ldg(f + c);
            const float* w0 = wgt + (size_t)c * 3;
            a0 = a0 + fv * __ldg(w0); a1 = a1 + fv * __ldg(w0 + 1); a2 = a2 + fv * __ldg(w0 + 2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 = a0 + __shfl_xor_sync(SD_FULL, a0, o);
            a1 = a1 + __shfl_xor_sync(SD_FULL, a1, o);
            a2 = a2 + __shfl_xor_sync(SD_FULL, a2, o);
        }
        if (lane == 0) {
            out[p * 3] = a0 + __ldg(bias); out[p * 3 + 1] = a1 + __ldg(bias + 1); out[p * 3 + 2] = a2 + __ldg(bias + 2);
        }
    }
}

// in [B][ih][iw][3], weights [4][4][3 out][3 in] (TF conv2d_transpose layout kh, kw, out, in), bias [3], skip and out
// [B][2 ih][2 iw][3]:  out = conv2d_transpose(in, 4x4, stride 2, 'same') + bias + skip.
// Output pixel (y, x) receives the taps ky = y + 1 - 2*iy in [0, 4): two input rows, two input columns (zero outside).
__global__ void __launch_bounds__(256)
deconv4x4s2_add_kernel(const float* __restrict__ in, const float* __restrict__ wgt, const float* __restrict__ bias,
                       const float* __restrict__ skip, float* __restrict__ out, int batch, int ih, int iw) {
    __shared__ float s_w[4 * 4 * 3 * 3];
    for (int i = threadIdx.x; i < 144; i += blockDim.x) s_w[i] = wgt[i];
    __syncthreads();
    const int oh = 2 * ih, ow = 2 * iw;
    const long long total = (long long)batch * oh * ow;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(t % ow), y = (int)((t / ow) % oh), b = (int)(t / ((long long)ow * oh));
        const int iy0 = ((y + 1) >> 1) - 1, ix0 = ((x + 1) >> 1) - 1;
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const int iy = iy0 + dy, ky = y + 1 - 2 * iy;
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int ix = ix0 + dx, kx = x + 1 - 2 * ix;
                float v[3] = {0.f, 0.f, 0.f};                 // zero outside the map: the products are added all the same
                if (iy >= 0 && iy < ih && ix >= 0 && ix < iw) {
                    const float* p = in + (((size_t)b * ih + iy) * iw + ix) * 3;
                    v[0] = __ldg(p); v[1] = __ldg(p + 1); v[2] = __ldg(p + 2);
                }
                const float* w = s_w + (ky * 4 + kx) * 9;
#pragma unroll
                for (int co = 0; co < 3; ++co)
#pragma unroll
                    for (int ci = 0; ci < 3; ++ci) acc[co] = acc[co] + v[ci] * w[co * 3 + ci];
            }
        }
        const size_t o = (size_t)t * 3;
#pragma unroll
        for (int co = 0; co < 3; ++co) out[o + co] = (acc[co] + __ldg(bias + co)) + __ldg(skip + o + co);
    }
}

}  // namespace sd

int sd_launch_fcn_head(const float* l3, const float* l4, const float* l7, int batch, int h, int w, int c3, int c4, int c7,
                       const float* w3, const float* b3, const float* w4, const float* b4, const float* w7, const float* b7,
                       const float* wd1, const float* bd1, const float* wd2, const float* bd2,
                       float* s7, float* s4, float* s3, float* first_skip, float* second_skip, cudaStream_t st) {
    using namespace sd;
    const long long n3 = (long long)batch * h * w, n4 = n3 / 4, n7 = n3 / 16;
    auto grid_conv = [](long long npix) { return (unsigned)max(1ll, min((npix + 7) / 8, (long long)148 * 8)); };
    conv1x1_to3_kernel<<<grid_conv(n7), kConvThreads, 0, st>>>(l7, w7, b7, s7, n7, c7);
    conv1x1_to3_kernel<<<grid_conv(n4), kConvThreads, 0, st>>>(l4, w4, b4, s4, n4, c4);
    conv1x1_to3_kernel<<<grid_conv(n3), kConvThreads, 0, st>>>(l3, w3, b3, s3, n3, c3);
    auto grid_dc = [](long long n) { return (unsigned)max(1ll, min((n + 255) / 256, (long long)148 * 8)); };
    deconv4x4s2_add_kernel<<<grid_dc(n4), 256, 0, st>>>(s7, wd1, bd1, s4, first_skip, batch, h / 4, w / 4);
    deconv4x4s2_add_kernel<<<grid_dc(n3), 256, 0, st>>>(first_skip, wd2, bd2, s3, second_skip, batch, h / 2, w / 2);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
