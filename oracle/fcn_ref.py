"""NumPy restatement of the FCN-8s decoder head -- TEST INFRASTRUCTURE ONLY (see oracle/__init__).

Follows /root/reference/fcn8s/fcn.py:159-205 (``layers`` up to ``second_skip``): three 1x1 convolutions of the VGG
feature maps to ``num_classes`` = 3, two ``conv2d_transpose(4x4, stride 2, 'same')`` and the two skip additions.  The
last layer (fcn.py:207-213) is ``oracle.frame_ref.upsample_scores``.

TensorFlow's own summation order cannot be reproduced (cuDNN picks it), so parity is **unpinned** against TF; the
contract -- shared with the CUDA kernels (csrc/sd_fcn_head.cu) and stated in include/sd_fusion.h -- is fp32 without
FMA in this order:

* 1x1 convolution: 32 partial sums, partial ``l`` adds the channels ``l, l+32, l+64, ...`` in ascending order starting
  from 0.0; the partials are combined by the tree ``a[i] + a[i ^ 16]``, then ``^ 8, 4, 2, 1``; bias last.
* transposed convolution: taps added from 0.0 in the order (input row, input column, input channel) ascending; bias
  last; then ``+ skip`` (``tf.add(deconv, conv_1x1)``, fcn.py:194,205).
"""
from __future__ import annotations

import numpy as np


def conv1x1_to3(feat, weights, bias):
    """``tf.layers.conv2d(feat, 3, 1x1)`` (fcn.py:166-183).  feat [..., C] fp32, weights [C, 3], bias [3] -> [..., 3]."""
    f = np.asarray(feat, dtype=np.float32)
    w = np.asarray(weights, dtype=np.float32)
    C = f.shape[-1]
    flat = f.reshape(-1, C)
    part = np.zeros((flat.shape[0], 32, 3), dtype=np.float32)
    for c0 in range(0, C, 32):
        blk = flat[:, c0:c0 + 32]                                  # channels c0 .. c0+31 -> partials 0 .. 31
        n = blk.shape[1]
        part[:, :n, :] = part[:, :n, :] + blk[:, :, None] * w[None, c0:c0 + n, :]       # fp32 product, fp32 sum
    a = part
    for o in (16, 8, 4, 2, 1):                                     # lane i receives a[i] + a[i ^ o]; lane 0 ends with the sum
        a = a + a[:, np.arange(32) ^ o, :]
    out = a[:, 0, :] + np.asarray(bias, dtype=np.float32)[None, :]
    return out.reshape(f.shape[:-1] + (3,)).astype(np.float32)


def deconv4x4s2_add(x, weights, bias, skip):
    """``tf.add(conv2d_transpose(x, 3, 4x4, stride 2, 'same'), skip)`` (fcn.py:187-194 / 197-205).

    x [B, ih, iw, 3], weights [4, 4, 3(out), 3(in)] (TF layout kh, kw, out, in), bias [3], skip [B, 2ih, 2iw, 3].
    Output pixel (y, x) receives the taps ``ky = y + 1 - 2*iy`` in [0, 4): two input rows and two input columns
    (zeros outside the map take part in the sum like any other value)."""
    x = np.asarray(x, dtype=np.float32)
    w = np.asarray(weights, dtype=np.float32)
    b = np.asarray(bias, dtype=np.float32)
    B, ih, iw, _ = x.shape
    oh, ow = 2 * ih, 2 * iw
    y = np.arange(oh); xx = np.arange(ow)
    iy0 = (y + 1) // 2 - 1; ix0 = (xx + 1) // 2 - 1
    pad = np.zeros((B, ih + 2, iw + 2, 3), dtype=np.float32)
    pad[:, 1:-1, 1:-1] = x
    out = np.empty((B, oh, ow, 3), dtype=np.float32)
    for co in range(3):
        acc = np.zeros((B, oh, ow), dtype=np.float32)
        for dy in range(2):
            iy = iy0 + dy
            ky = y + 1 - 2 * iy
            for dx in range(2):
                ix = ix0 + dx
                kx = xx + 1 - 2 * ix
                for ci in range(3):
                    sv = pad[:, iy[:, None] + 1, ix[None, :] + 1, ci]
                    wv = w[ky[:, None], kx[None, :], co, ci]
                    acc = acc + sv * wv[None]
        out[..., co] = (acc + b[co]) + np.asarray(skip, dtype=np.float32)[..., co]
    return out


def fcn8s_head(layer3, layer4, layer7, weights):
    """fcn.py:159-205: (vgg_layer3_out, vgg_layer4_out, vgg_layer7_out) -> second_skip [B, h8, w8, 3].

    ``weights`` is a dict with the keys of ``semantic_depth_b200.fcn8s_head.init_head_weights``."""
    s7 = conv1x1_to3(layer7, weights["conv7_w"], weights["conv7_b"])
    s4 = conv1x1_to3(layer4, weights["conv4_w"], weights["conv4_b"])
    s3 = conv1x1_to3(layer3, weights["conv3_w"], weights["conv3_b"])
    first_skip = deconv4x4s2_add(s7, weights["deconv1_w"], weights["deconv1_b"], s4)
    return deconv4x4s2_add(first_skip, weights["deconv2_w"], weights["deconv2_b"], s3)
