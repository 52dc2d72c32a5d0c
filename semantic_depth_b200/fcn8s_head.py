"""FCN-8s decoder head on the GPU: the producer side of the score-map mode.

The reference's segmentation network (/root/reference/fcn8s/fcn.py) is a VGG16 encoder followed by ``layers``
(fcn.py:159-215): 1x1 convolutions of ``vgg_layer3/4/7_out`` to 3 classes, two stride-2 transposed convolutions with
skip additions (-> ``second_skip``) and a final 16x16 / stride-8 transposed convolution to full resolution.  The
fusion path already evaluates that last layer inside its label kernel (``FusionEngine.fuse_frames_scores``); this module
is everything between the VGG feature maps and ``second_skip``, so that the logits -- 12 B per pixel -- exist
neither on the host nor on the device.  The VGG16 encoder itself (and monodepth) stay input producers (north_star).

Weights are random-init exactly like the reference's (``tf.truncated_normal_initializer(stddev=0.01)``, zero biases,
fcn.py:161) unless given; TF layouts throughout, so a checkpoint's tensors can be passed as they are.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import SdFcnHeadWeights, check

VGG_CHANNELS = (256, 512, 4096)          # vgg_layer3_out, vgg_layer4_out, vgg_layer7_out (fcn.py:95-103)
_SHAPES = {"conv3_w": None, "conv3_b": (3,), "conv4_w": None, "conv4_b": (3,), "conv7_w": None, "conv7_b": (3,),
           "deconv1_w": (4, 4, 3, 3), "deconv1_b": (3,), "deconv2_w": (4, 4, 3, 3), "deconv2_b": (3,)}


def init_head_weights(seed: int = 0, channels=VGG_CHANNELS, bias_std: float = 0.0) -> dict:
    """Random-init weights of the head: truncated normal (|z| <= 2) x 0.01 like fcn.py:161; biases zero (TF default)
    unless ``bias_std`` > 0 (tests use non-zero biases so that the bias path is exercised)."""
    rng = np.random.default_rng(seed)

    def tn(shape):
        z = rng.standard_normal(shape)
        while True:
            bad = np.abs(z) > 2.0
            if not bad.any():
                break
            z[bad] = rng.standard_normal(int(bad.sum()))
        return (z * 0.01).astype(np.float32)

    c3, c4, c7 = channels
    w = {"conv3_w": tn((c3, 3)), "conv4_w": tn((c4, 3)), "conv7_w": tn((c7, 3)),
         "deconv1_w": tn((4, 4, 3, 3)), "deconv2_w": tn((4, 4, 3, 3))}
    for k in ("conv3_b", "conv4_b", "conv7_b", "deconv1_b", "deconv2_b"):
        w[k] = (rng.standard_normal(3) * bias_std).astype(np.float32)
    return w


class Fcn8sHead:
    """``scores = head(layer3, layer4, layer7)``: CUDA fp32 NHWC feature maps in, ``second_skip`` [B, h8, w8, 3] out."""

    def __init__(self, weights: dict | None = None, device="cuda:0", channels=VGG_CHANNELS):
        if not torch.cuda.is_available():
            raise _lib.SdError("Fcn8sHead needs a CUDA device: there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.channels = tuple(int(c) for c in channels)
        host = weights if weights is not None else init_head_weights(0, self.channels)
        c3, c4, c7 = self.channels
        want = dict(_SHAPES, conv3_w=(c3, 3), conv4_w=(c4, 3), conv7_w=(c7, 3))
        self.w = {}
        for k, shape in want.items():
            a = np.ascontiguousarray(np.asarray(host[k], dtype=np.float32))
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"{k}: expected shape {shape}, got {a.shape}")
            self.w[k] = torch.from_numpy(a).to(self.device)
        self._struct = SdFcnHeadWeights(**{k: self.w[k].data_ptr() for k in want})
        self._scratch = None

    def __call__(self, layer3: torch.Tensor, layer4: torch.Tensor, layer7: torch.Tensor, out: torch.Tensor | None = None):
        c3, c4, c7 = self.channels
        for t in (layer3, layer4, layer7):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise TypeError("Fcn8sHead needs contiguous CUDA float32 NHWC tensors")
        b, h, w = layer3.shape[:3]
        if tuple(layer3.shape) != (b, h, w, c3) or tuple(layer4.shape) != (b, h // 2, w // 2, c4) or \
                tuple(layer7.shape) != (b, h // 4, w // 4, c7) or h % 4 or w % 4:
            raise ValueError(f"expected layer3 [B,h,w,{c3}], layer4 [B,h/2,w/2,{c4}], layer7 [B,h/4,w/4,{c7}] with h, w multiples "
                             f"of 4; got {tuple(layer3.shape)}, {tuple(layer4.shape)}, {tuple(layer7.shape)}")
        need = int(self.lib.sd_fcn8s_head_scratch_bytes(b, h, w))
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        if out is None:
            out = torch.empty((b, h, w, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.sd_fcn8s_head(layer3.data_ptr(), layer4.data_ptr(), layer7.data_ptr(), b, h, w, c3, c4, c7,
                                         C.byref(self._struct), self._scratch.data_ptr(), need, out.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "sd_fcn8s_head")
        return out
