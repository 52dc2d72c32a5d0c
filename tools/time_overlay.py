"""Device timing of the overlay kernels on a 5-frame 1024x2048 batch (bytes: 4 B/px in + 3 B/px out + 1 B/px count)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semantic_depth_b200 import frame_ops

b, h, w = 5, 1024, 2048
g = torch.Generator(device="cuda").manual_seed(0)
frames = torch.randint(0, 256, (b, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
labels = torch.randint(0, 3, (b, h * w), dtype=torch.uint8, device="cuda", generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    frame_ops._overlay_labels(frames, labels, frame_ops.ROAD_RGBA, frame_ops.FENCE_RGBA)
ts = []
for _ in range(10):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); frame_ops._overlay_labels(frames, labels, frame_ops.ROAD_RGBA, frame_ops.FENCE_RGBA); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print(json.dumps({"op": "overlay_masks", "frames": b, "ms": ms, "GB_per_s": 8 * b * h * w / ms / 1e6,
                  "note": "includes torch.empty of the output and scratch; L2 flushed between iterations"}))
