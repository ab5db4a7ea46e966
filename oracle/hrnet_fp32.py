"""Oracle (TEST INFRASTRUCTURE): HRNet pose network in torch fp32 on the CPU.

The reference runs an opaque ONNX artifact through onnxruntime
(human_body_length_est/modules/pose_estimator.py:47-59) or Triton; neither the
artifact nor onnxruntime exists offline, so -- as BASELINE.json prescribes --
the network is the PUBLIC HRNet definition ("pose_hrnet", Sun et al. CVPR 2019)
with random-init weights, evaluated here in float32 as the stand-in for "fp32
onnxruntime".  PARITY UNPINNED against the reference (it ships no network); this
file is an independent wiring of the architecture, sharing only parameter NAMES
(public state_dict keys) and values with the CUDA engine.

Input (B,3,H,W) RGB in [0,1] -> heatmaps (B,17,H/4,W/4).  BatchNorm is already
folded into (weight, bias) per conv.
"""
import torch
import torch.nn.functional as F


class HRNetFP32:
    def __init__(self, weights, width=32):
        """weights: {name: (W ndarray (cout,cin,k,k), b ndarray (cout,))}"""
        self.C = width
        self.p = {k: (torch.from_numpy(w).float(), torch.from_numpy(b).float()) for k, (w, b) in weights.items()}

    def cv(self, name, x, stride=1, relu=True):
        w, b = self.p[name]
        y = F.conv2d(x, w, b, stride=stride, padding=w.shape[-1] // 2)
        return F.relu(y) if relu else y

    def basic(self, pre, x):
        y = self.cv(pre + ".conv1", x)
        y = self.cv(pre + ".conv2", y, relu=False)
        return F.relu(y + x)

    def bottleneck(self, pre, x, project):
        idn = self.cv(pre + ".downsample.0", x, relu=False) if project else x
        y = self.cv(pre + ".conv1", x)
        y = self.cv(pre + ".conv2", y)
        y = self.cv(pre + ".conv3", y, relu=False)
        return F.relu(y + idn)

    def module(self, pre, xs, multi_scale=True):
        nb = len(xs)
        xs = list(xs)
        for i in range(nb):
            for blk in range(4):
                xs[i] = self.basic("%s.branches.%d.%d" % (pre, i, blk), xs[i])
        outs = []
        for i in range(nb if multi_scale else 1):
            acc = None
            for j in range(nb):
                if j == i:
                    t = xs[j]
                elif j > i:
                    t = self.cv("%s.fuse_layers.%d.%d.0" % (pre, i, j), xs[j], relu=False)
                    t = F.interpolate(t, scale_factor=2 ** (j - i), mode="nearest")
                else:
                    t = xs[j]
                    for k in range(i - j):
                        t = self.cv("%s.fuse_layers.%d.%d.%d.0" % (pre, i, j, k), t, stride=2,
                                    relu=(k != i - j - 1))
                acc = t if acc is None else acc + t
            outs.append(F.relu(acc))
        return outs

    @torch.no_grad()
    def forward(self, x, return_features=False, stages=None):
        """stages: optional dict that receives the activations at the stage boundaries, NCHW fp32:
        "conv2", "layer1", "transition1.{0,1}", "<module>.<branch>" for every HighResolutionModule output."""
        keep = (lambda k, t: stages.__setitem__(k, t.numpy().copy())) if stages is not None else (lambda k, t: None)
        x = torch.as_tensor(x).float()
        x = self.cv("conv1", x, stride=2)
        x = self.cv("conv2", x, stride=2)
        keep("conv2", x)
        for b in range(4):
            x = self.bottleneck("layer1.%d" % b, x, project=(b == 0))
        keep("layer1", x)
        xs = [self.cv("transition1.0.0", x), self.cv("transition1.1.0.0", x, stride=2)]
        keep("transition1.0", xs[0]); keep("transition1.1", xs[1])
        xs = self.module("stage2.0", xs)
        for i, t in enumerate(xs):
            keep("stage2.0.%d" % i, t)
        xs.append(self.cv("transition2.2.0.0", xs[-1], stride=2))
        for m in range(4):
            xs = self.module("stage3.%d" % m, xs)
            for i, t in enumerate(xs):
                keep("stage3.%d.%d" % (m, i), t)
        xs.append(self.cv("transition3.3.0.0", xs[-1], stride=2))
        for m in range(3):
            xs = self.module("stage4.%d" % m, xs, multi_scale=(m < 2))
            for i, t in enumerate(xs):
                keep("stage4.%d.%d" % (m, i), t)
        hm = self.cv("final_layer", xs[0], relu=False)
        if return_features:
            return hm, xs[0]
        return hm

    __call__ = forward
