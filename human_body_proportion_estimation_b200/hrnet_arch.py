"""HRNet-W32/W48 weights for the CUDA engine.

The reference ships no network source or weights (README.md:13-26 points at
Google-Drive artifacts); BASELINE.json asks for random-init weights of the named
architecture.  The conv program (names, shapes, blob offsets) is owned by the
C++ builder (csrc/hrnet.cu) and read back through hbp_hrnet_describe, so this
file only fills it:

* `random_weights`  -- seeded, *conditioned* random init of BN-folded convs:
  He-scaled convs, damped residual-branch outputs and 1/sqrt(n) fuse terms so
  that activations stay O(1) through all 36 residual blocks (plain Kaiming
  doubles the variance per block and overflows fp16; HRNet's own std=0.001 init
  vanishes).
* `pack`            -- {name: (W[cout,cin,k,k], b[cout])} -> fp16 blob in
  [tap][cout][cin] order + fp32 bias blob.
* `from_state_dict` -- fold a public HRNet checkpoint (conv + BN) into the same form.
"""
import numpy as np

from . import _capi


def layer_table(width=32, in_h=256, in_w=192):
    return _capi.describe_hrnet(width, in_h, in_w)


def _gain(name):
    if name == "final_layer":
        return 0.25
    if ".fuse_layers." in name:
        return 0.4
    if name.endswith(".conv2") and ".branches." in name:
        return 0.3           # BasicBlock residual branch output
    if name.endswith(".conv3"):
        return 0.3           # Bottleneck residual branch output
    if ".downsample." in name:
        return 1.0
    return 1.41421356        # conv followed by ReLU


def random_weights(width=32, in_h=256, in_w=192, seed=0):
    rows, _, _ = layer_table(width, in_h, in_w)
    rng = np.random.default_rng(seed)
    out = {}
    for name, cin, cout, k, stride, w_off, b_off in rows:
        fan_in = cin * k * k
        w = rng.standard_normal((cout, cin, k, k)).astype(np.float32) * np.float32(_gain(name) / np.sqrt(fan_in))
        b = (rng.standard_normal(cout) * 0.05).astype(np.float32)
        if name == "final_layer":
            b[:] = 0
        # the engine computes in fp16: round once here so every consumer (CUDA,
        # oracle) sees identical parameter values
        out[name] = (w.astype(np.float16).astype(np.float32), b)
    return out


def pack(weights, width=32, in_h=256, in_w=192):
    rows, nw, nb = layer_table(width, in_h, in_w)
    wb = np.zeros(nw, np.float16)
    bb = np.zeros(nb, np.float32)
    for name, cin, cout, k, stride, w_off, b_off in rows:
        w, b = weights[name]
        assert w.shape == (cout, cin, k, k), (name, w.shape, (cout, cin, k, k))
        wb[w_off:w_off + w.size] = np.transpose(w, (2, 3, 0, 1)).reshape(-1).astype(np.float16)
        bb[b_off:b_off + cout] = b
    return wb, bb


def from_state_dict(sd, width=32, in_h=256, in_w=192, eps=1e-5):
    """Fold conv+BN of a public pose_hrnet state_dict (numpy arrays) into
    {name: (W, b)}; BN modules sit next to their conv: conv1/bn1, x.0/x.1."""
    rows, _, _ = layer_table(width, in_h, in_w)
    out = {}
    for name, cin, cout, k, stride, _, _ in rows:
        w = np.asarray(sd[name + ".weight"], np.float32)
        if name == "final_layer":
            out[name] = (w, np.asarray(sd[name + ".bias"], np.float32))
            continue
        if name.endswith(".0"):
            bn = name[:-2] + ".1"
        else:
            head, last = name.rsplit(".", 1) if "." in name else ("", name)
            bn = (head + "." if head else "") + last.replace("conv", "bn")
        g, beta = np.asarray(sd[bn + ".weight"], np.float32), np.asarray(sd[bn + ".bias"], np.float32)
        mu, var = np.asarray(sd[bn + ".running_mean"], np.float32), np.asarray(sd[bn + ".running_var"], np.float32)
        s = g / np.sqrt(var + eps)
        out[name] = (w * s[:, None, None, None], beta - mu * s)
    return out


def flops_per_crop(width=32, in_h=256, in_w=192):
    """2 x MACs over all 293 convolutions of one forward (the algorithmic FLOP figure
    the roofline uses), and the per-shape-class breakdown {(cin,cout,k,stride,ho,wo): flops}."""
    rows, _, _ = _capi.describe_hrnet(width, in_h, in_w, full=True)
    total, classes = 0, {}
    for name, cin, cout, k, stride, _, _, ho, wo, up in rows:
        f = 2 * cin * cout * k * k * ho * wo
        total += f
        key = (cin, cout, k, stride, ho, wo)
        classes[key] = classes.get(key, 0) + f
    return total, classes
