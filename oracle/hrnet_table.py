"""Oracle (TEST INFRASTRUCTURE): the convolution table of the public HRNet "pose_hrnet" definition (Sun et al.
CVPR 2019) in pure Python -- names (public state_dict prefixes), channel counts, kernel size and stride of every
convolution -- and the seeded random initialisation used for parity runs.

The product reads the same table out of its C library (hbp_hrnet_describe); this copy exists so that the CPU
reference arm (`bench.py --impl reference`) and the oracle network never touch the library under test.
tests/test_host_cpu.py checks that both tables and both weight generators agree.
"""
import numpy as np


def _module(rows, pre, nb, ch, multi_scale):
    for i in range(nb):
        for blk in range(4):
            rows.append(("%s.branches.%d.%d.conv1" % (pre, i, blk), ch[i], ch[i], 3, 1))
            rows.append(("%s.branches.%d.%d.conv2" % (pre, i, blk), ch[i], ch[i], 3, 1))
    n_out = nb if multi_scale else 1
    n_levels = max(1, n_out - 1)
    # fuse layers in the order the engine schedules them: link k = level-1 of every stride-2 chain (i > j), then at
    # level 1 the 1x1 convs of the upsampled terms (j > i)
    for level in range(1, n_levels + 1):
        for i in range(level, n_out):
            for j in range(0, i - level + 1):
                k = level - 1
                last = level == i - j
                rows.append(("%s.fuse_layers.%d.%d.%d.0" % (pre, i, j, k), ch[j], ch[i] if last else ch[j], 3, 2))
        if level == 1:
            for i in range(min(n_out, nb - 1)):
                for j in range(i + 1, nb):
                    rows.append(("%s.fuse_layers.%d.%d.0" % (pre, i, j), ch[j], ch[i], 1, 1))


def layer_table(width=32):
    """[(name, cin, cout, k, stride)] in weight-blob order"""
    C = width
    rows = [("conv1", 3, 64, 3, 2), ("conv2", 64, 64, 3, 2)]
    cin = 64
    for b in range(4):
        p = "layer1.%d" % b
        if b == 0:
            rows.append((p + ".downsample.0", cin, 256, 1, 1))
        rows += [(p + ".conv1", cin, 64, 1, 1), (p + ".conv2", 64, 64, 3, 1), (p + ".conv3", 64, 256, 1, 1)]
        cin = 256
    rows += [("transition1.0.0", 256, C, 3, 1), ("transition1.1.0.0", 256, 2 * C, 3, 2)]
    ch = [C, 2 * C]
    _module(rows, "stage2.0", 2, ch, True)
    rows.append(("transition2.2.0.0", 2 * C, 4 * C, 3, 2))
    ch = ch + [4 * C]
    for m in range(4):
        _module(rows, "stage3.%d" % m, 3, ch, True)
    rows.append(("transition3.3.0.0", 4 * C, 8 * C, 3, 2))
    ch = ch + [8 * C]
    for m in range(3):
        _module(rows, "stage4.%d" % m, 4, ch, m < 2)
    rows.append(("final_layer", C, 17, 1, 1))
    return rows


def _gain(name):
    if name == "final_layer":
        return 0.25
    if ".fuse_layers." in name:
        return 0.4
    if name.endswith(".conv2") and ".branches." in name:
        return 0.3
    if name.endswith(".conv3"):
        return 0.3
    if ".downsample." in name:
        return 1.0
    return 1.41421356


def random_weights(width=32, seed=0):
    """{name: (W (cout,cin,k,k) float32 holding fp16-representable values, b (cout,) float32)} -- the same values the
    product's hrnet_arch.random_weights draws (same generator, same order)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, cin, cout, k, stride in layer_table(width):
        fan_in = cin * k * k
        w = rng.standard_normal((cout, cin, k, k)).astype(np.float32) * np.float32(_gain(name) / np.sqrt(fan_in))
        b = (rng.standard_normal(cout) * 0.05).astype(np.float32)
        if name == "final_layer":
            b[:] = 0
        out[name] = (w.astype(np.float16).astype(np.float32), b)
    return out


def flops_per_crop(width=32, in_h=256, in_w=192):
    """2 x MACs of every convolution of one forward"""
    total = 0
    # output sizes: stem /2, /4; branch i at /4 / 2^i
    def hw(name):
        if name == "conv1":
            return in_h // 2, in_w // 2
        if name == "conv2" or name.startswith("layer1") or name == "transition1.0.0" or name == "final_layer":
            return in_h // 4, in_w // 4
        if name == "transition1.1.0.0":
            return in_h // 8, in_w // 8
        if name.startswith("transition2"):
            return in_h // 16, in_w // 16
        if name.startswith("transition3"):
            return in_h // 32, in_w // 32
        parts = name.split(".")
        if "branches" in parts:
            i = int(parts[parts.index("branches") + 1])
            return in_h // 4 >> i, in_w // 4 >> i
        # fuse layers: stageS.M.fuse_layers.i.j[.k].0
        f = parts.index("fuse_layers")
        i, j = int(parts[f + 1]), int(parts[f + 2])
        if j > i:                       # 1x1 at the low resolution of branch j
            return in_h // 4 >> j, in_w // 4 >> j
        k = int(parts[f + 3])           # link k of the chain j -> i: output at branch j+k+1
        return in_h // 4 >> (j + k + 1), in_w // 4 >> (j + k + 1)
    for name, cin, cout, k, stride in layer_table(width):
        h, w = hw(name)
        total += 2 * cin * cout * k * k * h * w
    return total
