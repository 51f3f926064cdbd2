"""Seeded weights and synthetic image pairs (no network, no checkpoints).

The reference's pretrained checkpoint cannot be downloaded offline
(pretrained/download.sh), so parity and throughput are measured on seeded
random weights that carry the reference's ``state_dict`` key names and shapes
(names probed from networks/patch2pix.py:13-61,
networks/modules.py:56-99, networks/ncn/conv4d.py:118-120,
networks/resnet.py:96-123).  BatchNorm running statistics / affines and all
biases are randomised (a default-initialised BN is the identity and would hide
folding bugs) and the running variances are calibrated so that activations
stay O(1) through the regressor, like a trained network's.
"""
import math

import torch
import torch.nn.functional as F

_LAYERS = (('layer1', 3, 64), ('layer2', 4, 128), ('layer3', 6, 256))


def _xavier(gen, *shape, fan_in=None, fan_out=None):
    rf = 1
    for s in shape[2:]:
        rf *= s
    fan_in = fan_in if fan_in is not None else shape[1] * rf
    fan_out = fan_out if fan_out is not None else shape[0] * rf
    bound = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(*shape, generator=gen) * 2 - 1) * bound


def _bn(sd, gen, name, c, var):
    sd[name + '.weight'] = 0.8 + 0.4 * torch.rand(c, generator=gen)
    sd[name + '.bias'] = 0.1 * torch.randn(c, generator=gen)
    sd[name + '.running_mean'] = 0.1 * math.sqrt(var) * torch.randn(c, generator=gen)
    sd[name + '.running_var'] = var * (0.7 + 0.6 * torch.rand(c, generator=gen))


def make_seeded_state_dict(seed=0, backbone=True, regressors=True, nc_init='uniform'):
    """Flat fp32 dict with the reference's state_dict names (layer4 / num_batches_tracked omitted).

    nc_init: 'uniform' -- NeighConsensus weights uniform in +-0.1 (an untrained net: its output is
    unrelated to the correlation, a pair yields 10-20 mutual matches and exact zeros / ties abound);
    'consensus' -- trained-like filters (positive on the 4D-diagonal taps, slightly negative elsewhere):
    coherent match neighbourhoods are reinforced, the rest is zeroed by the ReLUs, so a pair of overlapping
    views yields ~1000 distinct mutual matches at 640x480 with top-1/top-2 margins >~ 1e-5 (the benchmark workload).
    Every other tensor is identical between the two modes."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    if backbone:
        sd['extract.conv1.weight'] = _xavier(g, 64, 3, 7, 7)
        _bn(sd, g, 'extract.bn1', 64, 1.0)
        cin = 64
        for lname, nblk, c in _LAYERS:
            for i in range(nblk):
                p = f'extract.{lname}.{i}'
                sd[p + '.conv1.weight'] = _xavier(g, c, cin if i == 0 else c, 3, 3)
                _bn(sd, g, p + '.bn1', c, 1.0)
                sd[p + '.conv2.weight'] = _xavier(g, c, c, 3, 3)
                _bn(sd, g, p + '.bn2', c, 1.0)
                if i == 0 and cin != c:
                    sd[p + '.downsample.0.weight'] = _xavier(g, c, cin, 1, 1)
                    _bn(sd, g, p + '.downsample.1', c, 1.0)
            cin = c
    # NCNet: Conv4d weights are stored pre-permuted [k1, Cout, Cin, k2, k3, k4]
    sd['ncn.conv.0.weight'] = (torch.rand(3, 16, 1, 3, 3, 3, generator=g) * 2 - 1) * 0.1
    sd['ncn.conv.0.bias'] = 0.01 * torch.randn(16, generator=g)
    sd['ncn.conv.2.weight'] = (torch.rand(3, 1, 16, 3, 3, 3, generator=g) * 2 - 1) * 0.1
    sd['ncn.conv.2.bias'] = 0.01 * torch.randn(1, generator=g)
    if nc_init == 'consensus':
        # what a trained neighbourhood-consensus filter looks like: positive weights on the 9 "diagonal" taps, where the
        # A-offset equals the B-offset ((a+d, b+d) neighbours of a true match are matches too), slightly negative
        # elsewhere, so that incoherent / constant regions cancel and ReLU zeroes them (~97 % exact zeros at 640x480)
        gn = torch.Generator().manual_seed(12345 + seed)      # own stream: the tensors below stay as in 'uniform'

        def layer(cout, cin):                                   # layout [k1, Cout, Cin, k2, k3, k4] = taps (a, b, d, e)
            w = (torch.rand(3, cout, cin, 3, 3, 3, generator=gn) * 2 - 1) * 0.02 - 0.045
            for a in range(3):
                for b in range(3):
                    w[a, :, :, b, a, b] += 0.4 * (0.5 + torch.rand(cout, cin, generator=gn))
            return w
        sd['ncn.conv.0.weight'], sd['ncn.conv.2.weight'] = layer(16, 1), layer(1, 16)
    elif nc_init != 'uniform':
        raise ValueError("nc_init must be 'uniform' or 'consensus'")
    if regressors:
        for r in ('regress_mid', 'regress_fine'):
            sd[f'{r}.conv.0.weight'] = _xavier(g, 512, 518, 3, 3)
            _bn(sd, g, f'{r}.conv.1', 512, 0.0039)
            sd[f'{r}.conv.2.weight'] = _xavier(g, 512, 512, 3, 3)
            _bn(sd, g, f'{r}.conv.3', 512, 1.0)
            sd[f'{r}.fc.0.weight'] = _xavier(g, 512, 512)
            sd[f'{r}.fc.0.bias'] = 0.05 * torch.randn(512, generator=g)
            _bn(sd, g, f'{r}.fc.1', 512, 4.0)
            sd[f'{r}.fc.3.weight'] = _xavier(g, 256, 512)
            sd[f'{r}.fc.3.bias'] = 0.05 * torch.randn(256, generator=g)
            _bn(sd, g, f'{r}.fc.4', 256, 0.7)
            sd[f'{r}.fc.6.weight'] = _xavier(g, 5, 256)
            sd[f'{r}.fc.6.bias'] = 0.05 * torch.randn(5, generator=g)
    return sd


def synthetic_pair(pair_idx, height, width):
    """Deterministic 'two views of one texture' pair: a low-frequency
    random texture plus noise, cropped twice with an (8,-8) px shift so that some true
    mutual matches exist.  Returns im1, im2 as [1,3,H,W] fp32 on CPU."""
    g = torch.Generator().manual_seed(1000 + int(pair_idx))
    low = torch.randn(1, 3, height // 8 + 4, width // 8 + 4, generator=g)
    base = F.interpolate(low, size=(height + 32, width + 32), mode='bicubic', align_corners=False)
    base = base + 0.3 * torch.randn(1, 3, height + 32, width + 32, generator=g)
    im1 = base[:, :, 16:16 + height, 16:16 + width].contiguous()
    im2 = base[:, :, 8:8 + height, 24:24 + width].contiguous()
    return im1, im2


def shifted_pair_offset(pair_idx):
    """(dx, dy) of synthetic_pair_shifted: multiples of 16 px, never (0, 0)."""
    g = torch.Generator().manual_seed(2000 + int(pair_idx))
    dx = 16 * int(torch.randint(-3, 4, (1,), generator=g))
    dy = 16 * int(torch.randint(-2, 3, (1,), generator=g))
    if dx == 0 and dy == 0:
        dx = 16
    return dx, dy


def synthetic_pair_shifted(pair_idx, height, width, noise=0.6):
    """Benchmark workload (round 2): two overlapping views of one texture whose offset is a multiple
    of 16 px (= one pooled correlation cell at ksize 2), so that coarse cells correspond one to one in
    the overlap, plus strong independent noise on the second view (std 0.6 against a unit-variance texture:
    with less noise the four true fine-level matches inside a 2^4 pooling window all have cosine
    1 - O(1e-6) and ~10 % of the reference's own relocalisation deltas are fp32 coin flips).  With the
    'consensus' NC weights this gives ~1000 distinct mutual matches at 640x480 (vs 13-17 for
    `synthetic_pair`), i.e. filter_coarse(ptmax=400) samples 400 DISTINCT proposals.
    Returns im1, im2 as [1,3,H,W] fp32 on CPU."""
    g = torch.Generator().manual_seed(2000 + int(pair_idx))
    dx = 16 * int(torch.randint(-3, 4, (1,), generator=g))
    dy = 16 * int(torch.randint(-2, 3, (1,), generator=g))
    if dx == 0 and dy == 0:
        dx = 16
    low = torch.randn(1, 3, height // 8 + 12, width // 8 + 12, generator=g)
    base = F.interpolate(low, size=(height + 96, width + 96), mode='bicubic', align_corners=False)
    base = base + 0.3 * torch.randn(1, 3, height + 96, width + 96, generator=g)
    im1 = base[:, :, 48:48 + height, 48:48 + width].contiguous()
    im2 = base[:, :, 48 + dy:48 + dy + height, 48 + dx:48 + dx + width].contiguous()
    im2 = im2 + noise * torch.randn(im2.shape, generator=g)
    return im1, im2
