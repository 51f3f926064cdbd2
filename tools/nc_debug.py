"""Fault isolation for the tensor-core NeighConsensus (nc_umma.cu): compares xp, the hidden tensor and the partial
maps of one p2p_neigh_consensus call with fp64 torch restatements.  Needs a GPU.
    python tools/nc_debug.py [hA wA hB wB] [weights: uniform|consensus] [nc_l2_mode]"""
import ctypes as C
import os
import sys
from argparse import Namespace

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import p2p_oracle as O  # noqa: E402
from patch2pix_b200 import _lib  # noqa: E402
from patch2pix_b200.model import Patch2PixB200  # noqa: E402
from patch2pix_b200.synth import make_seeded_state_dict  # noqa: E402


def run(net, sd, dims, mode, verbose):
    hA, wA, hB, wB = dims
    net.set_option('nc_l2_mode', mode)
    h = net._ready()
    lib = h.lib
    lib.p2p_debug_nc_scratch.restype = C.c_int
    lib.p2p_debug_nc_scratch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    g = torch.Generator().manual_seed(hA * 100 + wB)
    x = torch.rand(1, 1, hA, wA, hB, wB, generator=g) - 0.1
    xd = x.cuda()
    out = torch.empty_like(xd)
    _lib.check(lib.p2p_neigh_consensus(h.h, _lib.ptr(xd), hA, wA, hB, wB, _lib.ptr(out), h.stream()))
    torch.cuda.synchronize()
    nA, nB = hA * wA, hB * wB
    V = nA * nB
    WP = (wB + 32 + 3) & ~3

    def grab(which, dtype, n):
        a = np.empty(n, dtype=dtype)
        _lib.check(lib.p2p_debug_nc_scratch(h.h, which, a.ctypes.data_as(C.c_void_p), a.nbytes))
        return a

    xmax = grab(3, np.float32, 1)[0]
    e = int(np.frexp(xmax)[1])
    sx = 2.0 ** (12 - e)
    xp = grab(2, np.uint32, (hA + 2) * (wA + 2) * (hB + 2) * WP).reshape(hA + 2, wA + 2, hB + 2, WP)
    hi = (xp & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float64)
    lo = (xp >> 16).astype(np.uint16).view(np.float16).astype(np.float64)
    rec = (hi + lo) / sx
    want = np.zeros_like(rec)
    want[1:hA + 1, 1:wA + 1, 1:hB + 1, 1:wB + 1] = x[0, 0].double().numpy()
    xp_err = np.abs(rec - want).max()
    w1, b1 = sd['ncn.conv.0.weight'].double(), sd['ncn.conv.0.bias'].double()
    w2 = sd['ncn.conv.2.weight'].double()
    xd64 = x.double()
    h0 = F.relu(O.conv4d(xd64, w1, b1))[0]                                               # [16,hA,wA,hB,wB]
    h1 = F.relu(O.conv4d(xd64.permute(0, 1, 4, 5, 2, 3), w1, b1)).permute(0, 1, 4, 5, 2, 3)[0]
    wsum = float(np.float32(w1.permute(1, 0, 2, 3, 4, 5).reshape(16, -1).float().abs().sum(1).max().item()) * np.float32(1.0001))
    hb = np.float32(wsum) * np.float32(xmax) + np.float32(b1.abs().max().item())
    sh = 2.0 ** (12 - int(np.frexp(np.float32(hb))[1]))
    hid = grab(0, np.float16, V * 64).astype(np.float64).reshape(nA, nB, 2, 2, 16)       # [a][b][net][hi|lo][ch]
    got = (hid[:, :, :, 0] + hid[:, :, :, 1]) / sh                                       # [a][b][net][ch]
    exp = torch.stack([h0, h1], 0).reshape(2, 16, nA, nB).permute(2, 3, 0, 1).double().numpy()
    err = np.abs(got - exp)
    if verbose and err.max() > 1e-4 * max(exp.max(), 1e-9):
        print('   hidden: sh', sh, 'max |exp|', exp.max(), 'net0 err', err[:, :, 0].max(), 'net1 err', err[:, :, 1].max(),
              'hi-only err', np.abs(hid[:, :, :, 0] / sh - exp).max())
        for a, b, n, c in np.argwhere(err > 0.5 * err.max())[:6]:
            print('     bad hidden a', a, 'b', b, 'net', n, 'ch', c, 'got', got[a, b, n, c], 'exp', exp[a, b, n, c])
    part = grab(1, np.float32, 18 * V).astype(np.float64).reshape(2, 9, nA, hB, wB)
    hdev = torch.from_numpy(got).permute(2, 3, 0, 1).reshape(2, 16, nA, hB, wB)        # [net][ch][a][k][l]
    worst, scale = 0.0, 0.0
    for net_i in range(2):
        for d in range(9):
            if net_i == 0:       # partial (ta, tb) = d: kernel over (tk, tl)
                ker = w2[d // 3, 0, :, d % 3]                                           # [16,3,3]
            else:                # partial (tk, tl) = d: kernel over (ta, tb)
                ker = w2[:, 0, :, :, d // 3, d % 3].permute(1, 0, 2)                    # [16,ta,tb]
            e2 = F.conv2d(hdev[net_i].permute(1, 0, 2, 3), ker[None], padding=1)[:, 0].numpy()
            df = np.abs(part[net_i, d] - e2)
            worst = max(worst, df.max())
            scale = max(scale, np.abs(e2).max())
            if verbose and df.max() > 1e-4 * max(np.abs(e2).max(), 1e-9):
                a, k, l = np.unravel_index(df.argmax(), df.shape)
                print(f'     partial net {net_i} d {d}: max err {df.max():.4g} at a {a} k {k} l {l}: got {part[net_i, d, a, k, l]:.6g} '
                      f'exp {e2[a, k, l]:.6g} (max |exp| {np.abs(e2).max():.4g}); rows wrong {int((df > 1e-4 * np.abs(e2).max()).sum())}/{df.size}')
    ref = O.neigh_consensus(x, sd)
    print(f'dims {dims} mode {mode}: xp err {xp_err:.3g} | hidden err {err.max():.3g} of {exp.max():.3g} | partial err {worst:.3g} of '
          f'{scale:.3g} | output err {float((out.cpu() - ref).abs().max()):.3g} of {float(ref.abs().max()):.3g}', flush=True)


def main():
    weights = sys.argv[1] if len(sys.argv) > 1 else 'uniform'
    sd = make_seeded_state_dict(0, nc_init=weights)
    cfg = Namespace(training=False, device='cuda:0', regr_batch=1200, backbone='ResNet34', feat_idx=[0, 1, 2, 3],
                    weights_dict=sd, change_stride=True,
                    regressor_config=Namespace(conv_dims=[512, 512], conv_kers=[3, 3], conv_strs=[2, 1], fc_dims=[512, 256],
                                               feat_comb='pre', psize=[16, 16], pshift=8, panc=1, shared=False))
    net = Patch2PixB200(cfg)
    first = True
    for dims in ((3, 4, 5, 6), (2, 3, 6, 8), (3, 2, 30, 40), (2, 3, 12, 64), (2, 2, 3, 200), (3, 2, 45, 37)):
        for mode in (1, 2):
            run(net, sd, dims, mode, verbose=first or dims == (3, 2, 30, 40))
        first = False


if __name__ == '__main__':
    main()
