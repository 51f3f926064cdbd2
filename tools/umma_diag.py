"""Bring-up diagnostic for the tcgen05 GEMM: structured operands whose product reveals any
row / column / K permutation, written to gpurun_out/ for offline inspection."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from patch2pix_b200 import _lib  # noqa: E402

out = os.path.join(ROOT, 'gpurun_out')
os.makedirs(out, exist_ok=True)
h = _lib.default_handle('cuda:0')
rep = {}
for (M, N, K) in [(128, 256, 64), (128, 256, 128), (256, 512, 256)]:
    a = torch.zeros(M, K)
    a[torch.arange(M), torch.arange(M) % K] = 1.0            # C[i][j] = B[j][i % K]
    b = (torch.arange(N).view(-1, 1) * 64 + torch.arange(K).view(1, -1)).float() / 64.0
    ref = a @ b.t()
    ad, bd = a.cuda(), b.cuda()
    for passes, seg in ((1, 0), (3, 0), (3, 1)):
        c = torch.full((M, N), -7.0, device='cuda')
        rc = h.lib.p2p_test_gemm(h.h, _lib.ptr(ad), _lib.ptr(bd), _lib.ptr(c), M, N, K, passes, seg, 8.0, h.stream())
        key = f'{M}x{N}x{K}_p{passes}_s{seg}'
        if rc != 0:
            rep[key] = {'rc': rc, 'err': h.lib.p2p_last_error().decode()}
            continue
        try:
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            rep[key] = {'sync_error': str(e)[:300]}
            print(json.dumps(rep, indent=1))
            json.dump(rep, open(os.path.join(out, 'umma_diag.json'), 'w'), indent=1)
            sys.exit(1)
        cc = c.cpu()
        err = (cc - ref).abs().max().item()
        rep[key] = {'max_abs_err': err, 'untouched': int((cc == -7.0).sum()), 'nan': int(torch.isnan(cc).sum())}
        if err > 1e-3:
            np.save(os.path.join(out, f'umma_diag_{key}.npy'), cc.numpy()[:160, :320])
print(json.dumps(rep, indent=1))
json.dump(rep, open(os.path.join(out, 'umma_diag.json'), 'w'), indent=1)
