"""Run-to-run determinism of the refine stage at the benchmark size, per conv1 variant (fuse_gather 3 / 1 / 0):
the same anchors through mid + fine several times; reports how many rows differ between repetitions."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import model_config  # noqa: E402
from patch2pix_b200.model import Patch2PixB200  # noqa: E402
from patch2pix_b200.synth import make_seeded_state_dict, synthetic_pair_shifted  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
cfg = model_config(torch.device('cuda:0'), 8)
cfg.weights_dict = make_seeded_state_dict(0, nc_init='consensus')
net = Patch2PixB200(cfg)
rep = {}
with torch.no_grad():
    im1, im2 = synthetic_pair_shifted(3, 480, 640)
    f1, f2 = net.extract_pair(im1.cuda(), im2.cuda())
    np.random.seed(11)
    base = net.match_from_feats(f1, f2, 2, ptmax=400, return_all=True)
    anch = base[4]
    for fg in (3, 1, 0):
        net.set_option('fuse_gather', fg)
        for band in (26, 0):
            net.set_option('mid_band', band)
            outs = []
            for rep_i in range(6):
                mid, midp = net.forward_fine_match(f1, f2, anch, 16, 'center', net.regress_mid)
                fine, finep = net.forward_fine_match(f1, f2, mid, 16, 'center', net.regress_fine)
                torch.cuda.synchronize()
                outs.append((mid[0].clone(), fine[0].clone()))
            dm = [int(((o[0] != outs[0][0]).any(1)).sum()) for o in outs[1:]]
            df = [int(((o[1] != outs[0][1]).any(1)).sum()) for o in outs[1:]]
            mx = max(float((o[1] - outs[0][1]).abs().max()) for o in outs[1:])
            rep[f'fuse_gather{fg}_band{band}'] = {'mid_rows_differing': dm, 'fine_rows_differing': df, 'max_fine_diff_px': mx}
            print(f'fuse_gather {fg} band {band}: mid rows differing {dm} fine rows differing {df} max diff {mx:.3e}', flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(rep, open(os.path.join(ROOT, 'gpurun_out', 'determinism_probe.json'), 'w'), indent=1)
