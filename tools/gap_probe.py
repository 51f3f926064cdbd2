"""Where do launch gaps come from?  Hot loop (features resident) for 60 steps under: profile events on/off x NVML
sampler process on/off x pipeline depth 2/3; prints ms/step of each (3 repeats) and the host time per step."""
import json
import os
import sys
import time
from collections import deque

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from patch2pix_b200.model import Patch2PixB200  # noqa: E402
from patch2pix_b200.synth import make_seeded_state_dict, synthetic_pair_shifted  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
cfg = bench.model_config(torch.device('cuda:0'), 8)
cfg.weights_dict = make_seeded_state_dict(0, nc_init='consensus')
net = Patch2PixB200(cfg)
feats = []
with torch.no_grad():
    for p in range(8):
        a, b = synthetic_pair_shifted(p, 480, 640)
        feats.append((net.extract.forward_all(a.cuda(), [], True), net.extract.forward_all(b.cuda(), [], True)))


def loop(steps, depth):
    q = deque()
    host = 0.0
    for j in range(steps):
        t0 = time.perf_counter()
        q.append((j, net.submit_coarse(*feats[j % 8], 2, True)))
        if len(q) >= depth:
            i, tk = q.popleft()
            np.random.seed(i)
            net.finish_match(tk, 0.0, 400)
        host += time.perf_counter() - t0
    while q:
        i, tk = q.popleft()
        np.random.seed(i)
        net.finish_match(tk, 0.0, 400)
    return host / steps * 1e3


rep = []
with torch.no_grad():
    loop(10, 3)
    torch.cuda.synchronize()
    for prof in (0, 1):
        for samp in (0, 1):
            for depth in (2, 3):
                for r in range(3):
                    net.set_option('profile', prof)
                    s = bench.ClockSampler(0) if samp else None
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    if s:
                        s.start()
                    e0.record()
                    host_ms = loop(60, depth)
                    e1.record()
                    torch.cuda.synchronize()
                    if s:
                        s.finish()
                    if prof:
                        net._handle.profile_read()
                    ms = e0.elapsed_time(e1) / 60
                    rep.append({'profile': prof, 'sampler': samp, 'depth': depth, 'ms_per_step': ms, 'host_ms_per_step': host_ms})
                    print(rep[-1], flush=True)
json.dump(rep, open(os.path.join(ROOT, 'gpurun_out', 'gap_probe.json'), 'w'), indent=1)
