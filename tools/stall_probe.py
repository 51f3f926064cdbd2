"""Where do the sporadic 20-70 ms steps of the hot loop come from?  Re-runs bench.py's hot loop with host timestamps
around submit / finish, and prints allocator and cgroup counters before / after.  Needs a GPU."""
import gc
import json
import os
import sys
import time
from argparse import Namespace
from collections import deque

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from patch2pix_b200.model import Patch2PixB200  # noqa: E402


def cg():
    out = {}
    for f in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu.stat', '/proc/loadavg', '/proc/pressure/cpu'):
        try:
            out[f] = open(f).read().strip().replace('\n', ' | ')
        except Exception as e:
            out[f] = repr(e)
    return out


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    use_gc = len(sys.argv) > 2 and sys.argv[2] == 'gc'
    use_sampler = len(sys.argv) > 2 and sys.argv[2] == 'sampler'
    args = Namespace(height=480, width=640, ptmax=400, legacy_workload=False)
    dev = torch.device('cuda', 0)
    sd, gen = bench.make_workload(args)
    cfg = bench.model_config(dev, 8)
    cfg.weights_dict = sd
    net = Patch2PixB200(cfg)
    feats = []
    with torch.no_grad():
        for k in range(8):
            a, b = gen(k, 480, 640)
            feats.append((net.extract.forward_all(a.to(dev), [], True), net.extract.forward_all(b.to(dev), [], True)))
        results = torch.zeros(steps, 3200, 5, device=dev)
        print('cpus', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)), cg(), flush=True)

        def run(n, record):
            q = deque()
            rows = []
            for j in range(n):
                t0 = time.perf_counter()
                tk = net.submit_coarse(*feats[j % 8], 2, True)
                t1 = time.perf_counter()
                q.append((j, tk, t0, t1))
                if len(q) >= 3:
                    i, tk, a0, a1 = q.popleft()
                    np.random.seed(i)
                    t2 = time.perf_counter()
                    fine, fine_p, cm = net.finish_match(tk, 0.0, 400)
                    t3 = time.perf_counter()
                    if record:
                        results[i, :, :4] = fine[0]
                        results[i, :, 4] = fine_p[0]
                    t4 = time.perf_counter()
                    rows.append((i, a1 - a0, t3 - t2, t4 - t3, t4))
            while q:
                i, tk, a0, a1 = q.popleft()
                np.random.seed(i)
                net.finish_match(tk, 0.0, 400)
            return rows
        run(8, False)
        torch.cuda.synchronize()
        ms0 = {k: v for k, v in torch.cuda.memory_stats().items() if k in ('num_device_alloc', 'num_alloc_retries', 'num_device_free', 'segment.all.allocated')}
        try:
            hs0 = dict(torch.cuda.host_memory_stats())
        except Exception:
            hs0 = {}
        if not use_gc:
            gc.collect()
            gc.disable()
        c0 = cg()
        sampler = bench.ClockSampler(0) if use_sampler else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rows = run(steps, True)
        e1.record()
        torch.cuda.synchronize()
        gc.enable()
        if sampler:
            print('sampler', sampler.finish())
        c1 = cg()
        ms1 = {k: v for k, v in torch.cuda.memory_stats().items() if k in ms0}
        try:
            hs1 = dict(torch.cuda.host_memory_stats())
        except Exception:
            hs1 = {}
        print('ms/step', e0.elapsed_time(e1) / steps, 'gc' if use_gc else 'nogc')
        print('device allocator before/after', ms0, ms1)
        print('host allocator keys changed', {k: (hs0.get(k), v) for k, v in hs1.items() if hs0.get(k) != v and ('alloc' in k or 'segment' in k)})
        print('cgroup before', c0)
        print('cgroup after ', c1)
        per = [(rows[i][4] - rows[i - 1][4]) * 1e3 for i in range(1, len(rows))]
        order = sorted(range(len(per)), key=lambda i: -per[i])[:8]
        for i in order:
            r = rows[i + 1]
            print(f'step {r[0]}: host period {per[i]:.2f} ms  submit {r[1] * 1e3:.2f}  finish {r[2] * 1e3:.2f}  store {r[3] * 1e3:.2f}')
        print('host period quantiles ms', np.percentile(per, [10, 50, 90, 99, 100]).round(2).tolist())


if __name__ == '__main__':
    main()
