#!/usr/bin/env python
"""Benchmark: image-pairs/sec of the Patch2Pix correlate-and-refine hot path at 640x480,
ptmax=400, panc=8 (BASELINE.json configs[2]; training-loop forward sequence under eval,
train_patch2pix.py:97-118), on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm (CPU oracle port) on host cores

One "step" = one image pair per GPU through the whole hot path (weak scaling: pair p of step s
goes to rank p % N; no data-path collective, NCCL only broadcasts the pair indices and gathers
the matches).  `value` times the hot path with the feature pyramids already in HBM;
`e2e` times pinned-host images -> H2D -> cuDNN backbone -> hot path -> D2H of the matches.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H_DEF, W_DEF, PTMAX_DEF, PANC_DEF = 480, 640, 400, 8
MAC_CONV1, MAC_CONV2 = 152764416, 150994944        # per patch, dense count


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--height', type=int, default=H_DEF)
    ap.add_argument('--width', type=int, default=W_DEF)
    ap.add_argument('--ptmax', type=int, default=PTMAX_DEF)
    ap.add_argument('--mid-passes', type=int, default=None)
    ap.add_argument('--fine-passes', type=int, default=None)
    ap.add_argument('--corr-passes', type=int, default=None)
    ap.add_argument('--seg-len', type=int, default=None)
    ap.add_argument('--mid-band', type=int, default=None)
    ap.add_argument('--fuse-gather', type=int, default=None)
    ap.add_argument('--gemm-pair', type=int, default=None, help='bitmask of GEMM launches on the CTA-pair kernel')
    ap.add_argument('--backbone-fp32', action='store_true', help='keep cuDNN TF32 off in the e2e backbone')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-sample-patches', type=int, default=200)
    return ap.parse_args()


def model_config(device, panc):
    from argparse import Namespace
    rc = Namespace(conv_dims=[512, 512], conv_kers=[3, 3], conv_strs=[2, 1], fc_dims=[512, 256], feat_comb='pre',
                   psize=[16, 16], pshift=8, panc=panc, shared=False)
    return Namespace(training=False, device=device, regr_batch=1200, backbone='ResNet34', feat_idx=[0, 1, 2, 3],
                     weights_dict=None, change_stride=True, regressor_config=rc)


def load_traffic():
    """DRAM bytes (read + write) per launch from the committed `ncu --set full` capture of the round
    (profiles/r01_traffic.json: kernel kind -> bytes), or {}."""
    p = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {'hbm_gbs': d['hbm_gbs'], 'tflops': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'src': 'measured'}
    return {'hbm_gbs': 6650.0, 'tflops': 1400.0, 'src': 'fallback'}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""

    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(',')]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._halt.wait(0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=6)
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = [float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': float(self.rows[0][1]),
                'power_w_max': max(float(r[2]) for r in self.rows), 'samples': len(self.rows), 'reasons': reasons}


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm restated in oracle/p2p_oracle.py (the Python reference itself
# cannot travel to the GPU box).  One step = the full coarse stage of one pair + the two refine
# stages on a bounded subset of the 3200 patches, extrapolated to the full pair.
# --------------------------------------------------------------------------------------------------
def cpu_threads():
    """Threads for the CPU arm: every host core up to 32 (torch's CPU kernels for this path -- hundreds of
    small conv3d / index ops -- get slower, not faster, beyond that; measured on the 128-core GPU box).
    Override with P2P_CPU_THREADS."""
    cores = os.cpu_count() or 1
    return int(os.environ.get('P2P_CPU_THREADS', min(cores, 32)))


def cpu_step(O, sd, im1, im2, ptmax, panc, n_sample, nc_slices=None):
    """One bounded sample of the reference algorithm on the CPU for one pair: full backbone, full
    correlation / max-pool / mutual matching / proposals, the NC 4D conv on `nc_slices` of its
    first-dimension output slices (all if None) and the two refine stages on `n_sample` patches;
    the sampled parts are scaled to the full pair."""
    t0 = time.perf_counter()
    with torch.no_grad():
        f1 = O.backbone_forward_all(im1, sd)
        f2 = O.backbone_forward_all(im2, sd)
        t1 = time.perf_counter()
        a, b = O.l2_normalize(f1[-1], 1), O.l2_normalize(f2[-1], 1)
        corr, mi, mj, mk, ml = O.maxpool4d(O.feat_correlation_4d(a, b), 2)
        corr = O.mutual_matching(corr)
        t2 = time.perf_counter()
        hA = corr.shape[2]
        sl = None if (nc_slices is None or nc_slices >= hA) else list(range(0, hA, max(1, hA // nc_slices)))[:nc_slices]
        nc = O.neigh_consensus(corr, sd, sl)
        t3 = time.perf_counter()
        nc_scale = 1.0 if sl is None else hA / len(sl)
        corr4d = O.mutual_matching(nc if sl is None else corr)
        cm, sc = O.cal_coarse_matches(corr4d, (mi, mj, mk, ml), 2, upsample=O.UPSAMPLE, center=True)
        np.random.seed(0)
        cm, sc = O.filter_coarse(cm, sc, 0.0, True, ptmax=ptmax)
        anch = O.shift_to_anchors(cm, panc)
        t4 = time.perf_counter()
        n_full = anch[0].shape[0]
        sub = [anch[0][:n_sample]]
        mid, _ = O.forward_fine_match(f1, f2, sub, sd, 'regress_mid.')
        fine, _ = O.forward_fine_match(f1, f2, mid, sd, 'regress_fine.')
        t5 = time.perf_counter()
    n_sub = sub[0].shape[0]
    t_nc = (t3 - t2) * nc_scale
    t_refine = (t5 - t4) * n_full / max(n_sub, 1)
    hot = (t2 - t1) + t_nc + (t4 - t3) + t_refine
    return {'backbone_s': t1 - t0, 'coarse_s': (t2 - t1) + t_nc + (t4 - t3), 'nc_s_extrapolated': t_nc,
            'refine_s_extrapolated': t_refine, 'hot_path_s': hot, 'e2e_s': (t1 - t0) + hot, 'wall_s': t5 - t0,
            'n_sample': n_sub, 'n_full': n_full, 'nc_slices': 'all' if sl is None else f'{len(sl)}/{hA}'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import p2p_oracle as O
    from patch2pix_b200.synth import make_seeded_state_dict, synthetic_pair
    threads = cpu_threads()
    torch.set_num_threads(threads)
    sd = make_seeded_state_dict(0)
    H, W = args.height, args.width
    pairs = [synthetic_pair(p, H, W) for p in range(2)]
    # size the per-step sample so that the whole run stays within ~3 minutes
    budget = 150.0 / max(args.steps + min(args.warmup, 1), 1)
    nc_slices, n_sample = (None, args.cpu_sample_patches) if budget > 12 else ((8, 96) if budget > 4 else (3, 32))
    for i in range(min(args.warmup, 1)):
        cpu_step(O, sd, *pairs[i % 2], args.ptmax, PANC_DEF, n_sample, nc_slices)
    rs = [cpu_step(O, sd, *pairs[i % 2], args.ptmax, PANC_DEF, n_sample, nc_slices) for i in range(args.steps)]
    hot = sum(r['hot_path_s'] for r in rs) / len(rs)
    e2e = sum(r['e2e_s'] for r in rs) / len(rs)
    sample = (f'per step, one {W}x{H} pair: full backbone + correlation/max-pool/mutual/proposals, NC 4D conv on '
              f'{rs[0]["nc_slices"]} output slices, mid+fine refine on {rs[0]["n_sample"]} of {rs[0]["n_full"]} patches; '
              f'sampled parts scaled to the full pair (mean wall {sum(r["wall_s"] for r in rs) / len(rs):.2f} s/step)')
    line = {'impl': 'reference', 'metric': 'image-pairs/sec', 'value': 1.0 / e2e, 'unit': 'pairs/s', 'n_gpus': 0,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': e2e * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'{W}x{H} pair, ptmax={args.ptmax} panc={PANC_DEF} (BASELINE configs[2])',
                       'sequence': 'train_patch2pix.py:97-118 under eval/no_grad', 'includes_backbone': True},
            'cpu_baseline': {'value': 1.0 / e2e, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port', 'sample': sample,
                             'host_cores': os.cpu_count(), 'hot_path_only_pairs_per_s': 1.0 / hot},
            'e2e': {'value': 1.0 / e2e, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from patch2pix_b200.model import Patch2PixB200
    from patch2pix_b200.sharding import PairSharder
    from patch2pix_b200.synth import make_seeded_state_dict, synthetic_pair

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py (impl ours) needs a CUDA device: there is no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    sharder = PairSharder(rank, world, dev)

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    H, W, K, Wm = args.height, args.width, args.steps, args.warmup
    cfg = model_config(dev, PANC_DEF)
    cfg.weights_dict = make_seeded_state_dict(0)
    net = Patch2PixB200(cfg)
    for key, v in (('mid_passes', args.mid_passes), ('fine_passes', args.fine_passes), ('corr_passes', args.corr_passes),
                   ('seg_len', args.seg_len), ('mid_band', args.mid_band), ('fuse_gather', args.fuse_gather),
                   ('gemm_pair', args.gemm_pair)):
        if v is not None:
            net.set_option(key, v)
    opts = {k: net._handle.get_option(k) for k in ('mid_passes', 'fine_passes', 'corr_passes', 'seg_len', 'mid_band', 'fuse_gather',
                                                    'gemm_pair')}

    # pair indices: rank 0 decides, NCCL broadcasts (the "scatter pair indices" step)
    total_steps = K + Wm
    all_pairs = sharder.scatter_pair_indices(torch.arange(total_steps * world, dtype=torch.int64))
    n_distinct = 4                                      # distinct synthetic pairs cycled per rank
    imgs = [synthetic_pair(int(all_pairs[i % len(all_pairs)]) % 64, H, W) for i in range(n_distinct)]
    pinned = [(a.pin_memory(), b.pin_memory()) for a, b in imgs]
    with torch.no_grad():
        feats = []
        for a, b in imgs:
            f1 = net.extract.forward_all(a.to(dev), [], True)
            f2 = net.extract.forward_all(b.to(dev), [], True)
            feats.append((f1, f2))
    n_patches = args.ptmax * PANC_DEF
    results = torch.zeros(K, n_patches, 5, device=dev)

    # Two pairs are kept in flight: the coarse stage of pair i is enqueued before the host waits for the
    # mutual-match count of pair i-1 (filter_coarse's host sync), so the GPU never idles on that sync.
    def hot_submit(i):
        f1, f2 = feats[i % n_distinct]
        return (i, net.submit_coarse(f1, f2, 2, True))

    def hot_finish(tk, out_slot=None):
        i, ticket = tk
        np.random.seed(i)
        fine, fine_p, _ = net.finish_match(ticket, 0.0, args.ptmax)
        if out_slot is not None:
            results[out_slot, :, :4] = fine[0]
            results[out_slot, :, 4] = fine_p[0]

    def hot_loop(first, steps, record):
        prev = None
        for j in range(steps):
            tk = hot_submit(first + j)
            if prev is not None:
                hot_finish(prev[0], prev[1])
            prev = (tk, j if record else None)
        hot_finish(prev[0], prev[1])

    def e2e_submit(i):
        a, b = pinned[i % n_distinct]
        f1, f2 = net.extract_pair(a, b, slot=i)    # pinned host images: H2D into the graph's input, then the backbone
        return (i, net.submit_coarse(f1, f2, 2, True))

    def e2e_finish(tk, host_out):
        i, ticket = tk
        np.random.seed(i)
        fine, fine_p, _ = net.finish_match(ticket, 0.0, args.ptmax)
        host_out[:, :4].copy_(fine[0], non_blocking=True)          # D2H read of this step's result
        host_out[:, 4].copy_(fine_p[0], non_blocking=True)

    def e2e_loop(first, steps, host_outs):
        prev = None
        for j in range(steps):
            tk = e2e_submit(first + j)
            if prev is not None:
                e2e_finish(prev, host_outs[(j - 1) % len(host_outs)])
            prev = tk
        e2e_finish(prev, host_outs[(steps - 1) % len(host_outs)])
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(steps)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return ms.item()

    with torch.no_grad():
        # ---- hot path, features resident in HBM -------------------------------------------------
        hot_loop(0, Wm, False)
        sharder.gather_results(results)                  # warm-up of the collective (NCCL sets up channels lazily)
        net.set_option('profile', 1)
        net._handle.profile_read()
        l0 = net._handle.launch_count()
        sampler = ClockSampler(local) if rank == 0 else None

        def hot_region(steps):
            hot_loop(Wm, steps, True)
            sharder.gather_results(results)              # NCCL gather of the matches (inside the timed region)
        ms_hot = timed(hot_region, K, sampler)
        launches = net._handle.launch_count() - l0
        clocks = sampler.finish() if sampler else None
        prof = net._handle.profile_read()
        net.set_option('profile', 0)

        band_rows = net._handle.get_option('band_rows')
        # ---- end to end: pinned host images -> matches on the host ------------------------------
        # backbone in PyTorch's default cuDNN mode (TF32 convolutions allowed, as the reference would run)
        torch.backends.cudnn.allow_tf32 = not args.backbone_fp32
        torch.backends.cudnn.benchmark = True
        net.enable_backbone_graphs(H, W, instances=2)
        host_outs = [torch.empty(n_patches, 5).pin_memory() for _ in range(2)]
        e2e_loop(0, max(min(Wm, 3), 1), host_outs)

        def e2e_region(steps):
            e2e_loop(Wm, steps, host_outs)
        ms_e2e = timed(e2e_region, K)

    if rank == 0:
        peaks = load_peaks()
        pairs = K * world
        value = pairs / (ms_hot / 1e3)
        # dominant kernel: the conv implicit GEMMs of the refine stage
        kern = {k: {'ms_per_launch': v[0] / v[1], 'launches': v[1]} for k, v in prof.items() if v[1] > 0}
        banded = opts['mid_passes'] == 3 and opts['mid_band'] > 0
        macs = {'conv1': MAC_CONV1, 'conv2': MAC_CONV2}
        for k in list(kern):
            base, _, stage = k.partition('_')
            if base not in macs:
                continue
            rows = band_rows if stage == 'band' else n_patches
            ps = 3 if stage == 'band' else (opts['fine_passes'] if stage == 'fine' else (1 if banded else opts['mid_passes']))
            fl = 2.0 * macs[base] * rows
            kern[k].update({'rows': rows, 'tensor_passes': ps,
                            'algorithmic_tflops': fl / (kern[k]['ms_per_launch'] * 1e-3) / 1e12})
            kern[k]['issued_tflops'] = kern[k]['algorithmic_tflops'] * ps
        gemm_names = [k for k in kern if k.startswith('conv')]
        dom = max(gemm_names, key=lambda k: kern[k]['ms_per_launch'] * kern[k]['launches'], default=None)
        roofline = None
        if dom:
            ach = kern[dom]['algorithmic_tflops']
            gemm_ms = sum(kern[k]['ms_per_launch'] * kern[k]['launches'] for k in gemm_names)
            kname = 'umma_conv1_fused_kernel' if (dom.startswith('conv1') and not dom.endswith('band') and opts['fuse_gather'] == 1) \
                else 'umma_gemm_kernel'
            traffic = load_traffic().get(dom if kname == 'umma_gemm_kernel' else 'conv1_fused')
            roofline = {'kernel': f'{kname} ({dom})', 'bound': 'tensor', 'achieved': ach, 'peak': peaks['tflops'],
                        'unit': 'TFLOP/s', 'frac': ach / peaks['tflops'], 'traffic': traffic,
                        'traffic_unit': 'bytes of DRAM read+write per launch (ncu --set full, profiles/)',
                        'peak_source': peaks['src'] + ' bf16 sustained (fp16 runs at the same tensor rate)',
                        'tensor_passes': kern[dom]['tensor_passes'],
                        'issued_frac': kern[dom]['issued_tflops'] / peaks['tflops'],
                        'share_of_step': kern[dom]['ms_per_launch'] * kern[dom]['launches'] / ms_hot,
                        'all_umma_gemm_share_of_step': gemm_ms / ms_hot,
                        'band_rows_recomputed_3pass': band_rows if banded else None}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import p2p_oracle as O
            threads = cpu_threads()
            torch.set_num_threads(threads)
            sd_cpu = make_seeded_state_dict(0)
            cpu_step(O, sd_cpu, *imgs[0], args.ptmax, PANC_DEF, 16, 2)          # warm-up
            r = cpu_step(O, sd_cpu, *imgs[0], args.ptmax, PANC_DEF, args.cpu_sample_patches, None)
            cpu = {'value': 1.0 / r['hot_path_s'], 'unit': 'pairs/s', 'cores': threads, 'host_cores': os.cpu_count(),
                   'kind': 'port',
                   'sample': (f'oracle port of the reference, {threads} threads: full coarse stage of one {W}x{H} pair '
                              f'({r["coarse_s"]:.2f} s) + mid/fine refine on {r["n_sample"]} of {r["n_full"]} patches scaled to '
                              f'the full pair ({r["refine_s_extrapolated"]:.2f} s); backbone excluded ({r["backbone_s"]:.2f} s)'),
                   'with_backbone_pairs_per_s': 1.0 / r['e2e_s']}
        line = {
            'metric': 'image-pairs/sec', 'value': value, 'unit': 'pairs/s', 'n_gpus': world, 'steps': K, 'warmup': Wm,
            'ms_per_step': ms_hot / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': f'f16 tensor-core operands (mid: {opts["mid_passes"]}-pass hi/lo split'
                     f'{" on the risk band, 1-pass elsewhere" if opts["mid_band"] and opts["mid_passes"] == 3 else ""}, '
                     f'fine: {opts["fine_passes"]}-pass, correlation: {opts["corr_passes"]}-pass), f32 accumulate; f32 NC conv',
            'data': 'synthetic',
            'config': {'workload': f'{W}x{H} pair, ptmax={args.ptmax} panc={PANC_DEF} -> {n_patches} patches/stage '
                                   f'(BASELINE configs[2]); hot path = correlation .. fine matches, features resident in HBM',
                       'sequence': 'train_patch2pix.py:97-118 under eval/no_grad', 'pairs_per_step': world,
                       'l2': 'distinct pair per step, per-step working set (~3 GB) >> 126 MB L2',
                       'pipelining': 'two pairs in flight per GPU (coarse of pair i is enqueued before the host sync of pair i-1)',
                       'options': opts},
            'e2e': {'value': pairs / (ms_e2e / 1e3), 'unit': 'pairs/s', 'ms_per_step': ms_e2e / K,
                    'h2d_bytes_per_step': 2 * 3 * H * W * 4, 'd2h_bytes_per_step': n_patches * 5 * 4,
                    'path': 'pinned host images -> H2D -> cuDNN ResNet34 pyramid, both images as one batch, CUDA graph (' + ('fp32' if args.backbone_fp32 else 'TF32 convs, PyTorch default') + ') -> hot path -> D2H matches+scores'},
            'gpu_launches': launches, 'roofline': roofline, 'kernels': kern, 'clocks': clocks, 'cpu_baseline': cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
