"""Mirror of the reference's matcher glue (utils/eval/model_helper.py:28-109), device-resident.

`estimate_matches` starts from normalised image tensors [1,3,H,W] plus the (sx, sy) scale factors of the loader and
reproduces the reference exactly -- `predict_coarse` / `predict_fine`, the `io_thres` inlier filter with its "keep
everything if nothing passes" rule, the rescaling to original-image pixels in float64 -- but filter and rescaling run
in one kernel and the result crosses PCIe in ONE copy.  `estimate_matches_from_files` adds the loader: only the
file-format decode stays on the host, resize / ToTensor / Normalize run on the GPU (patch2pix_b200.preprocess).
"""
from argparse import Namespace

import numpy as np
import torch

from .model import Patch2PixB200


def load_model(state_dict, regressor_config=None, device='cuda:0', method='patch2pix'):
    """utils/eval/model_helper.py:28-62 without the checkpoint file I/O: `state_dict` is the loaded
    `ckpt['state_dict']`; `regressor_config` the checkpoint's Namespace (panc is forced to 1 as in :46)."""
    config = Namespace(training=False, device=torch.device(device), regr_batch=1200, backbone='ResNet34',
                       feat_idx=None, weights_dict=state_dict, regressor_config=None, change_stride=True)
    if 'patch2pix' in method:
        if regressor_config is None:
            regressor_config = Namespace(conv_dims=[512, 512], conv_kers=[3, 3], conv_strs=[2, 1], fc_dims=[512, 256],
                                         feat_comb='pre', psize=[16, 16], pshift=8, panc=1, shared=False)
        config.feat_idx = [0, 1, 2, 3]
        config.regressor_config = regressor_config
        config.regressor_config.panc = 1
    return Patch2PixB200(config)


def load_checkpoint(ckpt_path, device='cuda:0', method='patch2pix', lprint=print):
    """utils/eval/model_helper.py:28-62 + utils/common/setup_helper.py:25-30 including the file I/O.

    Released checkpoints are pickled dicts {'backbone', 'feat_idx', 'state_dict', 'regressor_config': Namespace,
    ['last_epoch']}; torch >= 2.6 refuses the Namespace under `weights_only=True`, so the file is read with
    `weights_only=False` (SURVEY.md s8c shim 3) onto the CPU -- `pack_weights` does the one upload. An 'nc'
    checkpoint may be a bare state_dict (:54-57). Only the released architecture is accepted."""
    ckpt = torch.load(ckpt_path, map_location='cpu', weights_only=False)
    lprint('\nLoad model method:{} '.format(method))
    if 'patch2pix' in method:
        if ckpt.get('backbone', 'ResNet34') != 'ResNet34' or list(ckpt.get('feat_idx', [0, 1, 2, 3])) != [0, 1, 2, 3]:
            raise RuntimeError('only the released ResNet34 / feat_idx [0,1,2,3] configuration is supported, got '
                               f"{ckpt.get('backbone')} / {ckpt.get('feat_idx')}")
        if 'last_epoch' in ckpt:
            lprint(f"Ckpt:{ckpt_path} epochs:{ckpt['last_epoch'] + 1}")
        else:
            lprint(f'Ckpt:{ckpt_path}')
        return load_model(ckpt['state_dict'], ckpt.get('regressor_config'), device=device, method=method)
    if 'nc' in method:
        sd = ckpt['state_dict'] if isinstance(ckpt, dict) and 'state_dict' in ckpt else ckpt
        lprint('Load pretrained weights: {}'.format(ckpt_path))
        return load_model(sd, None, device=device, method=method)
    raise ValueError('Wrong method name.')


def _finalize(net, fine, scores, coarse, io_thres, upscale):
    """One launch + ONE device->host copy for the tail of estimate_matches (model_helper.py:97-109)."""
    import ctypes as C
    from . import _lib
    h = net._handle
    n = int(scores.shape[0])
    dev = scores.device
    packed = torch.empty(n * 9 + 1, dtype=torch.float64, device=dev)
    up = (C.c_double * 4)(*[float(v) for v in upscale])
    fine_c = fine.reshape(-1, 4).contiguous() if fine is not None else None
    scores_c = scores.reshape(-1).contiguous()
    coarse_c = coarse.contiguous()
    with torch.cuda.device(dev):
        _lib.check(h.lib.p2p_finalize_matches(h.h, _lib.ptr(fine_c), _lib.ptr(scores_c), _lib.ptr(coarse_c), n, float(io_thres),
                                              up, _lib.ptr(packed), h.stream()))
    host = packed.cpu().numpy()                      # the single synchronising copy
    m = int(host[-1])
    rows = host[:-1].reshape(n, 9)[:m]
    return rows[:, 0:4].copy(), rows[:, 4].astype(np.float32), rows[:, 5:9].copy()


def estimate_matches(net, im1, im2, scale1=(1.0, 1.0), scale2=(1.0, 1.0), ksize=2, ncn_thres=0.0, mutual=True,
                     io_thres=0.25, eval_type='fine'):
    """utils/eval/model_helper.py:64-109 on image tensors -> (matches, scores, coarse_matches) numpy arrays
    (float64 matches in original-image pixels, float32 scores), with the inlier filter and the rescaling on the device
    and a single device->host copy (the reference does three `.cpu()` round trips)."""
    upscale = tuple(scale1) + tuple(scale2)
    im1 = im1.to(net.device)
    im2 = im2.to(net.device)
    with torch.no_grad():
        if eval_type == 'coarse':
            coarse_matches, scores = net.predict_coarse(im1, im2, ksize=ksize, ncn_thres=ncn_thres, mutual=mutual)
            m, s, _ = _finalize(net, None, scores[0], coarse_matches[0], float('-inf'), upscale)
            return m, s, m
        if eval_type != 'fine':
            raise ValueError("eval_type must be 'coarse' or 'fine'")
        fine_matches, fine_scores, coarse_matches = net.predict_fine(im1, im2, ksize=ksize, ncn_thres=ncn_thres,
                                                                    mutual=mutual)
    return _finalize(net, fine_matches[0], fine_scores[0], coarse_matches[0], io_thres, upscale)


def estimate_matches_from_files(net, im1_path, im2_path, ksize=2, ncn_thres=0.0, mutual=True, io_thres=0.25,
                                eval_type='fine', imsize=None):
    """utils/eval/model_helper.py:64-72 + the above: image files in, numpy matches out.  Only the file decode runs on
    the host; resize / ToTensor / Normalize are GPU kernels (patch2pix_b200.preprocess)."""
    from .preprocess import load_im_flexible
    im1, sc1 = load_im_flexible(im1_path, ksize, net.upsample, imsize=imsize, device=net.device, handle=net._handle)
    im2, sc2 = load_im_flexible(im2_path, ksize, net.upsample, imsize=imsize, device=net.device, handle=net._handle)
    return estimate_matches(net, im1.unsqueeze(0), im2.unsqueeze(0), sc1, sc2, ksize, ncn_thres, mutual, io_thres, eval_type)
