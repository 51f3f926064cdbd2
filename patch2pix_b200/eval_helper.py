"""Tensor-level mirror of the reference's matcher glue (utils/eval/model_helper.py:28-109).

Image decoding / resizing (`load_im_flexible`, PIL) is host I/O and out of scope:
these helpers start from already normalised image tensors [1,3,H,W] plus the (sx, sy) scale factors the
reference's loader would have returned, and reproduce the rest of `estimate_matches` exactly:
`predict_coarse` / `predict_fine`, the `io_thres` inlier filter with its "keep everything if nothing
passes" rule, and the rescaling of the matches to original-image pixels.
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


def estimate_matches(net, im1, im2, scale1=(1.0, 1.0), scale2=(1.0, 1.0), ksize=2, ncn_thres=0.0, mutual=True,
                     io_thres=0.25, eval_type='fine'):
    """utils/eval/model_helper.py:64-109 on image tensors -> (matches, scores, coarse_matches) numpy arrays."""
    upscale = np.array([tuple(scale1) + tuple(scale2)])
    im1 = im1.to(net.device)
    im2 = im2.to(net.device)
    with torch.no_grad():
        if eval_type == 'coarse':
            coarse_matches, scores = net.predict_coarse(im1, im2, ksize=ksize, ncn_thres=ncn_thres, mutual=mutual)
            matches = upscale * coarse_matches[0].cpu().data.numpy()
            return matches, scores[0].cpu().data.numpy(), matches
        fine_matches, fine_scores, coarse_matches = net.predict_fine(im1, im2, ksize=ksize, ncn_thres=ncn_thres,
                                                                    mutual=mutual)
    coarse_matches = coarse_matches[0].cpu().data.numpy()
    fine_matches = fine_matches[0].cpu().data.numpy().reshape(-1, 4)
    fine_scores = fine_scores[0].cpu().data.numpy().reshape(-1)
    pos_ids = np.where(fine_scores > io_thres)[0]
    if len(pos_ids) > 0:
        coarse_matches, matches, scores = coarse_matches[pos_ids], fine_matches[pos_ids], fine_scores[pos_ids]
    else:
        matches, scores = fine_matches, fine_scores
    return upscale * matches, scores, upscale * coarse_matches
