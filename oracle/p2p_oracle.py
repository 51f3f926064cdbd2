"""CPU oracle for the Patch2Pix correlate -> NC-filter -> propose -> refine path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``patch2pix_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker
or the timed CPU baseline -- never as a product code path.

This is a functional (stateless) restatement, on torch CPU fp32, of the
algorithm the reference implements with nn.Modules.  Every function cites the
reference file:line (relative to the reference repo root) it follows.  It
deliberately uses the *same library primitives in the same order* as the
reference (bmm, 16-slice max, conv3d loop, softmax+max, np.unique, conv2d, ...)
so that (a) on the same CPU it reproduces the reference bit-for-bit -- which is
how it is pinned, see tests/golden/make_golden.py and tests/test_oracle_golden.py
-- and (b) its wall-clock is representative of the reference's CPU path.

Parity status: the reference ships NO tests or golden vectors, so
the oracle is pinned against outputs of the live reference generated in the
authoring container (tests/golden/*.npz, generator script committed).

Weights are passed as a flat dict with the reference's ``state_dict`` key names
(``ncn.conv.0.weight`` ... ``regress_mid.fc.6.bias``, ``extract.*``).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

UPSAMPLE = 8                        # networks/patch2pix.py:18-27 with change_stride=True
FEATS_DOWNSAMPLE = [1, 2, 2, 2, 1]  # networks/patch2pix.py:19,27
FEAT_IDX = [0, 1, 2, 3]             # released model (examples/visualize_matches.ipynb output)
PSIZE = 16
PSHIFT = 8
REGR_BATCH = 1200                   # utils/eval/model_helper.py:34
BN_EPS = 1e-5


# --------------------------------------------------------------------------
# backbone (feeds the path; networks/resnet.py:125-157, ResNet34, layer3 stride 1)
# --------------------------------------------------------------------------
def _bn(x, sd, name):
    return F.batch_norm(x, sd[name + '.running_mean'], sd[name + '.running_var'],
                        sd[name + '.weight'], sd[name + '.bias'], False, 0.0, BN_EPS)


def _basic_block(x, sd, name, stride):
    """networks/resnet.py:18-47."""
    out = F.conv2d(x, sd[name + '.conv1.weight'], None, stride, 1)
    out = F.relu(_bn(out, sd, name + '.bn1'))
    out = F.conv2d(out, sd[name + '.conv2.weight'], None, 1, 1)
    out = _bn(out, sd, name + '.bn2')
    if (name + '.downsample.0.weight') in sd:
        x = F.conv2d(x, sd[name + '.downsample.0.weight'], None, stride, 0)
        x = _bn(x, sd, name + '.downsample.1')
    return F.relu(out + x)


def backbone_forward_all(im, sd, prefix='extract.'):
    """ResNet34.forward_all(early_feat=True): networks/resnet.py:138-157.

    Returns [image, conv1-relu, layer1, layer2, layer3]; layer3 runs at stride 1
    (change_stride, networks/resnet.py:169-173).
    """
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    feats = [im]
    x = F.relu(_bn(F.conv2d(im, sd['conv1.weight'], None, 2, 3), sd, 'bn1'))
    feats.append(x)
    x = F.max_pool2d(x, 3, 2, 1)
    for lname, nblk, stride in (('layer1', 3, 1), ('layer2', 4, 2), ('layer3', 6, 1)):
        for i in range(nblk):
            x = _basic_block(x, sd, f'{lname}.{i}', stride if i == 0 else 1)
        feats.append(x)
    return feats


# --------------------------------------------------------------------------
# coarse stage
# --------------------------------------------------------------------------
def l2_normalize(feat, dim):
    """networks/modules.py:6 -- eps sits inside the square root."""
    return feat / torch.pow(torch.sum(torch.pow(feat, 2), dim=dim) + 1e-6, 0.5).unsqueeze(dim)


def feat_correlation_4d(feat1, feat2):
    """networks/modules.py:41-53 -- [b,c,h1,w1] x [b,c,h2,w2] -> [b,1,h1,w1,h2,w2]."""
    b, c, h1, w1 = feat1.shape
    _, _, h2, w2 = feat2.shape
    corr = torch.bmm(feat1.view(b, c, h1 * w1).transpose(1, 2), feat2.view(b, c, h2 * w2))
    return corr.view(b, h1, w1, h2, w2).unsqueeze(1)


def maxpool4d(corr, k):
    """networks/modules.py:11-34 -- k^4 max with the argmax split into 4 deltas.

    Slice order idx = ((i*k+j)*k+kk)*k+l, ties resolve to the lowest idx
    (torch.max returns the first maximal index).
    """
    slices = [corr[:, :, i::k, j::k, kk::k, l::k]
              for i in range(k) for j in range(k) for kk in range(k) for l in range(k)]
    val, idx = torch.max(torch.cat(slices, dim=1), dim=1, keepdim=True)
    max_l = torch.fmod(idx, k)
    max_k = torch.fmod(idx.sub(max_l).floor_divide(k), k)
    max_j = torch.fmod(idx.sub(max_l).floor_divide(k).sub(max_k).floor_divide(k), k)
    max_i = idx.sub(max_l).floor_divide(k).sub(max_k).floor_divide(k).sub(max_j).floor_divide(k)
    return val, max_i, max_j, max_k, max_l


def mutual_matching(corr):
    """networks/ncn/model.py:157-176."""
    b, _, s1, s2, s3, s4 = corr.shape
    cb = corr.view(b, s1 * s2, s3, s4)
    ca = corr.view(b, s1, s2, s3 * s4)
    cb_max, _ = torch.max(cb, dim=1, keepdim=True)
    ca_max, _ = torch.max(ca, dim=3, keepdim=True)
    eps = 1e-5
    cb = (cb / (cb_max + eps)).view(b, 1, s1, s2, s3, s4)
    ca = (ca / (ca_max + eps)).view(b, 1, s1, s2, s3, s4)
    return corr * (ca * cb)


def conv4d(x, w, bias, out_slices=None):
    """networks/ncn/conv4d.py:12-74 with pre-permuted filters [k1,Cout,Cin,k2,k3,k4]
    (conv4d.py:118-120): loop over the first spatial dim, three conv3d per slice,
    bias added with the centre tap only.  `out_slices` (timing samples only, bench.py's CPU
    baseline) restricts the loop to a subset of output slices; the other slices stay zero."""
    b, c, h, ww, d, t = x.shape
    xp = x.permute(2, 0, 1, 3, 4, 5).contiguous()
    pad = w.shape[0] // 2
    z = torch.zeros(pad, b, c, ww, d, t)
    xp = torch.cat((z, xp, z), 0)
    out = torch.zeros(h, b, w.shape[1], ww, d, t)
    for i in (range(h) if out_slices is None else out_slices):
        out[i] = F.conv3d(xp[i + pad], w[pad], bias=bias, stride=1, padding=pad)
        for p in range(1, pad + 1):
            out[i] = out[i] + F.conv3d(xp[i + pad - p], w[pad - p], bias=None, stride=1, padding=pad)
            out[i] = out[i] + F.conv3d(xp[i + pad + p], w[pad + p], bias=None, stride=1, padding=pad)
    return out.permute(1, 2, 0, 3, 4, 5).contiguous()


def neigh_consensus(x, sd, out_slices=None):
    """networks/ncn/model.py:124-155 (symmetric mode, 2 layers 1->16->1, ReLU after each)."""
    def net(y):
        y = F.relu(conv4d(y, sd['ncn.conv.0.weight'], sd['ncn.conv.0.bias'], out_slices))
        return F.relu(conv4d(y, sd['ncn.conv.2.weight'], sd['ncn.conv.2.bias'], out_slices))
    return net(x) + net(x.permute(0, 1, 4, 5, 2, 3)).permute(0, 1, 4, 5, 2, 3)


def forward_coarse_match(feat1, feat2, sd, ksize=1, stages=None):
    """networks/patch2pix.py:120-136.  ``stages`` (optional dict) receives intermediates."""
    f1 = l2_normalize(feat1, 1)
    f2 = l2_normalize(feat2, 1)
    corr = feat_correlation_4d(f1, f2)
    delta4d = None
    if ksize > 1:
        corr, mi, mj, mk, ml = maxpool4d(corr, ksize)
        delta4d = (mi, mj, mk, ml)
    if stages is not None:
        stages['pooled'] = corr
    corr = mutual_matching(corr)
    if stages is not None:
        stages['mutual1'] = corr
    corr = neigh_consensus(corr, sd)
    if stages is not None:
        stages['ncn'] = corr
    corr = mutual_matching(corr)
    return corr, delta4d


def corr_to_matches(corr4d, delta4d=None, ksize=1, do_softmax=True, invert=False):
    """networks/ncn/extract_ncmatches.py:6-94 (return_indices=True branch).

    invert=False: softmax over A (dim 1 of [b,hA*wA,hB,wB]) -> best A for every B cell.
    invert=True : softmax over B (dim 3 of [b,hA,wA,hB*wB]) -> best B for every A cell.
    Returns (jA, iA, jB, iB, score).
    """
    b, _, s1, s2, s3, s4 = corr4d.shape
    JA, IA = np.meshgrid(range(s2), range(s1))
    JB, IB = np.meshgrid(range(s4), range(s3))
    JA, IA = torch.LongTensor(JA).view(1, -1), torch.LongTensor(IA).view(1, -1)
    JB, IB = torch.LongTensor(JB).view(1, -1), torch.LongTensor(IB).view(1, -1)
    if invert:
        v = corr4d.view(b, s1, s2, s3 * s4)
        if do_softmax:
            v = F.softmax(v, dim=3)
        vals, idx = torch.max(v, dim=3)
        score = vals.view(b, -1)
        iB = IB.view(-1)[idx.view(-1)].view(b, -1).contiguous()
        jB = JB.view(-1)[idx.view(-1)].view(b, -1).contiguous()
        iA = IA.expand_as(iB).contiguous()
        jA = JA.expand_as(jB).contiguous()
    else:
        v = corr4d.view(b, s1 * s2, s3, s4)
        if do_softmax:
            v = F.softmax(v, dim=1)
        vals, idx = torch.max(v, dim=1)
        score = vals.view(b, -1)
        iA = IA.view(-1)[idx.view(-1)].view(b, -1).contiguous()
        jA = JA.view(-1)[idx.view(-1)].view(b, -1).contiguous()
        iB = IB.expand_as(iA).contiguous()
        jB = JB.expand_as(jA).contiguous()
    if delta4d is not None:
        d_iA, d_jA, d_iB, d_jB = delta4d
        for n in range(b):
            sel = (iA[n], jA[n], iB[n], jB[n])
            diA, djA = d_iA[n][0][sel], d_jA[n][0][sel]
            diB, djB = d_iB[n][0][sel], d_jB[n][0][sel]
            iA[n] = iA[n] * ksize + diA
            jA[n] = jA[n] * ksize + djA
            iB[n] = iB[n] * ksize + diB
            jB[n] = jB[n] * ksize + djB
    return jA, iA, jB, iB, score


def cal_coarse_matches(corr4d, delta4d, ksize=1, do_softmax=True, upsample=16, center=True):
    """networks/patch2pix.py:340-375 (sort=False).  -> matches [b,Nc,4] int64, scores [b,Nc]."""
    xa, ya, xb, yb, sc = corr_to_matches(corr4d, delta4d, ksize, do_softmax, invert=False)
    xa2, ya2, xb2, yb2, sc2 = corr_to_matches(corr4d, delta4d, ksize, do_softmax, invert=True)
    cols = [torch.cat((p, q), 1).unsqueeze(1) for p, q in ((xa, xa2), (ya, ya2), (xb, xb2), (yb, yb2))]
    matches = upsample * torch.cat(cols, dim=1).permute(0, 2, 1)
    if center:
        matches = matches + upsample // 2
    return matches, torch.cat((sc, sc2), 1)


def filter_coarse(coarse_matches, match_scores, ncn_thres=0.0, mutual=True, ptmax=None):
    """networks/utils.py:38-72, including its quirks: lexicographic order from
    np.unique, first-occurrence scores, 'skip the filter if it would empty the set',
    and ptmax shuffle/tile drawn from the *global* numpy RNG."""
    out_m, out_s = [], []
    for im, sc in zip(coarse_matches, match_scores):
        _, ids, counts = np.unique(im.cpu().numpy(), axis=0, return_index=True, return_counts=True)
        if mutual:
            ids = ids[counts > 1]
        if len(ids) > 0:
            sc = sc[ids]
            im = im[ids]
        ids = torch.nonzero(sc.flatten() > ncn_thres, as_tuple=False).flatten()
        if ptmax:
            if len(ids) == 0:
                ids = torch.tensor([0, 0, 0, 0]).long()
            iids = np.arange(len(ids))
            np.random.shuffle(iids)
            iids = np.tile(iids, (ptmax // len(ids) + 1))[:ptmax]
            ids = ids[iids]
        if len(ids) > 0:
            sc = sc[ids]
            im = im[ids]
        out_m.append(im)
        out_s.append(sc)
    return out_m, out_s


def shift_to_anchors(matches, panc, pshift=PSHIFT):
    """networks/patch2pix.py:377-402 -- 8-row template (not 16, despite the comment)."""
    if panc == 1:
        return matches
    p = pshift
    tmpl = torch.tensor([[-p, -p, 0, 0], [p, -p, 0, 0], [-p, p, 0, 0], [p, p, 0, 0],
                         [0, 0, -p, -p], [0, 0, p, -p], [0, 0, -p, p], [0, 0, p, p]])
    return [(m.unsqueeze(1) + tmpl).reshape(-1, 4) for m in matches]


# --------------------------------------------------------------------------
# refine stage
# --------------------------------------------------------------------------
def select_local_patch_feats(feats1, feats2, ibatch, imatches, psize=PSIZE):
    """networks/utils.py:4-36 (ptype='center', feat_idx=[0,1,2,3]).

    (x,y)=imatches.long() truncates; level-j index = clamp((x+dx)//ds, 0, W//ds-1)
    with W,H taken from feats[0] (the image).  -> f1s,f2s [259, N*psize*psize].
    """
    dy, dx = torch.meshgrid(torch.arange(psize), torch.arange(psize), indexing='ij')
    dx = dx.flatten().view(1, -1) - psize // 2
    dy = dy.flatten().view(1, -1) - psize // 2
    _, _, h1, w1 = feats1[0].shape
    _, _, h2, w2 = feats2[0].shape
    x1, y1, x2, y2 = imatches.permute(1, 0).long()
    f1s, f2s = [], []
    for j, (fm1, fm2) in enumerate(zip(feats1, feats2)):
        if j not in FEAT_IDX:
            continue
        ds = int(np.prod(FEATS_DOWNSAMPLE[0:j + 1]))
        xs = lambda x, w: ((x.view(-1, 1) + dx).view(-1) // ds).long().clamp(min=0, max=w // ds - 1)
        ys = lambda y, h: ((y.view(-1, 1) + dy).view(-1) // ds).long().clamp(min=0, max=h // ds - 1)
        f1s.append(fm1[ibatch, :, ys(y1, h1), xs(x1, w1)])
        f2s.append(fm2[ibatch, :, ys(y2, h2), xs(x2, w2)])
    return torch.cat(f1s, dim=0), torch.cat(f2s, dim=0)


def feat_regress_net(f1, f2, sd, prefix):
    """networks/modules.py:56-112, feat_comb='pre', eval-mode BatchNorm.
    conv(518->512,k3,s2,p1) BN conv(512->512,k3,s1,p1) BN ReLU MaxPool(8) | FC 512-512-256-5."""
    g = lambda k: sd[prefix + k]
    x = torch.cat([f1, f2], dim=1)
    x = F.conv2d(x, g('conv.0.weight'), None, 2, 1)
    x = F.batch_norm(x, g('conv.1.running_mean'), g('conv.1.running_var'), g('conv.1.weight'), g('conv.1.bias'), False, 0.0, BN_EPS)
    x = F.conv2d(x, g('conv.2.weight'), None, 1, 1)
    x = F.batch_norm(x, g('conv.3.running_mean'), g('conv.3.running_var'), g('conv.3.weight'), g('conv.3.bias'), False, 0.0, BN_EPS)
    x = F.max_pool2d(F.relu(x), x.shape[-1])
    x = x.view(-1, x.shape[1])
    x = F.linear(x, g('fc.0.weight'), g('fc.0.bias'))
    x = F.relu(F.batch_norm(x, g('fc.1.running_mean'), g('fc.1.running_var'), g('fc.1.weight'), g('fc.1.bias'), False, 0.0, BN_EPS))
    x = F.linear(x, g('fc.3.weight'), g('fc.3.bias'))
    x = F.relu(F.batch_norm(x, g('fc.4.running_mean'), g('fc.4.running_var'), g('fc.4.weight'), g('fc.4.bias'), False, 0.0, BN_EPS))
    return F.linear(x, g('fc.6.weight'), g('fc.6.bias'))


def parse_regressor_out(out, psize, imatches, max_val):
    """networks/patch2pix.py:138-155 (ptype='center')."""
    w1, h1, w2, h2 = max_val
    off = psize * torch.tanh(F.relu(out[:, :4])) - psize // 2
    fm = imatches.float() + off
    probs = torch.sigmoid(out[:, 4])
    fm = torch.stack([fm[:, 0].clamp(min=0, max=w1), fm[:, 1].clamp(min=0, max=h1),
                      fm[:, 2].clamp(min=0, max=w2), fm[:, 3].clamp(min=0, max=h2)], dim=-1)
    return fm, probs


def forward_fine_match_mini_batch(feats1, feats2, ibatch, imatches, sd, prefix, psize=PSIZE):
    """networks/patch2pix.py:157-184."""
    n = imatches.shape[0]
    _, _, h1, w1 = feats1[0].shape
    _, _, h2, w2 = feats2[0].shape
    f1s, f2s = select_local_patch_feats(feats1, feats2, ibatch, imatches, psize)
    f1s = l2_normalize(f1s, 0).view(-1, n, psize, psize).permute(1, 0, 2, 3)
    f2s = l2_normalize(f2s, 0).view(-1, n, psize, psize).permute(1, 0, 2, 3)
    out = feat_regress_net(f1s, f2s, sd, prefix)
    return parse_regressor_out(out, psize, imatches, [w1, h1, w2, h2])


def forward_fine_match(feats1, feats2, coarse_matches, sd, prefix, psize=PSIZE, regr_batch=REGR_BATCH):
    """networks/patch2pix.py:186-218 -- chunks of regr_batch; a trailing chunk of
    one row is merged into the previous chunk; multi-chunk results are squeezed."""
    fine, probs = [], []
    for ib, im in enumerate(coarse_matches):
        n = im.shape[0]
        if n > regr_batch:
            edges = [regr_batch * i for i in range(n // regr_batch + 1)]
            if edges[-1] < n:
                if n - edges[-1] == 1:
                    edges[-1] = n
                else:
                    edges += [n]
            fm, pr = [], []
            for s, e in zip(edges[:-1], edges[1:]):
                r = forward_fine_match_mini_batch(feats1, feats2, ib, im[s:e], sd, prefix, psize)
                fm.append(r[0])
                pr.append(r[1])
            fm, pr = torch.cat(fm, 0).squeeze(), torch.cat(pr, 0).squeeze()
        else:
            fm, pr = forward_fine_match_mini_batch(feats1, feats2, ib, im, sd, prefix, psize)
        fine.append(fm)
        probs.append(pr)
    return fine, probs


# --------------------------------------------------------------------------
# top-level sequences
# --------------------------------------------------------------------------
def forward(im1, im2, sd, ksize=1, stages=None):
    """networks/patch2pix.py:220-237 (return_feats=True)."""
    feats1 = backbone_forward_all(im1, sd)
    feats2 = backbone_forward_all(im2, sd)
    corr4d, delta4d = forward_coarse_match(feats1[-1], feats2[-1], sd, ksize, stages)
    return corr4d, delta4d, feats1, feats2


def predict_coarse(im1, im2, sd, ksize=2, ncn_thres=0.0, mutual=False, center=True):
    """networks/patch2pix.py:240-248."""
    corr4d, delta4d, _, _ = forward(im1, im2, sd, ksize)
    cm, sc = cal_coarse_matches(corr4d, delta4d, ksize, upsample=UPSAMPLE, center=center)
    return filter_coarse(cm, sc, ncn_thres, mutual)


def hot_path_from_feats(feats1, feats2, sd, ksize=2, ncn_thres=0.0, mutual=True, ptmax=None, panc=1,
                        return_all=False):
    """Everything after the backbone.  panc=1, ptmax=None -> predict_fine
    (networks/patch2pix.py:250-276); panc=8, ptmax>0 -> the training-loop forward
    sequence (train_patch2pix.py:97-118), which is the benchmark configuration."""
    corr4d, delta4d = forward_coarse_match(feats1[-1], feats2[-1], sd, ksize)
    cm, sc = cal_coarse_matches(corr4d, delta4d, ksize, upsample=UPSAMPLE, center=True)
    if ptmax:
        if panc > 1 and ptmax > 0:
            cm, sc = filter_coarse(cm, sc, 0.0, True, ptmax=ptmax)
    else:
        cm, sc = filter_coarse(cm, sc, ncn_thres, mutual)
    cm = shift_to_anchors(cm, panc)
    mid, mid_p = forward_fine_match(feats1, feats2, cm, sd, 'regress_mid.')
    fine, fine_p = forward_fine_match(feats1, feats2, mid, sd, 'regress_fine.')
    if return_all:
        return fine, fine_p, mid, mid_p, cm
    return fine, fine_p, cm


def predict_fine(im1, im2, sd, ksize=2, ncn_thres=0.0, mutual=True, return_all=False):
    """networks/patch2pix.py:250-276."""
    feats1 = backbone_forward_all(im1, sd)
    feats2 = backbone_forward_all(im2, sd)
    return hot_path_from_feats(feats1, feats2, sd, ksize, ncn_thres, mutual, None, 1, return_all)


def train_forward_sequence(im1, im2, sd, ksize=2, ptmax=400, panc=8, return_all=False):
    """train_patch2pix.py:97-118 under eval()/no_grad (benchmark 'ptmax=400 panc=8')."""
    feats1 = backbone_forward_all(im1, sd)
    feats2 = backbone_forward_all(im2, sd)
    return hot_path_from_feats(feats1, feats2, sd, ksize, 0.0, True, ptmax, panc, return_all)


def refine_matches(im1, im2, coarse_matches, sd, io_thres):
    """networks/patch2pix.py:278-318."""
    if len(coarse_matches) == 0:
        return np.empty((0, 4)), np.empty((0,)), np.empty((0, 4))
    if isinstance(coarse_matches, np.ndarray):
        cm_ = torch.from_numpy(coarse_matches).unsqueeze(0)
    else:
        cm_ = coarse_matches.unsqueeze(0)
        coarse_matches = coarse_matches.cpu().numpy()
    feats1 = backbone_forward_all(im1, sd)
    feats2 = backbone_forward_all(im2, sd)
    mid, _ = forward_fine_match(feats1, feats2, cm_, sd, 'regress_mid.')
    fine, fine_p = forward_fine_match(feats1, feats2, mid, sd, 'regress_fine.')
    refined = fine[0].cpu().numpy()
    scores = fine_p[0].cpu().numpy()
    if io_thres > 0:
        pos = np.where(scores > io_thres)[0]
        if len(pos) > 0:
            coarse_matches, refined, scores = coarse_matches[pos], refined[pos], scores[pos]
    return refined, scores, coarse_matches
