"""ResNet34 feature pyramid that FEEDS the hot path (not part of it).

Plain PyTorch/cuDNN.  Same truncation as the reference's extractor (networks/resnet.py:96-157,
169-179): conv1/bn1/relu -> maxpool -> layer1 (3 blocks, 64) -> layer2 (4, 128, stride 2) ->
layer3 (6, 256, stride forced to 1 when change_stride) ; layer4 is never run and is not built.
Parameter names match the reference checkpoint (`extract.conv1.weight`, `extract.layer3.0.
downsample.1.running_var`, ...).
"""
import torch.nn as nn


class _Block(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = self.bn2(self.conv2(self.relu(self.bn1(self.conv1(x)))))
        if self.downsample is not None:
            x = self.downsample(x)
        return self.relu(y + x)


def _stage(cin, cout, n, stride):
    return nn.Sequential(*[_Block(cin if i == 0 else cout, cout, stride if i == 0 else 1) for i in range(n)])


class ResNet34Features(nn.Module):
    def __init__(self, change_stride=True):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = _stage(64, 64, 3, 1)
        self.layer2 = _stage(64, 128, 4, 2)
        self.layer3 = _stage(128, 256, 6, 2)
        if change_stride:
            self.change_stride('layer3')

    def change_stride(self, target='layer3'):
        blk = getattr(self, target)[0]
        blk.conv1.stride = (1, 1)
        blk.downsample[0].stride = (1, 1)

    def forward_all(self, x, feat_list=None, early_feat=True):
        """Appends [image, conv1-relu, layer1, layer2, layer3] to feat_list (and returns it)."""
        if not early_feat:
            raise RuntimeError('layer4 is never used by Patch2Pix and is not built (early_feat must be True)')
        feat_list = [] if feat_list is None else feat_list
        feat_list.append(x)
        x = self.relu(self.bn1(self.conv1(x)))
        feat_list.append(x)
        x = self.layer1(self.maxpool(x))
        feat_list.append(x)
        x = self.layer2(x)
        feat_list.append(x)
        x = self.layer3(x)
        feat_list.append(x)
        return feat_list

    def forward(self, x, early_feat=True):
        return self.forward_all(x, [], early_feat)[-1]
