"""TEST INFRASTRUCTURE -- never imported by the product package.

Restatement of the *caller* of the correlation path, so that the "final flow within
0.01 px mean EPE after 12 GRU iterations" criterion can be tested where the reference
checkout does not exist (the GPU box): RAFT's full-size model as flow-supervisor runs it,

    /root/reference/pytorch/core/raft.py:24-144        (RAFT.__init__, forward, upsample_flow)
    /root/reference/pytorch/core/extractor.py:6-57     (ResidualBlock)
    /root/reference/pytorch/core/extractor.py:118-192  (BasicEncoder)
    /root/reference/pytorch/core/update.py:6-16,33-60,79-96,113-139
                                                       (FlowHead, SepConvGRU, BasicMotionEncoder, BasicUpdateBlock)

Only the non-small variant (D = 256, 4 levels, radius 4) without dropout / mixed precision
is restated.  Sub-module names, construction order and initialisation follow the reference,
so that (a) a reference ``state_dict`` loads unchanged and (b) ``torch.manual_seed(s)`` followed
by ``Raft()`` yields the same weights as the reference under the same seed -- both are pinned
in tests/test_oracle_raft.py against the live reference and against a committed golden flow
(tests/golden/raft_seed1234_128x160.npz, made by oracle/make_golden_raft.py).

The correlation block is injected: ``Raft.forward(..., corr_block=cls)`` calls
``cls(fmap1, fmap2, num_levels=4, radius=4)`` and ``corr_fn(coords1)`` exactly where raft.py:104-107
and raft.py:124 do.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _norm(kind: str, ch: int) -> nn.Module:
    if kind == "instance":
        return nn.InstanceNorm2d(ch)
    if kind == "batch":
        return nn.BatchNorm2d(ch)
    raise ValueError(kind)


class _Res(nn.Module):
    """extractor.py:6-57: two 3x3 convs, each followed by norm + relu; a strided 1x1
    projection (+ norm) on the skip path when the block down-samples."""

    def __init__(self, cin, cout, kind, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride=stride, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1, self.norm2 = _norm(kind, cout), _norm(kind, cout)
        self.downsample = None
        if stride != 1:
            self.norm3 = _norm(kind, cout)
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride), self.norm3)

    def forward(self, x):
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.relu(self.norm2(self.conv2(y)))
        skip = x if self.downsample is None else self.downsample(x)
        return self.relu(skip + y)


class _Encoder(nn.Module):
    """extractor.py:118-192 (BasicEncoder): 7x7/2 stem, three stages of two residual
    blocks (64, 96/2, 128/2), 1x1 projection; 1/8 resolution."""

    def __init__(self, out_ch, kind):
        super().__init__()
        self.norm1 = _norm(kind, 64)
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3)
        self.relu1 = nn.ReLU(inplace=True)
        stages, cin = [], 64
        for cout, stride in ((64, 1), (96, 2), (128, 2)):
            stages.append(nn.Sequential(_Res(cin, cout, kind, stride), _Res(cout, cout, kind, 1)))
            cin = cout
        self.layer1, self.layer2, self.layer3 = stages
        self.conv2 = nn.Conv2d(128, out_ch, 1)
        for m in self.modules():                       # extractor.py:149-156
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d)):
                if m.weight is not None:
                    nn.init.constant_(m.weight, 1)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, x):
        pair = isinstance(x, (tuple, list))
        if pair:
            n = x[0].shape[0]
            x = torch.cat(list(x), 0)
        x = self.relu1(self.norm1(self.conv1(x)))
        x = self.conv2(self.layer3(self.layer2(self.layer1(x))))
        return torch.split(x, [n, n], 0) if pair else x


class _MotionEncoder(nn.Module):
    """update.py:79-96: 324 correlation channels + 2 flow channels -> 126 + 2 motion features."""

    def __init__(self, corr_ch):
        super().__init__()
        self.convc1 = nn.Conv2d(corr_ch, 256, 1)
        self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(64 + 192, 128 - 2, 3, padding=1)

    def forward(self, flow, corr):
        c = F.relu(self.convc2(F.relu(self.convc1(corr))))
        f = F.relu(self.convf2(F.relu(self.convf1(flow))))
        return torch.cat([F.relu(self.conv(torch.cat([c, f], 1))), flow], 1)


class _SepGru(nn.Module):
    """update.py:33-60: GRU with a 1x5 pass followed by a 5x1 pass."""

    def __init__(self, hidden, inp):
        super().__init__()
        for tag, k, p in (("1", (1, 5), (0, 2)), ("2", (5, 1), (2, 0))):
            for gate in "zrq":
                setattr(self, f"conv{gate}{tag}", nn.Conv2d(hidden + inp, hidden, k, padding=p))

    def forward(self, h, x):
        for tag in "12":
            hx = torch.cat([h, x], 1)
            z = torch.sigmoid(getattr(self, "convz" + tag)(hx))
            r = torch.sigmoid(getattr(self, "convr" + tag)(hx))
            q = torch.tanh(getattr(self, "convq" + tag)(torch.cat([r * h, x], 1)))
            h = (1 - z) * h + z * q
        return h


class _FlowHead(nn.Module):
    def __init__(self, cin, hidden):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, hidden, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden, 2, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.conv2(self.relu(self.conv1(x)))


class _UpdateBlock(nn.Module):
    """update.py:113-139 (BasicUpdateBlock)."""

    def __init__(self, corr_ch, hidden=128):
        super().__init__()
        self.encoder = _MotionEncoder(corr_ch)
        self.gru = _SepGru(hidden, 128 + hidden)
        self.flow_head = _FlowHead(hidden, 256)
        self.mask = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(256, 64 * 9, 1))

    def forward(self, net, inp, corr, flow):
        net = self.gru(net, torch.cat([inp, self.encoder(flow, corr)], 1))
        return net, 0.25 * self.mask(net), self.flow_head(net)


def grid_xy(n, h, w, device):
    """utils.py:74-77: channel 0 = x, channel 1 = y."""
    ys, xs = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
    return torch.stack([xs, ys], 0).float()[None].repeat(n, 1, 1, 1)


def convex_upsample(flow, mask):
    """raft.py:72-83: every fine pixel is a softmax-weighted mix of its 3x3 coarse neighbourhood."""
    n, _, h, w = flow.shape
    wgt = torch.softmax(mask.view(n, 1, 9, 8, 8, h, w), dim=2)
    nb = F.unfold(8 * flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w)
    up = torch.sum(wgt * nb, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(n, 2, 8 * h, 8 * w)


class Raft(nn.Module):
    """raft.py:24-144, non-small configuration, test_mode=True semantics of ``forward``."""

    LEVELS, RADIUS, HIDDEN, CONTEXT = 4, 4, 128, 128

    def __init__(self):
        super().__init__()
        self.fnet = _Encoder(256, "instance")
        self.cnet = _Encoder(self.HIDDEN + self.CONTEXT, "batch")
        self.update_block = _UpdateBlock(self.LEVELS * (2 * self.RADIUS + 1) ** 2, self.HIDDEN)

    def forward(self, image1, image2, iters=12, corr_block=None, return_all=False):
        image1 = (2 * (image1 / 255.0) - 1.0).contiguous()
        image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
        fmap1, fmap2 = self.fnet([image1, image2])
        corr_fn = corr_block(fmap1.float(), fmap2.float(), num_levels=self.LEVELS, radius=self.RADIUS)
        net, inp = torch.split(self.cnet(image1), [self.HIDDEN, self.CONTEXT], 1)
        net, inp = torch.tanh(net), torch.relu(inp)
        n, _, h, w = image1.shape
        coords0 = grid_xy(n, h // 8, w // 8, image1.device)
        coords1 = coords0.clone()
        ups = []
        for _ in range(iters):
            coords1 = coords1.detach()
            corr = corr_fn(coords1)
            net, mask, delta = self.update_block(net, inp, corr, coords1 - coords0)
            coords1 = coords1 + delta
            ups.append(convex_upsample(coords1 - coords0, mask))
        return ups if return_all else (coords1 - coords0, ups[-1])
