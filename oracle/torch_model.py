"""ORACLE (test infrastructure — only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may import this; the product path never does).

PyTorch restatement of the reference's own oracle, /root/reference/pytorch_inference.py:
  ResnetBlock  :29-82   bottleneck: conv1x1 -> bn -> relu -> conv3x3(stride) -> bn -> relu -> conv1x1
                        -> bn -> += shortcut -> relu
  make_layer   :85-110  first block carries the stride and (when shape changes) a conv1x1+bn shortcut
  Resnet152    :113-162 conv7x7/2 -> bn -> relu -> maxpool3/2 -> 4 layers -> adaptive avgpool -> fc
generalised to the depths BASELINE.json's configs name (18/34 use the torchvision BasicBlock, which
the reference does not have — SURVEY.md section 0). Parameter names match torchvision's state_dict, which
is what save_weights.py dumps, so `load_state_dict(strict=True)` works for both.

PIN: tests/test_oracle_pin.py checks (in the build container, where /root/reference exists) that
this module is bit-identical to the reference's Resnet152 class executed from its own source file,
and everywhere that it is bit-identical to torchvision.models.resnet*; golden logits generated from
it are committed under tests/golden/.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

ARCH_SPECS = {
    "resnet18": (False, (2, 2, 2, 2)),
    "resnet34": (False, (3, 4, 6, 3)),
    "resnet50": (True, (3, 4, 6, 3)),
    "resnet101": (True, (3, 4, 23, 3)),
    "resnet152": (True, (3, 8, 36, 3)),  # pytorch_inference.py:127-130
}


def _shortcut(in_ch, out_ch, stride):
    # pytorch_inference.py:86-101
    if stride != 1 or in_ch != out_ch:
        return nn.Sequential(nn.Conv2d(in_ch, out_ch, kernel_size=1, stride=stride, bias=False),
                             nn.BatchNorm2d(out_ch))
    return nn.Identity()


class Bottleneck(nn.Module):
    """pytorch_inference.py:29-82 (ResnetBlock)."""

    def __init__(self, in_ch, mid_ch, out_ch, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(in_ch, mid_ch, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid_ch)
        self.conv2 = nn.Conv2d(mid_ch, mid_ch, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(mid_ch)
        self.conv3 = nn.Conv2d(mid_ch, out_ch, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(out_ch)
        self.downsample = downsample if downsample is not None else nn.Identity()

    def forward(self, x):
        shortcut = self.downsample(x)
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        y = y + shortcut
        return F.relu(y)


class BasicBlock(nn.Module):
    """torchvision BasicBlock semantics (absent from the reference)."""

    def __init__(self, in_ch, out_ch, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(out_ch)
        self.conv2 = nn.Conv2d(out_ch, out_ch, kernel_size=3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(out_ch)
        self.downsample = downsample if downsample is not None else nn.Identity()

    def forward(self, x):
        shortcut = self.downsample(x)
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        y = y + shortcut
        return F.relu(y)


def _make_layer(bottleneck, in_ch, mid_ch, n_blocks, stride):
    out_ch = mid_ch * 4 if bottleneck else mid_ch
    blocks = []
    for i in range(n_blocks):
        s = stride if i == 0 else 1
        ic = in_ch if i == 0 else out_ch
        ds = _shortcut(ic, out_ch, s) if i == 0 else None
        if isinstance(ds, nn.Identity):
            ds = None
        blocks.append(Bottleneck(ic, mid_ch, out_ch, s, ds) if bottleneck else BasicBlock(ic, out_ch, s, ds))
    return nn.Sequential(*blocks), out_ch


class OracleResNet(nn.Module):
    """pytorch_inference.py:113-162 (Resnet152) for any depth in ARCH_SPECS."""

    def __init__(self, arch: str, n_classes: int = 1000):
        super().__init__()
        bottleneck, counts = ARCH_SPECS[arch]
        self.arch = arch
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        ch = 64
        self.layer1, ch = _make_layer(bottleneck, ch, 64, counts[0], 1)
        self.layer2, ch = _make_layer(bottleneck, ch, 128, counts[1], 2)
        self.layer3, ch = _make_layer(bottleneck, ch, 256, counts[2], 2)
        self.layer4, ch = _make_layer(bottleneck, ch, 512, counts[3], 2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(ch, n_classes)

    def forward(self, x, taps=None):
        """`taps`, if a dict, receives the intermediate activations by the engine's names."""
        y = F.relu(self.bn1(self.conv1(x)))
        if taps is not None:
            taps["stem"] = y
        y = self.maxpool(y)
        if taps is not None:
            taps["maxpool"] = y
        for li, layer in enumerate((self.layer1, self.layer2, self.layer3, self.layer4), start=1):
            for bi, block in enumerate(layer):
                y = block(y)
                if taps is not None:
                    taps[f"layer{li}.{bi}"] = y
        y = self.avgpool(y)
        if taps is not None:
            taps["avgpool"] = y.flatten(1)
        return self.fc(y.flatten(-3))


def build(arch: str, state_dict, dtype=torch.float32) -> OracleResNet:
    m = OracleResNet(arch, n_classes=state_dict["fc.bias"].numel())
    m.load_state_dict(state_dict, strict=True)
    return m.to(dtype).eval()


@torch.no_grad()
def run(arch: str, state_dict, x: torch.Tensor, dtype=torch.float32, taps=None) -> torch.Tensor:
    """Logits of the oracle on CPU in `dtype` (fp32 = the reference path, fp64 = tie arbiter)."""
    m = build(arch, state_dict, dtype)
    return m(x.to(dtype), taps)
