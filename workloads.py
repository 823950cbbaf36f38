"""Small synthetic models used by the golden fixtures, tests, smoke and bench.

TinyModel has the same architecture as the reference's test fixture
(testing/models.py:13-31: Linear(10,20,bias=False) -> ReLU -> Linear(20,10) ->
Softmax).  SmallConvNet covers every layer geometry on the hot path: k3/s1/p1
with bias, k3/s2/p1 without bias, 1x1, 1x1 stride 2, and a biased Linear.
ResNet-32/ResNet-50 are own re-statements of the standard architectures
(examples/vision/cifar_resnet.py:155-219; torchvision resnet50) used for the
BASELINE.json workloads with random-init weights and synthetic data.
"""
from __future__ import annotations

import torch
from torch import nn
import torch.nn.functional as F


class TinyModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.linear1 = nn.Linear(10, 20, bias=False)
        self.activation = nn.ReLU()
        self.linear2 = nn.Linear(20, 10)
        self.softmax = nn.Softmax(dim=1)

    def forward(self, x):
        return self.softmax(self.linear2(self.activation(self.linear1(x))))


class SmallConvNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 6, 3, stride=1, padding=1, bias=True)
        self.conv2 = nn.Conv2d(6, 8, 3, stride=2, padding=1, bias=False)
        self.conv3 = nn.Conv2d(8, 8, 1, bias=False)
        self.down = nn.Conv2d(8, 10, 1, stride=2, bias=True)
        self.fc = nn.Linear(10, 5)

    def forward(self, x):
        x = F.relu(self.conv1(x))
        x = F.relu(self.conv2(x))
        x = F.relu(self.conv3(x))
        x = F.relu(self.down(x))
        x = F.adaptive_avg_pool2d(x, 1).flatten(1)
        return self.fc(x)


class GeomConvNet(nn.Module):
    """Asymmetric kernels / strides / paddings (modules.py:210-237 handles each axis on its
    own): (3,2) stride (2,1) pad (1,0) with bias; 5x5 stride 2 pad 2; (1,3) pad (0,1); Linear
    without bias."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(2, 4, (3, 2), stride=(2, 1), padding=(1, 0), bias=True)
        self.conv2 = nn.Conv2d(4, 5, 5, stride=2, padding=2, bias=False)
        self.conv3 = nn.Conv2d(5, 6, (1, 3), stride=1, padding=(0, 1), bias=True)
        self.fc = nn.Linear(6, 4, bias=False)

    def forward(self, x):
        x = F.relu(self.conv1(x))
        x = F.relu(self.conv2(x))
        x = F.relu(self.conv3(x))
        return self.fc(F.adaptive_avg_pool2d(x, 1).flatten(1))


class SeqModel(nn.Module):
    """Linear layers on (batch, seq, features) inputs: all leading dims are flattened into
    rows (modules.py:129,140)."""

    def __init__(self):
        super().__init__()
        self.proj = nn.Linear(12, 16)
        self.out = nn.Linear(16, 7, bias=False)

    def forward(self, x):
        return self.out(torch.tanh(self.proj(x)))


MODEL_ZOO = {'TinyModel': TinyModel, 'SmallConvNet': SmallConvNet, 'GeomConvNet': GeomConvNet,
             'SeqModel': SeqModel}


# ------------------------------------------------------------------ ResNets
class _BasicBlock(nn.Module):
    def __init__(self, inp, out, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(inp, out, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(out)
        self.conv2 = nn.Conv2d(out, out, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(out)
        self.stride, self.inp, self.out = stride, inp, out

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        if self.stride != 1 or self.inp != self.out:
            # CIFAR "option A" parameter-free shortcut (cifar_resnet.py:82-100)
            sc = x[:, :, ::2, ::2]
            pad = self.out // 4
            sc = F.pad(sc, (0, 0, 0, 0, pad, pad))
        else:
            sc = x
        return F.relu(y + sc)


class CifarResNet(nn.Module):
    def __init__(self, n_blocks=5, num_classes=10):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 16, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(16)
        layers, inp = [], 16
        for out, stride in ((16, 1), (32, 2), (64, 2)):
            blocks = []
            for b in range(n_blocks):
                blocks.append(_BasicBlock(inp, out, stride if b == 0 else 1))
                inp = out
            layers.append(nn.Sequential(*blocks))
        self.layer1, self.layer2, self.layer3 = layers
        self.linear = nn.Linear(64, num_classes)

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        x = self.layer3(self.layer2(self.layer1(x)))
        x = F.adaptive_avg_pool2d(x, 1).flatten(1)
        return self.linear(x)


def resnet32():
    return CifarResNet(5)


class _Bottleneck(nn.Module):
    def __init__(self, inp, mid, stride):
        super().__init__()
        out = mid * 4
        self.conv1 = nn.Conv2d(inp, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid)
        self.conv2 = nn.Conv2d(mid, mid, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(mid)
        self.conv3 = nn.Conv2d(mid, out, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(out)
        self.downsample = None
        if stride != 1 or inp != out:
            self.downsample = nn.Sequential(
                nn.Conv2d(inp, out, 1, stride, bias=False), nn.BatchNorm2d(out))

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        sc = x if self.downsample is None else self.downsample(x)
        return F.relu(y + sc)


class ResNet50(nn.Module):
    def __init__(self, num_classes=1000, width=64, blocks=(3, 4, 6, 3)):
        super().__init__()
        self.conv1 = nn.Conv2d(3, width, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        inp, stages = width, []
        for i, n in enumerate(blocks):
            mid = width * 2 ** i
            seq = []
            for b in range(n):
                seq.append(_Bottleneck(inp, mid, (2 if i > 0 else 1) if b == 0 else 1))
                inp = mid * 4
            stages.append(nn.Sequential(*seq))
        self.layer1, self.layer2, self.layer3, self.layer4 = stages
        self.fc = nn.Linear(inp, num_classes)

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, 2, 1)
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(F.adaptive_avg_pool2d(x, 1).flatten(1))


def resnet50():
    return ResNet50()


# ------------------------------------------------------------------ GPT-NeoX-125M-width Linear stack
class ColumnParallelLinear(nn.Linear):
    """Named like the Megatron layer the reference's GPT-NeoX front end registers
    (kfac/gpt_neox/preconditioner.py:481-502; testing/gpt_neox.py:16-25); TP = 1."""


class RowParallelLinear(nn.Linear):
    """See ColumnParallelLinear."""


class NeoXBlock(nn.Module):
    """The four K-FAC-registered Linear layers of one GPT-NeoX transformer block at TP = 1
    (hidden 768: qkv 768->2304, dense 768->768, h_to_4h 768->3072, 4h_to_h 3072->768).  The
    attention mixing itself holds no K-FAC layer and is replaced by a head-wise product."""

    def __init__(self, hidden=768):
        super().__init__()
        self.query_key_value = ColumnParallelLinear(hidden, 3 * hidden)
        self.dense = RowParallelLinear(hidden, hidden)
        self.dense_h_to_4h = ColumnParallelLinear(hidden, 4 * hidden)
        self.dense_4h_to_h = RowParallelLinear(4 * hidden, hidden)

    def forward(self, x):
        q, k, v = self.query_key_value(x).chunk(3, dim=-1)
        x = x + self.dense(torch.tanh(q * k) * v)
        return x + self.dense_4h_to_h(F.gelu(self.dense_h_to_4h(x)))


class NeoXStack(nn.Module):
    def __init__(self, layers=1, hidden=768):
        super().__init__()
        self.blocks = nn.Sequential(*[NeoXBlock(hidden) for _ in range(layers)])

    def forward(self, x):
        return self.blocks(x)
