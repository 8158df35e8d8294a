"""The reference's gain-regressor CNNs, kept in PyTorch (north_star: "The CNN gain regressor itself
stays in PyTorch"), behind the same forward contract:

    forward(x: float32 [B, 4, 1025, T]) -> (masked [B, 1025, T], (m1, m2, m3, m4) each [B, 1])

Module attribute names, construction order and hyper-parameters follow the reference so that (a) a
reference ``state_dict`` loads unchanged and (b) the same ``torch.manual_seed`` gives the same initial
weights (tests compare seeded outputs with goldens produced by the reference).  Only the tail --
``masked = sum_i m_i * x[:, i]`` (model_scalar_1s.py:260-275, model_scalar_2s.py:117-132,
model_resnet.py:120-126) -- is on the hot path; it runs as one fused CUDA kernel with autograd
(mixing.feature_mix) instead of the reference's eight elementwise launches.

References: models/model_scalar_1s.py:151-275, models/model_scalar_2s.py:9-132,
models/model_resnet.py:6-130.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import mixing


class ConvBlock2d(nn.Module):
    """Conv2D -> BatchNorm(momentum 0.90, eps 1e-3) -> ReLU -> Dropout (training only), no padding
    (model_scalar_1s.py:151-190)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dilation=1, dropout_p=-1.0):
        super().__init__()
        self.add_padding = None
        self.conv = nn.Conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size,
                              stride=stride, padding=0, dilation=dilation)
        self.batch_norm = nn.BatchNorm2d(num_features=out_channels, momentum=0.90, eps=0.001)
        self.activation = nn.ReLU()
        self.dropout = nn.Dropout(dropout_p) if dropout_p != -1 else None

    def forward(self, x):
        out = self.activation(self.batch_norm(self.conv(x)))
        if self.training and self.dropout:
            out = self.dropout(out)
        return out


def dB_to_amplitude(x: torch.Tensor):
    """10**(0.5*x), sic (model_scalar_1s.py:193-197)."""
    return torch.pow(10.0, 0.5 * x)


def amplitude_to_dB(x: torch.Tensor):
    return 20 * torch.log10(x)


class _FourHeadTail(nn.Module):
    """Four (1x1 conv -> ReLU -> FC) heads and the gain-weighted stem sum shared by all models."""

    def _make_heads(self, channels, flattened_dim):
        for i in range(1, 5):
            setattr(self, 'conv_head%d' % i, nn.Conv2d(channels, 1, kernel_size=(1, 1)))
            setattr(self, 'fc_head%d' % i, nn.Linear(flattened_dim, 1))

    def _heads_and_tail(self, x, out):
        ms = []
        for i in range(1, 5):
            m = F.relu(getattr(self, 'conv_head%d' % i)(out))
            ms.append(getattr(self, 'fc_head%d' % i)(m.view((x.size()[0], -1))))
        masked = mixing.feature_mix(x, torch.cat(ms, dim=1))
        return masked, tuple(ms)


class MixingModelScalar1s(_FourHeadTail):
    """models/model_scalar_1s.py:207-275; accepts T in {87, 88} (flattened_dim 10290)."""

    def __init__(self):
        super().__init__()
        self.conv_b1 = ConvBlock2d(4, 16, 3, dropout_p=0.2, stride=2)
        self.conv_b2 = ConvBlock2d(16, 32, 5, dropout_p=0.2)
        self.conv_b3 = ConvBlock2d(32, 48, 5, dropout_p=0.2)
        self.conv_b4 = ConvBlock2d(48, 64, 7, dropout_p=0.2)
        self.conv_b5 = ConvBlock2d(64, 128, 9, dropout_p=0.3)
        self._make_heads(128, 10290)

    def forward(self, x):
        out = self.conv_b5(self.conv_b4(self.conv_b3(self.conv_b2(self.conv_b1(x)))))
        return self._heads_and_tail(x, out)


class MixingModelScalar2s(_FourHeadTail):
    """models/model_scalar_2s.py:64-132; first conv dilation 2, T in {173, 174} (30807)."""

    def __init__(self):
        super().__init__()
        self.conv_b1 = ConvBlock2d(4, 16, 3, dropout_p=0.2, stride=2, dilation=2)
        self.conv_b2 = ConvBlock2d(16, 32, 5, dropout_p=0.2)
        self.conv_b3 = ConvBlock2d(32, 48, 5, dropout_p=0.2)
        self.conv_b4 = ConvBlock2d(48, 64, 7, dropout_p=0.2)
        self.conv_b5 = ConvBlock2d(64, 128, 9, dropout_p=0.3)
        self._make_heads(128, 30807)

    def forward(self, x):
        out = self.conv_b5(self.conv_b4(self.conv_b3(self.conv_b2(self.conv_b1(x)))))
        return self._heads_and_tail(x, out)


class BasicBlock(nn.Module):
    """models/model_resnet.py:6-28."""
    expansion = 1

    def __init__(self, in_channels, out_channels, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_channels != self.expansion * out_channels:
            self.shortcut = nn.Sequential(
                nn.Conv2d(in_channels, self.expansion * out_channels, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(self.expansion * out_channels))

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        out = out + self.shortcut(x)
        return F.relu(out)


class ResNet(_FourHeadTail):
    """models/model_resnet.py:59-126: six stages 16..256 channels, flattened_dim 231 (T = 216)."""

    def __init__(self, block, num_blocks):
        super().__init__()
        self.in_planes = 16
        self.conv1 = nn.Conv2d(4, 16, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(16)
        self.layer1 = self._make_layer(block, 16, num_blocks[0], stride=1)
        self.layer2 = self._make_layer(block, 32, num_blocks[1], stride=2)
        self.layer3 = self._make_layer(block, 64, num_blocks[2], stride=2)
        self.layer4 = self._make_layer(block, 96, num_blocks[3], stride=2)
        self.layer5 = self._make_layer(block, 128, num_blocks[4], stride=2)
        self.layer6 = self._make_layer(block, 256, num_blocks[5], stride=2)
        self._make_heads(256, 231)

    def _make_layer(self, block, planes, num_blocks, stride):
        layers = []
        for st in [stride] + [1] * (num_blocks - 1):
            layers.append(block(self.in_planes, planes, st))
            self.in_planes = planes * block.expansion
        return nn.Sequential(*layers)

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        for i in range(1, 7):
            out = getattr(self, 'layer%d' % i)(out)
        return self._heads_and_tail(x, out)


def ResNet18():
    return ResNet(BasicBlock, [2, 2, 2, 2, 2, 2])


class DummyModel(nn.Module):
    """models/baselines/dummy_model.py:18-34: dB -> amplitude -> sum -> dB (kernel K6)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return mixing.db_sum(x)


class RandomModel:
    """models/baselines/random_model.py: one U(gain_from, gain_to) gain per stem (python float x
    float32 array -> float32)."""

    def __init__(self, gain_from=0.5, gain_to=1.5):
        self.tracklist = ('bass', 'drums', 'vocals', 'other')
        self._gain_from = gain_from
        self._gain_to = gain_to

    def draw(self):
        import numpy as np
        return [float(np.random.uniform(self._gain_from, self._gain_to)) for _ in self.tracklist]

    def forward(self, x: dict) -> dict:
        from .inference_utils import stack_tracks, unstack_tracks
        stems, host = stack_tracks(x, self.tracklist)
        gains = [[g] for g in self.draw()]
        out = mixing.gain_mix(stems, gains, hold_single=True, compute_dtype=stems.dtype,
                              want_mix=False, want_stems=True)['stems']
        return unstack_tracks(out, self.tracklist, host)


class MeanLoudnessModel:
    """models/baselines/mean_loudness_model.py: each stem is loudness-normalised to the train-set
    mean loudness of its track."""

    def __init__(self, d_mean_loudness: dict, sr=44100):
        from .loudness import Meter
        self.mean_loudness = d_mean_loudness
        self.meter = Meter(sr)
        self.tracklist = ('bass', 'drums', 'vocals', 'other')

    def forward(self, x: dict) -> dict:
        import numpy as np
        from .inference_utils import stack_tracks, unstack_tracks
        stems, host = stack_tracks(x, self.tracklist)
        loud = self.meter.integrated_loudness_batch(stems)
        gains = [[float(np.power(10.0, (self.mean_loudness[t] - l) / 20.0))] for t, l in zip(self.tracklist, loud)]
        # pyln.normalize.loudness: float64 gain x data, in float64 (numpy >= 2 promotion)
        out = mixing.gain_mix(stems, gains, hold_single=True, compute_dtype=torch.float64,
                              want_mix=False, want_stems=True)['stems']
        return unstack_tracks(out, self.tracklist, host)
