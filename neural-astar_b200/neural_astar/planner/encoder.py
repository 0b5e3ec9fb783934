"""Cost-map encoders (plain PyTorch; the north star keeps them on cuDNN).

Same classes, constructor arguments and state-dict layout as
/root/reference/src/neural_astar/planner/encoder.py:12-97 so reference checkpoints load:
`model.{0,3,6,...}` are 3x3 convs, `model.{1,4,7,...}` BatchNorms, ReLU (and MaxPool for
CNNDownSize) in between, trailing activation(s) dropped; output = sigmoid(model(x)) * const.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn


class EncoderBase(nn.Module):
    def __init__(self, input_dim: int, encoder_depth: int = 4, const: Optional[float] = None):
        super().__init__()
        self.model = self.construct_encoder(input_dim, encoder_depth)
        # learnable scale when given (reference :25-28); plain 1.0 otherwise
        self.const = nn.Parameter(torch.ones(1) * const) if const is not None else 1.0

    def construct_encoder(self, input_dim: int, encoder_depth: int) -> nn.Module:
        raise NotImplementedError

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.sigmoid(self.model(x)) * self.const


def _conv_stack(widths: List[int], pool: bool) -> nn.Sequential:
    layers: List[nn.Module] = []
    tail = 2 if pool else 1  # modules to drop after the last BatchNorm
    for cin, cout in zip(widths[:-1], widths[1:]):
        layers += [nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(cout), nn.ReLU()]
        if pool:
            layers.append(nn.MaxPool2d((2, 2)))
    return nn.Sequential(*layers[:-tail])


class CNN(EncoderBase):
    CHANNELS = [32, 64, 128, 256]

    def construct_encoder(self, input_dim: int, encoder_depth: int) -> nn.Module:
        return _conv_stack([input_dim] + self.CHANNELS[:encoder_depth] + [1], pool=False)


class CNNDownSize(CNN):
    def construct_encoder(self, input_dim: int, encoder_depth: int) -> nn.Module:
        return _conv_stack([input_dim] + self.CHANNELS[:encoder_depth] + [1], pool=True)


class Unet(EncoderBase):
    """U-Net encoder; needs the optional third-party `segmentation_models_pytorch` (reference :37-57)."""

    DECODER_CHANNELS = [256, 128, 64, 32, 16]

    def construct_encoder(self, input_dim: int, encoder_depth: int) -> nn.Module:
        try:
            import segmentation_models_pytorch as smp
        except ImportError as e:  # not shipped in this image; used by no reference config
            raise ImportError("encoder_arch='Unet' needs segmentation_models_pytorch==0.3.1") from e
        return smp.Unet(
            encoder_name="vgg16_bn",
            encoder_weights=None,
            classes=1,
            in_channels=input_dim,
            encoder_depth=encoder_depth,
            decoder_channels=self.DECODER_CHANNELS[:encoder_depth],
        )
