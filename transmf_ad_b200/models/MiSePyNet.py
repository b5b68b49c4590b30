"""MiSePyNet / Mnet baseline on the B200 kernels -- drop-in for the reference ``models/MiSePyNet.py`` (:5-163).

Same class names, constructor / forward signatures and ``state_dict`` keys (the ``nn.Conv3d`` / ``nn.BatchNorm3d`` children are
parameter containers; arithmetic goes through ``transmf_ad_b200.mnet_functional`` -> csrc/mnet_ops.cu).  The reference's quirk is
kept: ``spatial_cnn.forward`` applies ``self.conv1`` to all three slice outputs (:89-94); ``conv2`` / ``conv3`` exist only as
(dead) parameters so that checkpoints interchange.
"""
from __future__ import annotations

import torch
from torch import nn

from transmf_ad_b200 import mnet_functional as MF
from .networks import Linear


class slice_cnn(nn.Module):
    def __init__(self, dim) -> None:
        super().__init__()
        k2, k3 = (dim + 1) // 2, (dim + 2) // 3
        self.conv1 = nn.Sequential(nn.Conv3d(1, 8, kernel_size=(1, 1, dim)), nn.BatchNorm3d(8), nn.ReLU())
        self.conv2 = nn.Sequential(nn.Conv3d(1, 8, kernel_size=(1, 1, k2)), nn.BatchNorm3d(8), nn.ReLU(),
                                   nn.Conv3d(8, 8, kernel_size=(1, 1, k2)), nn.BatchNorm3d(8), nn.ReLU())
        self.conv3 = nn.Sequential(nn.Conv3d(1, 8, kernel_size=(1, 1, k3)), nn.BatchNorm3d(8), nn.ReLU(),
                                   nn.Conv3d(8, 8, kernel_size=(1, 1, k3)), nn.BatchNorm3d(8), nn.ReLU(),
                                   nn.Conv3d(8, 8, kernel_size=(1, 1, k3)), nn.BatchNorm3d(8), nn.ReLU())

    def _stack(self, seq, x):
        for i in range(0, len(seq), 3):
            x = MF.conv_bn_relu(x, seq[i], seq[i + 1], self.training)
        return x

    def forward(self, img):
        return self._stack(self.conv1, img), self._stack(self.conv2, img), self._stack(self.conv3, img)


class spatial_cnn(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.conv1 = nn.Sequential(
            nn.Conv3d(8, 16, kernel_size=(11, 11, 1), stride=(2, 2, 2)), nn.BatchNorm3d(16), nn.ReLU(), nn.MaxPool3d(kernel_size=(3, 3, 1)),
            nn.Conv3d(16, 32, kernel_size=(11, 11, 1)), nn.BatchNorm3d(32), nn.ReLU(), nn.MaxPool3d(kernel_size=(3, 3, 1)),
            nn.Conv3d(32, 64, kernel_size=(1, 1, 1)), nn.BatchNorm3d(64), nn.ReLU())
        self.conv2 = nn.Sequential(
            nn.Conv3d(8, 16, kernel_size=(7, 7, 1), stride=(2, 2, 2)), nn.BatchNorm3d(16), nn.ReLU(), nn.MaxPool3d(kernel_size=(2, 2, 1), padding=1),
            nn.Conv3d(16, 32, kernel_size=(7, 7, 1)), nn.BatchNorm3d(32), nn.ReLU(), nn.MaxPool3d(kernel_size=(2, 2, 1), padding=1),
            nn.Conv3d(32, 64, kernel_size=(7, 7, 1)), nn.BatchNorm3d(64), nn.ReLU(), nn.MaxPool3d(kernel_size=(2, 2, 1), padding=1))
        self.conv3 = nn.Sequential(
            nn.Conv3d(8, 16, kernel_size=(3, 3, 1), stride=(2, 2, 2)), nn.BatchNorm3d(16), nn.ReLU(), nn.MaxPool3d(kernel_size=(2, 2, 1), padding=1),
            nn.Conv3d(16, 32, kernel_size=(3, 3, 1)), nn.BatchNorm3d(32), nn.ReLU(), nn.MaxPool3d(kernel_size=(2, 2, 1), padding=1),
            nn.Conv3d(32, 64, kernel_size=(3, 3, 1)), nn.BatchNorm3d(64), nn.ReLU(),
            nn.Conv3d(64, 64, kernel_size=(3, 3, 1)), nn.BatchNorm3d(64), nn.ReLU(), nn.MaxPool3d(kernel_size=(2, 2, 1), padding=1))

    def _conv1(self, x):
        s = self.conv1
        x = MF.maxpool_xy(MF.conv_bn_relu(x, s[0], s[1], self.training), s[3])
        x = MF.maxpool_xy(MF.conv_bn_relu(x, s[4], s[5], self.training), s[7])
        return MF.conv_bn_relu(x, s[8], s[9], self.training)

    def forward(self, slices1, slices2, slices3):
        # reference :89-94 -- conv1 for all three inputs
        return self._conv1(slices1) + self._conv1(slices2) + self._conv1(slices3)


class MiSePyNet(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.slice_cnn_axial = slice_cnn(91)
        self.spatial_cnn_axial = spatial_cnn()
        self.slice_cnn_col = slice_cnn(109)
        self.spatial_cnn_col = spatial_cnn()
        self.slice_cnn_sag = slice_cnn(91)
        self.spatial_cnn_sag = spatial_cnn()

    def forward(self, img):
        if hasattr(img, "as_tensor"):
            img = img.as_tensor()
        B = img.shape[0]
        # the slice convolutions run along the LAST axis of each view: materialise the permuted views once (3.6 MB each)
        views = ((self.slice_cnn_axial, self.spatial_cnn_axial, img.contiguous()),
                 (self.slice_cnn_col, self.spatial_cnn_col, img.permute(0, 1, 2, 4, 3).contiguous()),
                 (self.slice_cnn_sag, self.spatial_cnn_sag, img.permute(0, 1, 4, 3, 2).contiguous()))
        feats = [spat(*slc(v)).reshape(B, -1) for slc, spat, v in views]
        return torch.cat(feats, dim=1)


class Mnet(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.mri = MiSePyNet()
        self.pet = MiSePyNet()
        self.fc = nn.Sequential(Linear(640, 512), nn.BatchNorm1d(512), nn.ReLU(), nn.Dropout(0.5),
                                Linear(512, 64), nn.BatchNorm1d(64), nn.ReLU(), nn.Dropout(0.5),
                                Linear(64, 2))

    def forward(self, mri, pet):
        return self.fc(torch.cat([self.mri(mri), self.pet(pet)], dim=-1))
