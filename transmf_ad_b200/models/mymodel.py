"""Model assemblies of TransMF_AD on the B200 kernels -- drop-in for the reference ``models/mymodel.py``.

Same class names, constructor / forward signatures, return arity and ``state_dict`` keys as the reference
(``model_single`` :13-37, ``model_CNN`` :40-66, ``model_transformer`` :69-98, ``model_transformer_res`` :101-141,
``model_CNN_ad`` :144-179, ``model_ad`` :182-222), so ``kfold_train_adversarial.py`` / ``kfold_train_single.py``
run them unchanged.  The two sNet towers are executed as one grouped launch sequence.
"""
from __future__ import annotations

import torch
from torch import nn

from transmf_ad_b200 import functional as TF
from .gradient_reversal import revgrad
from .networks import (CrossTransformer, CrossTransformer_MOD_AVG, Linear, sNet, snet_pair_forward, tokens_of)

GRL_LAMBDA = 2.0          # reference mymodel.py:167,209  alpha = torch.Tensor([2])


def _init_cnn_weights(module):
    """kaiming-normal (fan_out, relu) conv weights; BatchNorm3d weight 1 / bias 0 (reference mymodel.py:195-202)."""
    for m in module.modules():
        if isinstance(m, nn.Conv3d):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        elif isinstance(m, nn.BatchNorm3d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


class _GapFlatten(nn.Module):
    """AdaptiveAvgPool3d(1) + 'b c x y z -> b (c x y z)' as one token-mean kernel (reference mymodel.py:193)."""

    def forward(self, feat):
        return TF.token_pool(tokens_of(feat), True, False)


def _mlp_head_transformer(in_dim):
    """Linear BN1d ReLU Dropout(.5) Linear BN1d ReLU Dropout(.5) Linear (reference mymodel.py:190-192).
    Dropout stays torch's so that its Philox stream is the reference's."""
    return nn.Sequential(Linear(in_dim, 512), nn.BatchNorm1d(512), nn.ReLU(), nn.Dropout(0.5),
                         Linear(512, 64), nn.BatchNorm1d(64), nn.ReLU(), nn.Dropout(0.5),
                         Linear(64, 2))


def _discriminator(dim):
    return nn.Sequential(Linear(dim, 128), nn.BatchNorm1d(128), nn.ReLU(), Linear(128, 2))


class model_single(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.cnn = sNet(dim)
        self.avgpool = _GapFlatten()
        self.fc = nn.Sequential(Linear(128, 64), nn.ReLU(), Linear(64, 2))
        _init_cnn_weights(self)

    def forward(self, img):
        return self.fc(self.avgpool(self.cnn(img)))


class model_CNN(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.mri_cnn = sNet(dim)
        self.pet_cnn = sNet(dim)
        self.transform = _GapFlatten()
        self.fc = nn.Sequential(Linear(dim * 2, 128), nn.ReLU(), Linear(128, 2))
        _init_cnn_weights(self)

    def forward(self, mri, pet):
        fm, fp = snet_pair_forward(self.mri_cnn, self.pet_cnn, mri, pet)
        return self.fc(torch.cat([self.transform(fm), self.transform(fp)], dim=1))


class model_transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout):
        super().__init__()
        self.mri_cnn = sNet(dim)
        self.pet_cnn = sNet(dim)
        self.fuse_transformer = CrossTransformer_MOD_AVG(dim, depth, heads, dim_head, mlp_dim, dropout)
        self.fc_cls = _mlp_head_transformer(dim * 4)
        _init_cnn_weights(self)

    def forward(self, mri, pet):
        fm, fp = snet_pair_forward(self.mri_cnn, self.pet_cnn, mri, pet)
        return self.fc_cls(self.fuse_transformer(tokens_of(fm), tokens_of(fp)))


class model_transformer_res(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout):
        super().__init__()
        self.mri_cnn = sNet(dim)
        self.pet_cnn = sNet(dim)
        self.fuse_transformer = CrossTransformer(dim, depth, heads, dim_head, mlp_dim, dropout)
        self.fc_cls = nn.Sequential(Linear(dim * 2, 512), nn.ReLU(), nn.Dropout(0.5),
                                    Linear(512, 64), nn.ReLU(), nn.Dropout(0.5),
                                    Linear(64, 2))
        self.gap = nn.Identity()      # parameter-free in the reference (Rearrange + AdaptiveAvgPool1d); kept for attribute parity
        self.gmp = nn.Identity()
        _init_cnn_weights(self)

    def forward(self, mri, pet):
        fm, fp = snet_pair_forward(self.mri_cnn, self.pet_cnn, mri, pet)
        tm, tp = tokens_of(fm), tokens_of(fp)
        fused_m, fused_p = self.fuse_transformer(tm, tp)
        cls = torch.cat([TF.token_pool(fused_m + tm, True, False), TF.token_pool(fused_p + tp, True, False)], dim=1)
        return self.fc_cls(cls)


class model_CNN_ad(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.mri_cnn = sNet(dim)
        self.pet_cnn = sNet(dim)
        self.fc_cls = nn.Sequential(Linear(dim * 2, 128), nn.ReLU(), Linear(128, 2))
        self.gap = _GapFlatten()
        self.D = _discriminator(dim)
        _init_cnn_weights(self)

    def forward(self, mri, pet):
        fm, fp = snet_pair_forward(self.mri_cnn, self.pet_cnn, mri, pet)
        gm, gp = self.gap(fm), self.gap(fp)
        D_MRI_logits = self.D(revgrad(gm, GRL_LAMBDA))          # MRI first, then PET: BatchNorm1d running-stat order
        D_PET_logits = self.D(revgrad(gp, GRL_LAMBDA))
        output_logits = self.fc_cls(torch.cat([gm, gp], dim=1))
        return output_logits, D_MRI_logits, D_PET_logits


class model_ad(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout):
        super().__init__()
        self.mri_cnn = sNet(dim)
        self.pet_cnn = sNet(dim)
        self.fuse_transformer = CrossTransformer_MOD_AVG(dim, depth, heads, dim_head, mlp_dim, dropout)
        self.fc_cls = _mlp_head_transformer(dim * 4)
        self.gap = _GapFlatten()
        self.D = _discriminator(dim)
        _init_cnn_weights(self)

    def forward(self, mri, pet):
        fm, fp = snet_pair_forward(self.mri_cnn, self.pet_cnn, mri, pet)
        D_MRI_logits = self.D(revgrad(self.gap(fm), GRL_LAMBDA))
        D_PET_logits = self.D(revgrad(self.gap(fp), GRL_LAMBDA))
        output_pos = self.fuse_transformer(tokens_of(fm), tokens_of(fp))
        output_logits = self.fc_cls(output_pos)
        return output_logits, D_MRI_logits, D_PET_logits
