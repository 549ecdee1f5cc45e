"""Noise-model objectives and their far-field gradients
(reference: src/tike/operators/cupy/objective.py:11-124).  Device tensors in,
device tensors out; the fused solvers evaluate the same formulas in-kernel."""
from __future__ import annotations

import torch


def _gaussian(data, intensity):
    diff = torch.sqrt(intensity) - torch.sqrt(data)
    return diff * diff


def gaussian(data, intensity):
    return torch.mean(_gaussian(data, intensity))


def gaussian_grad(data, farplane, intensity):
    return farplane * (1 - torch.sqrt(data) /
                       (torch.sqrt(intensity) + 1e-9))[..., None, None, :, :]


def gaussian_each_pattern(data, intensity):
    return torch.mean(_gaussian(data, intensity), dim=(-2, -1))


def _poisson(data, intensity):
    return intensity - data * torch.log(intensity + 1e-9)


def poisson(data, intensity):
    return torch.mean(_poisson(data, intensity))


def poisson_grad(data, farplane, intensity):
    return farplane * (1 - data / (intensity + 1e-9))[..., None, None, :, :]


def poisson_each_pattern(data, intensity):
    return torch.mean(_poisson(data, intensity), dim=(-2, -1))
