import torch


def meshgrid_ij(*tensors):
    return torch.meshgrid(*tensors, indexing="ij")
