"""mvdfusion/embedder.py of the reference: sinusoidal timestep embedding + the (unused on the path, but
state-dict-bearing) TimestepEmbedder / RayEmbedder that GridAttn constructs."""
import math

import torch
import torch.nn as nn


def timestep_embedding(timesteps, dim, max_period=10000, repeat_only=False):
    """mvdfusion/embedder.py:114-134 — host helper (the in-loop version is mvd_timestep_embedding)."""
    if repeat_only:
        return timesteps[:, None].expand(-1, dim)
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


class TimestepEmbedder(nn.Module):
    """mvdfusion/embedder.py:73-110: owned by GridAttn as `t_embedder`, never called in its forward; kept so that
    `view_attn.t_embedder.mlp.*` exists in the state dict."""

    def __init__(self, hidden_size, frequency_embedding_size=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size, bias=True), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size, bias=True))
        self.frequency_embedding_size = frequency_embedding_size


class RayEmbedder(nn.Module):
    """mvdfusion/embedder.py:12-69: parameter-free, constructed by GridAttn and never called on the path."""

    def __init__(self, input_size, n_harmonic=7, omega0=0.1):
        super().__init__()
        self.input_size = input_size
        self.output_dim = 6 * (2 * n_harmonic + 1)
