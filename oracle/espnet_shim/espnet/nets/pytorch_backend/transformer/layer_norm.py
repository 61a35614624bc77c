"""Stub of espnet's LayerNorm (reference import: variance_predictor.py:10).
eps is 1e-12; `dim` selects the normalised axis by transposing it last."""
import torch


class LayerNorm(torch.nn.LayerNorm):
    def __init__(self, nout, dim=-1):
        super().__init__(nout, eps=1e-12)
        self.dim = dim

    def forward(self, x):
        if self.dim == -1:
            return super().forward(x)
        return super().forward(x.transpose(1, -1)).transpose(1, -1)
