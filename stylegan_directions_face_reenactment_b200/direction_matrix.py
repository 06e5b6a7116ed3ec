"""The learned direction matrix A: R^k -> W+ offsets (reference libs/models/direction_matrix.py:6-48).

Stays in PyTorch (one addmm); its gradient (weight [out,in] + bias [out]) is the payload of the multi-GPU
all-reduce (dist.py).  Same constructor, attributes, state_dict keys (`linear.weight`, `linear.bias`) and forward
semantics as the reference; does not depend on the NumPy-1-only `np.product` the reference calls at :11-12.
"""
import math

import torch
from torch import nn


def _prod(v):
    return int(math.prod(v)) if isinstance(v, (tuple, list)) else int(v)


class DirectionMatrix(nn.Module):
    def __init__(self, shift_dim, input_dim=None, out_dim=None, inner_dim=512, bias=True, w_plus=False,
                 num_layers=14, initialization='normal'):
        super().__init__()
        self.shift_dim = shift_dim
        self.input_dim = input_dim if input_dim is not None else _prod(shift_dim)
        self.out_dim = out_dim if out_dim is not None else _prod(shift_dim)
        self.w_plus = w_plus
        self.num_layers = num_layers
        total_out = self.out_dim * num_layers if w_plus else self.out_dim
        self.linear = nn.Linear(self.input_dim, total_out, bias=bias)
        with torch.no_grad():
            self.linear.weight.zero_()
            if initialization == 'normal':
                nn.init.normal_(self.linear.weight, mean=0.0, std=0.03)
            elif initialization == 'eye':
                m = int(min(self.input_dim, total_out if not w_plus else self.out_dim))
                reps = num_layers if w_plus else 1
                for r in range(reps):
                    self.linear.weight[r * self.out_dim:r * self.out_dim + m, :m] = torch.eye(m)

    def forward(self, input):
        x = input.view(-1, self.input_dim)
        out = self.linear(x)
        if self.w_plus:
            out = out.view(x.shape[0], self.num_layers, self.shift_dim)
        return out
