"""Small models for smoke tests: an MLP and a tensor-parallel transformer MLP block."""
import torch
from torch import nn

from ..parallel.tensor_parallel import ColumnParallelLinear, RowParallelLinear


class MLP(nn.Module):
    def __init__(self, d=256, hidden=1024, depth=3, out=10):
        super().__init__()
        layers, cin = [], d
        for _ in range(depth):
            layers += [nn.Linear(cin, hidden), nn.GELU()]
            cin = hidden
        self.net = nn.Sequential(*layers, nn.Linear(cin, out))

    def forward(self, x):
        return self.net(x)


class TPTransformerBlock(nn.Module):
    """x + W2 gelu(W1 x) with W1 column- and W2 row-parallel (one allreduce forward, one backward)."""

    def __init__(self, d=512, hidden=2048, comm=None, dtype=None, device=None):
        super().__init__()
        self.norm = nn.LayerNorm(d, dtype=dtype, device=device)
        self.up = ColumnParallelLinear(d, hidden, comm=comm, dtype=dtype, device=device)
        self.down = RowParallelLinear(hidden, d, comm=comm, dtype=dtype, device=device)

    def forward(self, x):
        return x + self.down(torch.nn.functional.gelu(self.up(self.norm(x))))
