"""Reduced forms of large gradient tensors for the golden fixtures (test infrastructure, like everything under oracle/).

A conv weight gradient at the shipped AMFT shape is [512, 512, 3, 3] = 9.4 MB in fp32; the fixtures keep instead
  rows   sum over (Cin, ky, kx)      [Cout]     -- sensitive to every output channel
  cols   sum over (Cout, ky, kx)     [Cin]      -- sensitive to every input channel
  taps   sum over (Cout, Cin)        [3, 3]     -- sensitive to the tap geometry (flips / transposes)
  pick   4096 individual elements at seeded positions
which together pin any structured error (wrong tap, transposed channels, missing split-K partial, wrong scale) at a
few KB per tensor.
"""
import numpy as np
import torch


def _positions(numel, seed, n=4096):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (min(n, numel),), generator=g)


def weight_grad_digest(G, seed):
    G = torch.as_tensor(np.asarray(G.detach().cpu() if torch.is_tensor(G) else G)).double()
    assert G.dim() == 4
    return {
        "rows": G.sum((1, 2, 3)).float().numpy(),
        "cols": G.sum((0, 2, 3)).float().numpy(),
        "taps": G.sum((0, 1)).float().numpy(),
        "pick": G.reshape(-1)[_positions(G.numel(), seed)].float().numpy(),
    }
