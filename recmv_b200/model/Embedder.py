"""Positional encoding with per-band annealing weights -- API of model/Embedder.py:4-65.

Host-side (torch) form used by the autograd-composite paths; the fused kernels evaluate the same
encoding in their prologue (csrc/common.cuh positional_encode)."""
import numpy as np
import torch


def annealing_weights(multires, ratio):
    """utils/utils.py:40-46."""
    alpha = ratio * multires
    out = []
    for ind in range(multires):
        w = (1. - np.cos(np.pi * min(max(alpha - float(ind), 0.), 1.))) / 2.
        out.extend([w, w])
    return out


def ratio_to_weights(multires, ratio):
    """model/network.py:93-99: None -> ones, <= 0 -> zeros, else annealed."""
    if ratio is None:
        return [1.0] * (2 * multires)
    if ratio <= 0:
        return [0.0] * (2 * multires)
    return annealing_weights(multires, ratio)


class Embedder:
    def __init__(self, **kwargs):
        self.kwargs = kwargs
        d = kwargs['input_dims']
        self.include_input = kwargs['include_input']
        n = kwargs['num_freqs']
        mx = kwargs['max_freq_log2']
        if kwargs['log_sampling']:
            self.freq_bands = 2. ** torch.linspace(0., mx, n)
        else:
            self.freq_bands = torch.linspace(2. ** 0., 2. ** mx, n)
        self.periodic_fns = kwargs['periodic_fns']
        self.out_dim = (d if self.include_input else 0) + d * n * len(self.periodic_fns)

    def embed(self, inputs, ws=None):
        outs = [inputs] if self.include_input else []
        i = 0
        for freq in self.freq_bands:
            for fn in self.periodic_fns:
                w = 1. if ws is None else ws[i]
                outs.append(w * fn(inputs * freq))
                i += 1
        return torch.cat(outs, -1)


def get_embedder(multires):
    eo = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                  log_sampling=True, periodic_fns=[torch.sin, torch.cos])

    def embed(x, ws=None, eo=eo):
        return eo.embed(x, ws)
    return embed, eo.out_dim
