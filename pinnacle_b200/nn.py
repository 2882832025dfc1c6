"""Minimal host mirror of ``dde.nn.FNN`` (deepxde/nn/pytorch/fnn.py:9-48) so the fused path can be
driven where the reference package is not installed (GPU box, bench).  Same attribute names
(``linears``, ``activation``), same initialisation (Glorot normal weights, zero biases,
fnn.py:25-33 / nn/initializers.py:132-139), so parameters remain ordinary ``nn.Linear`` tensors and
checkpoints (model.py:933-951) interchange with the reference net."""
from __future__ import annotations

import torch


class FNN(torch.nn.Module):
    def __init__(self, layer_sizes, activation="tanh", kernel_initializer="Glorot normal"):
        super().__init__()
        if activation != "tanh":
            raise ValueError("the fused path supports activation='tanh' only")
        if kernel_initializer != "Glorot normal":
            raise ValueError("only 'Glorot normal' is mirrored")
        self.activation = torch.tanh
        self._input_transform = None
        self._output_transform = None
        self.regularizer = None
        self.auxiliary_vars = None
        self.linears = torch.nn.ModuleList()
        for i in range(1, len(layer_sizes)):
            lin = torch.nn.Linear(layer_sizes[i - 1], layer_sizes[i], dtype=torch.float32)
            torch.nn.init.xavier_normal_(lin.weight)
            torch.nn.init.zeros_(lin.bias)
            self.linears.append(lin)

    def forward(self, inputs):
        x = inputs
        for lin in self.linears[:-1]:
            x = self.activation(lin(x))
        return self.linears[-1](x)


class _LAAFlayer(torch.nn.Module):
    """tanh(n * a * (W x + b)) (src/model/laaf.py:5-22)."""

    def __init__(self, n, a, dim_in, dim_out):
        super().__init__()
        self.n, self.a = n, a
        self.activation = torch.nn.Tanh()
        self.fc = torch.nn.Linear(dim_in, dim_out)

    def forward(self, x):
        return self.activation(self.n * torch.mul(self.a, self.fc(x)))


class DNN_LAAF(torch.nn.Module):
    """Host mirror of the reference's layer-wise adaptive-activation net (src/model/laaf.py:25-51; benchmark.py --method laaf):
    same constructor, same attribute names (``a`` [n_layers, n_hidden], ``layers`` = Sequential of layer0..layerN), same
    initialisation (xavier-uniform slopes with gain 1.4, default nn.Linear init)."""
    per_neuron = True

    def __init__(self, n_layers, n_hidden, x_dim=1, u_dim=1):
        super().__init__()
        self.x_dim, self.u_dim, self.n_layers, self.n_hidden = x_dim, u_dim, n_layers, n_hidden
        self.regularizer = None
        self.a = torch.nn.Parameter(torch.empty(n_layers, n_hidden if self.per_neuron else 1))
        torch.nn.init.xavier_uniform_(self.a.data, gain=1.4)
        from collections import OrderedDict
        sel = (lambda i: self.a[i, :]) if self.per_neuron else (lambda i: self.a[i])
        layers = [("layer0", _LAAFlayer(10, sel(0), x_dim, n_hidden))]
        for i in range(n_layers - 1):
            layers.append(("layer%d" % (i + 1), _LAAFlayer(10, sel(i + 1), n_hidden, n_hidden)))
        layers.append(("layer%d" % n_layers, torch.nn.Linear(n_hidden, u_dim)))
        self.layers = torch.nn.Sequential(OrderedDict(layers))

    def forward(self, x):
        return self.layers(x)

    def _apply(self, fn, *args, **kwargs):
        # the layers hold VIEWS of self.a taken at construction (as in the reference); re-take them after .to(device)/.float()
        out = super()._apply(fn, *args, **kwargs)
        for i, lay in enumerate(list(self.layers)[:-1]):
            lay.a = self.a[i, :] if self.per_neuron else self.a[i]
        return out


class DNN_GAAF(DNN_LAAF):
    """Global (one slope per layer) variant, src/model/laaf.py:54-80 (benchmark.py --method gaaf)."""
    per_neuron = False
