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
