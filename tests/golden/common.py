"""Shared helpers for the golden fixtures (deterministic weights / index sets)."""
from __future__ import annotations

import os

import numpy as np

GOLDEN_DIR = os.path.dirname(os.path.abspath(__file__))
N_POINTS = 77            # ragged on purpose: not a multiple of any CUDA tile size
N_HIDDEN_SAMPLES = 1024
WEIGHT_SCALE = 1.25      # a little beyond Glorot init so tanh leaves its linear regime
BIAS_STD = 0.1


def golden_params(layer_shapes, seed: int, dtype=np.float64) -> np.ndarray:
    """Flat parameter vector in nn.Linear order (W0, b0, W1, b1, ...), Glorot-normal-like weights
    scaled by WEIGHT_SCALE and N(0, BIAS_STD^2) biases, from the frozen legacy MT19937 stream."""
    rs = np.random.RandomState(seed)
    parts = []
    for (n_out, n_in) in layer_shapes:
        std = WEIGHT_SCALE * np.sqrt(2.0 / (n_in + n_out))
        parts.append((rs.standard_normal((n_out, n_in)) * std).reshape(-1))
        parts.append(rs.standard_normal(n_out) * BIAS_STD)
    return np.concatenate(parts).astype(dtype)


def golden_index_set(layer_shapes, seed: int) -> np.ndarray:
    """Indices into the flat gradient that the fixture stores: every entry of the first layer, the
    last layer and all biases, plus N_HIDDEN_SAMPLES random entries of each hidden weight matrix."""
    rs = np.random.RandomState(seed + 1)
    idx, off = [], 0
    n_layers = len(layer_shapes)
    for li, (n_out, n_in) in enumerate(layer_shapes):
        nw = n_out * n_in
        if li == 0 or li == n_layers - 1 or nw <= N_HIDDEN_SAMPLES:
            idx.append(off + np.arange(nw))
        else:
            idx.append(off + np.sort(rs.choice(nw, N_HIDDEN_SAMPLES, replace=False)))
        off += nw
        idx.append(off + np.arange(n_out))
        off += n_out
    return np.concatenate(idx).astype(np.int64)


def probe_vectors(n_params: int, seed: int, n: int = 4) -> np.ndarray:
    rs = np.random.RandomState(seed + 2)
    return rs.standard_normal((n, n_params))


def fixture_path(name: str) -> str:
    return os.path.join(GOLDEN_DIR, f"{name}.npz")


def class_seed(name: str) -> int:
    return sum((i + 1) * ord(c) for i, c in enumerate(name)) % 100000
