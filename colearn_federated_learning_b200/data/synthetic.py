"""Synthetic UNSW-2018 IoT-Botnet-shaped data (there is no network for the real dataset).

``BASELINE.json``: "synthetic UNSW-IoT-shaped data / random-init weights".  The generator
produces the 10 model features already MinMax-scaled to [0, 1] with a binary ``attack`` label
that is a noisy function of a few features (so that loss-decrease tests are meaningful), and can
also write a CSV with the exact Bot-IoT "10-best" 19-column header so the whole CSV → dataset →
training path is exercised (reference example file: ``dataset_example/*.csv``).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from .datasets import CSV_HEADER, FEATURE_COLUMNS


def synthetic_unsw(n: int, seed: int = 0, n_features: int = 10, noise: float = 0.05,
                   device: Optional[torch.device] = None,
                   label_dtype: torch.dtype = torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns ``(X[n, n_features] in [0,1], y[n,1] in {0,1})``; deterministic in ``seed``."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, n_features, generator=g)
    w = torch.linspace(-1.5, 1.5, n_features)
    logit = (x - 0.5) @ w * 4.0 + noise * torch.randn(n, generator=g)
    y = (logit > 0).to(label_dtype).view(n, 1)
    if device is not None:
        x, y = x.to(device), y.to(device)
    return x, y


def synthetic_images(n: int, seed: int = 0, size: int = 32, classes: int = 10,
                     device: Optional[torch.device] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Synthetic ``[n,3,size,size]`` images with class-dependent mean (BASELINE config 4)."""
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, classes, (n,), generator=g)
    x = torch.randn(n, 3, size, size, generator=g) * 0.5
    x += (y.float().view(n, 1, 1, 1) / classes - 0.5)
    if device is not None:
        x, y = x.to(device), y.to(device)
    return x, y


def synthetic_for_model(model: str, n: int, seed: int = 0,
                        device: Optional[torch.device] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Synthetic data of the shape ``model`` consumes (``--synthetic N`` on the CLIs): 32x32 images with 10 classes
    for ``resnet18``, 28x28 single-channel digits-shaped inputs for ``net`` (784-128-64-10), two XOR-like features
    for ``testing_remote``, the ten UNSW-IoT features otherwise.  Labels are ``[n, 1]`` floats (class index for the
    cross-entropy models), which is what every fit executor accepts."""
    if model == "resnet18":
        x, y = synthetic_images(n, seed=seed)
        y = y.float().view(n, 1)
    elif model == "net":
        x, y = synthetic_images(n, seed=seed, size=28)
        x, y = x[:, :1].contiguous(), y.float().view(n, 1)
    elif model == "testing_remote":
        x, y = synthetic_unsw(n, seed=seed, n_features=2)
    else:
        x, y = synthetic_unsw(n, seed=seed)
    if device is not None:
        x, y = x.to(device), y.to(device)
    return x, y


def write_synthetic_csv(path: str, n: int, seed: int = 0, attack_fraction: Optional[float] = None) -> str:
    """Write a Bot-IoT-shaped CSV (19 columns, header identical to the reference example)."""
    import pandas as pd

    rng = np.random.default_rng(seed)
    x, y = synthetic_unsw(n, seed)
    x, y = x.numpy(), y.numpy().reshape(-1).astype(int)
    if attack_fraction is not None:
        y = (rng.random(n) < attack_fraction).astype(int)
    scale = {"seq": 262212, "stddev": 2.5, "N_IN_Conn_P_SrcIP": 100, "min": 5.0, "state_number": 11,
             "mean": 5.0, "N_IN_Conn_P_DstIP": 100, "drate": 60.0, "srate": 1000.0, "max": 5.0}
    df = pd.DataFrame({c: x[:, i] * scale[c] for i, c in enumerate(FEATURE_COLUMNS)})
    for c in ("seq", "N_IN_Conn_P_SrcIP", "state_number", "N_IN_Conn_P_DstIP"):
        df[c] = df[c].round().astype(int)
    df["pkSeqID"] = np.arange(1, n + 1)
    df["proto"] = rng.choice(["tcp", "udp"], n)
    df["saddr"] = [f"192.168.100.{rng.integers(1, 254)}" for _ in range(n)]
    df["sport"] = rng.integers(1024, 65535, n)
    df["daddr"] = [f"192.168.100.{rng.integers(1, 10)}" for _ in range(n)]
    df["dport"] = rng.choice([80, 53, 443, 8080], n)
    df["attack"] = y
    df["category"] = np.where(y == 1, rng.choice(["DDoS", "DoS", "Reconnaissance"], n), "Normal")
    df["subcategory"] = np.where(y == 1, rng.choice(["TCP", "UDP", "HTTP"], n), "Normal")
    df[CSV_HEADER].to_csv(path, index=False, float_format="%.6f")
    return path
