"""Encrypted MLP training step on secret shares (reference ``encrypted_training`` cf.py:130-170).

Forward: ``z = W·a + b`` via Beaver matmul + truncation; ReLU through shared sign bits; the
sigmoid head is the cubic ``0.5 + x/4 - x^3/48`` (PySyft also approximates sigmoid under SMPC).
Loss: ``((out - target)**2).sum().refresh() / batch`` (cf.py:158).  Backward is the same manual
back-propagation as ``ops.reference.mlp_backward`` but on shares; SGD uses a fixed-precision
learning rate (``optimizer.fix_precision()``, fc.py:443-444).
"""
from __future__ import annotations

from typing import Sequence

import torch

from .sharing import BASE, CryptoProvider, SharedTensor, fix_precision, float_precision, share


class SharedMLP:
    def __init__(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], provider: CryptoProvider,
                 sigmoid_out: bool) -> None:
        self.provider = provider
        self.sigmoid_out = sigmoid_out
        self.W = [share(fix_precision(w), provider) for w in weights]   # [out, in], scale BASE
        self.b = [share(fix_precision(b), provider) for b in biases]

    @classmethod
    def from_module(cls, model, provider: CryptoProvider) -> "SharedMLP":
        layers = model.layers()
        return cls([l.weight.detach() for l in layers], [l.bias.detach() for l in layers], provider,
                   model.spec.out_activation == "sigmoid")

    def reveal_into(self, model) -> None:
        """``model.get().float_precision()`` (fc.py:454)."""
        with torch.no_grad():
            for layer, w, b in zip(model.layers(), self.W, self.b):
                wv, bv = w.get(), b.get()          # (distributed demo: only the dealer gets the plaintext, the parties None)
                if wv is not None:
                    layer.weight.copy_(float_precision(wv).to(layer.weight.device))
                    layer.bias.copy_(float_precision(bv).to(layer.bias.device))

    # ------------------------------------------------------------------------------------------
    def forward(self, x: SharedTensor):
        acts, masks = [x], []
        h = x
        for li, (w, b) in enumerate(zip(self.W, self.b)):
            z = h.matmul(w.t()).truncate() + b          # scale BASE
            if li < len(self.W) - 1:
                m = z.positive_bit()                      # 0/1, unscaled
                h = z.mul(m)
                masks.append(m)
                acts.append(h)
            else:
                h = z
        pre = h
        if self.sigmoid_out:
            x2 = pre.mul(pre).truncate()
            x3 = x2.mul(pre).truncate()
            out = pre.mul_public(BASE // 4).truncate() - x3.mul_public(round(BASE / 48)).truncate() + (BASE // 2)
            dact = (-x2.mul_public(round(BASE / 16)).truncate()) + (BASE // 4)   # derivative of the cubic
        else:
            out, dact = pre, None
        return out, acts, masks, dact

    def step(self, x: SharedTensor, y: SharedTensor, lr: float) -> SharedTensor:
        """One SGD step; returns the shared loss value (scale BASE)."""
        bsz = x.shape[0]
        out, acts, masks, dact = self.forward(x)
        diff = out - y
        loss = diff.mul(diff).truncate().sum().refresh()
        loss = loss.truncate(bsz) if bsz > 1 else loss
        dz = diff.mul_public(2)
        if bsz > 1:
            dz = dz.truncate(bsz)
        if dact is not None:
            dz = dz.mul(dact).truncate()
        lr_fp = int(round(lr * BASE))
        for li in range(len(self.W) - 1, -1, -1):
            a = acts[li]
            gw = dz.t().matmul(a).truncate()
            gb = dz.sum(0)
            if li > 0:
                dh = dz.matmul(self.W[li]).truncate()
                dz_next = dh.mul(masks[li - 1])
            self.W[li] = self.W[li] - gw.mul_public(lr_fp).truncate()
            self.b[li] = self.b[li] - gb.mul_public(lr_fp).truncate()
            if li > 0:
                dz = dz_next
        return loss


def encrypted_sgd_step(model: SharedMLP, x: SharedTensor, y: SharedTensor, lr: float) -> SharedTensor:
    return model.step(x, y, lr)
