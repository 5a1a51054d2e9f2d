"""Helpers shared by the GPU parity tests."""
import numpy as np
import torch

from oracle import pyg_gatconv as O


def seeded_params(K, H, C, concat=False, seed=1, bias_scale=0.1):
    """glorot weights exactly as reset_parameters() draws them, under a fixed seed (plus a non-zero bias so
    that the bias path is exercised)."""
    g = torch.Generator().manual_seed(seed)
    W = O.glorot_(torch.empty(H * C, K), g)
    a_s = O.glorot_(torch.empty(1, H, C), g)
    a_d = O.glorot_(torch.empty(1, H, C), g)
    b = torch.randn(H * C if concat else C, generator=g) * bias_scale
    return W, a_s, a_d, b


def load_ckpt(model, path):
    npz = np.load(path)
    sd = {k: torch.from_numpy(npz[k]) for k in npz.files}
    for k in list(sd):
        if k.endswith("lin_src.weight"):
            sd[k.replace("lin_src", "lin_dst")] = sd[k]
    model.load_state_dict(sd, strict=True)
    return model


def maxabs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max()) if a.numel() else 0.0


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
