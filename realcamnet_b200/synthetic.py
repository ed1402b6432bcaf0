"""Seeded synthetic inputs and name-keyed deterministic weights for benchmarks, tools, fixtures and tests.

Not on the compute path: nothing here touches the CUDA library.  The reference ships no checkpoint and no data, so every
measurement and parity check of this repository runs on

* inputs: raw = uniform [0,1) packed-Bayer (B,4,T,T); cond = uniform (B,4,256,256) for single tiles;
  coord = 2-channel normalised pixel-centre coordinates, channel 0 = x in [-1,1] along W, channel 1 = y in [-1,1] along H (the
  reference smoke tests feed randn triples of the same shapes, LiteISP.py:2670-2672);
* weights that are a pure function of (parameter name, shape, seed): the same call fills the real reference model (in the
  authoring container, when generating tests/golden), the functional oracle and the CUDA product, wherever they run.  Scales are
  chosen so activations stay O(1) through the ~100-layer path (parity tests are sensitive to every layer) and so the entropy
  model sees non-trivial symbols.
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch


def coord_map(T, B=1, y0=-1.0, y1=1.0, x0=-1.0, x1=1.0):
    ys = torch.linspace(y0, y1, T)
    xs = torch.linspace(x0, x1, T)
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([xx, yy])[None].repeat(B, 1, 1, 1).contiguous()


def make_inputs(T, seed=1234, B=1, cond_size=256):
    g = torch.Generator().manual_seed(seed)
    raw = torch.rand(B, 4, T, T, generator=g)
    cond = torch.rand(B, 4, cond_size, cond_size, generator=g)
    return [raw, cond, coord_map(T, B)]


_SKIP = ("pedestal", "lower_bound", "likelihood_lower_bound", "target", "scale_table", "scale_bound",
         "_offset", "_quantized_cdf", "_cdf_length", "num_batches_tracked")


def _rng(name: str, seed: int):
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return np.random.default_rng(int.from_bytes(h[:8], "little"))


def make_tensor(name: str, shape, seed: int = 0, gain: float = 0.7):
    """Returns the deterministic value for parameter/buffer ``name`` or None to keep as is."""
    last = name.split(".")[-1]
    if any(s in name for s in _SKIP):
        return None
    g = _rng(name, seed)
    shape = tuple(shape)
    n = int(np.prod(shape)) if len(shape) else 1
    u = g.uniform(-1.0, 1.0, size=n).astype(np.float32).reshape(shape)
    if last == "beta":  # GDN beta (reparametrised): sqrt(b + pedestal), b in [0.8, 1.2]
        ped = (2.0 ** -18) ** 2
        return torch.from_numpy(np.sqrt(1.0 + 0.2 * u + ped).astype(np.float32))
    if last == "gamma":  # GDN gamma: sqrt(0.1*I + small non-negative off-diagonal + pedestal)
        ped = (2.0 ** -18) ** 2
        C = shape[0]
        gm = 0.1 * np.eye(C, dtype=np.float32) + (0.02 / np.sqrt(C)) * np.abs(u)
        return torch.from_numpy(np.sqrt(gm + ped).astype(np.float32))
    if last.startswith("_matrix"):
        return "matrix"  # handled by caller (perturb the constant default init)
    if last.startswith("_bias"):
        return torch.from_numpy(0.5 * u)
    if last.startswith("_factor"):
        return torch.from_numpy(0.2 * u)
    if last == "quantiles":  # (C,1,3): [lo, median, hi]
        q = np.zeros(shape, dtype=np.float32)
        med = 0.3 * u[..., 1]
        q[..., 1] = med
        q[..., 0] = med - (6.0 + 4.0 * np.abs(u[..., 0]))
        q[..., 2] = med + (6.0 + 4.0 * np.abs(u[..., 2]))
        return torch.from_numpy(q)
    if last == "relative_position_params":
        return torch.from_numpy(0.5 * u)
    if last == "running_mean":
        return torch.from_numpy(0.1 * u)
    if last == "running_var":
        return torch.from_numpy(1.0 + 0.3 * u)
    if last == "bias":
        if "cc_scale_transforms" in name and name.endswith(".4.bias"):
            # spread the predicted scales over the CDF table: log-uniform in [0.08, 12]
            return torch.from_numpy(np.exp(2.5 * u).astype(np.float32))
        return torch.from_numpy(0.1 * u)
    if last == "weight":
        if len(shape) <= 1:  # LayerNorm / InstanceNorm / BatchNorm affine
            return torch.from_numpy(1.0 + 0.1 * u)
        fan_in = int(np.prod(shape[1:]))
        bound = gain * np.sqrt(3.0 / fan_in)  # std = gain / sqrt(fan_in)
        return torch.from_numpy((bound * u).astype(np.float32))
    return None


def fill_(module_or_sd, seed: int = 0, gain: float = 0.7):
    """In-place fill of an nn.Module (parameters + buffers) or a state-dict-like mapping."""
    sd = module_or_sd.state_dict() if hasattr(module_or_sd, "state_dict") else module_or_sd
    with torch.no_grad():
        for name in sorted(sd.keys()):
            t = sd[name]
            if not torch.is_floating_point(t):
                continue
            # Haar DWT filters are fixed constants of the architecture (networks.py:224-249)
            if _is_fixed_dwt(name, t):
                continue
            v = make_tensor(name, t.shape, seed, gain)
            if v is None:
                continue
            if isinstance(v, str):  # "_matrixN": keep the reference's constant init, perturb by 10 %
                u = _rng(name, seed).uniform(-1.0, 1.0, size=tuple(t.shape)).astype(np.float32)
                t.add_(torch.from_numpy(0.1 * u).to(t.device) * t.abs().clamp_min(0.1))
                continue
            t.copy_(v.to(t.device, t.dtype))
    return module_or_sd


def _is_fixed_dwt(name: str, t: torch.Tensor) -> bool:
    if not name.endswith(".weight") or t.dim() != 4 or tuple(t.shape[1:]) != (1, 2, 2):
        return False
    return bool(torch.all(t.abs() == 0.5))


def checksum(sd) -> dict:
    """Order-independent fingerprint of a state dict (used to prove two models hold equal weights)."""
    tot, n = 0.0, 0
    for k in sorted(sd.keys()):
        t = sd[k]
        if torch.is_floating_point(t) and not any(x in k for x in _SKIP):
            tot += float(t.double().abs().sum())
            n += t.numel()
    return {"abs_sum": tot, "numel": n, "tensors": len(sd)}
