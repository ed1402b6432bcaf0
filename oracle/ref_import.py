"""Import the UNMODIFIED reference (``/root/reference/models/*.py``) in the authoring container.

TEST ORACLE ONLY.  The reference cannot be imported as shipped:
it hard-imports ``compressai``, ``timm``, ``thop``, ``ipdb`` (absent offline) and three
in-repo modules that were never committed (``models/cbam.py``, ``models/AWISP_utils.py``,
``models/AWISP_modules.py``; models/LiteISP.py:3,13,14).  ``install_shims()`` registers stub
modules for those names in ``sys.modules`` -- ``compressai.*`` resolves to the restatement in
``oracle/cai.py`` -- after which the five reference files import without modification.

``/root/reference`` does not exist on the GPU box; there the modules come from ``oracle/_ref`` (the sourceless byte-code that
``oracle/stage_ref.py`` compiles in the authoring container), when staged.  Callers: ``tests/golden/make_golden.py``, the
validation tests and ``bench.py --impl reference`` / its ``cpu_baseline`` leg.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("RCN_REFERENCE_ROOT", "/root/reference")
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/stage_ref.py output (sourceless .pyc)


def reference_root():
    """The source tree in the authoring container, else the byte-compiled copy staged by oracle/stage_ref.py, else None."""
    if os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "raw2bit.py")):
        return REFERENCE_ROOT
    if os.path.isfile(os.path.join(STAGED_ROOT, "models", "raw2bit.pyc")):
        try:
            ver = open(os.path.join(STAGED_ROOT, "PYTHON_VERSION")).read().strip()
        except OSError:
            ver = ""
        if ver == f"{sys.version_info.major}.{sys.version_info.minor}":
            return STAGED_ROOT
    return None


def reference_available() -> bool:
    return reference_root() is not None


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Placeholder(nn.Module):
    """Stand-in for classes from files the reference never committed (never on the hot path)."""

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, *a, **k):
        raise NotImplementedError("placeholder for a module absent from the reference repository")


class DropPath(nn.Module):
    """timm DropPath; identity in eval / at rate 0 (the only mode the oracle uses)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


_installed = False


def install_shims():
    global _installed
    if _installed:
        return
    from . import cai

    _mod("ipdb", set_trace=lambda *a, **k: None)
    _mod("thop", profile=lambda *a, **k: (0, 0), clever_format=lambda v, f=None: v)
    timm = _mod("timm")
    timm.data = _mod("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406),
                     IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
    timm.models = _mod("timm.models")
    timm.models.layers = _mod("timm.models.layers", DropPath=DropPath, to_2tuple=to_2tuple,
                              trunc_normal_=trunc_normal_)

    c = _mod("compressai")
    c.entropy_models = _mod("compressai.entropy_models", EntropyBottleneck=cai.EntropyBottleneck,
                            GaussianConditional=cai.GaussianConditional, EntropyModel=cai.EntropyModel)
    c.ans = _mod("compressai.ans", BufferedRansEncoder=cai.BufferedRansEncoder,
                 RansDecoder=cai.RansDecoder, RansEncoder=cai.RansEncoder)
    c.layers = _mod("compressai.layers", AttentionBlock=cai.AttentionBlock, ResidualBlock=cai.ResidualBlock,
                    ResidualBlockUpsample=cai.ResidualBlockUpsample,
                    ResidualBlockWithStride=cai.ResidualBlockWithStride, conv3x3=cai.conv3x3,
                    subpel_conv3x3=cai.subpel_conv3x3, GDN=cai.GDN, MaskedConv2d=cai.MaskedConv2d)
    c.models = _mod("compressai.models", CompressionModel=cai.CompressionModel)
    c.models.google = _mod("compressai.models.google", FactorizedPrior=_Placeholder,
                           ScaleHyperprior=_Placeholder, MeanScaleHyperprior=_Placeholder)
    c.models.utils = _mod("compressai.models.utils", conv=cai.conv, deconv=cai.deconv)
    c.datasets = _mod("compressai.datasets", ImageFolder=object, Vimeo90kDataset=object)
    c.zoo = _mod("compressai.zoo", models={})

    # modules missing from the reference repository itself (namespace package ``models``)
    _mod("models.cbam", CBAM=_Placeholder)
    _mod("models.AWISP_utils", DWT=_Placeholder, IWT=_Placeholder)
    _mod("models.AWISP_modules", **{n: _Placeholder for n in (
        "shortcutblock", "GCIWTResUp", "GCWTResDown", "GCRDB", "ContextBlock2d", "SE_net",
        "PSPModule", "last_upsample")})
    _installed = True


def import_reference():
    """Returns a namespace with the reference modules: .networks .LiteISP .groupmix .tcm .raw2bit"""
    root = reference_root()
    if root is None:
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT} and not staged under {STAGED_ROOT}")
    install_shims()
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib

    ns = types.SimpleNamespace()
    for name in ("networks", "LiteISP", "groupmix", "tcm", "raw2bit"):
        setattr(ns, name, importlib.import_module(f"models.{name}"))
    return ns
