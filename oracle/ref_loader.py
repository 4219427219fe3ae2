"""TEST / BENCHMARK INFRASTRUCTURE - the UNMODIFIED reference, made to travel.

``vendor()`` (run by ``__graft_entry__.build()`` in the authoring container, where /root/reference exists) copies the
few stage-1 files the per-ray hot path lives in - byte for byte, no edits - into ``oracle/_ref/S1/`` (git-ignored, but
shipped to the GPU box with the snapshot, like the built ``.so``).  ``load_s1()`` imports them with inert stubs for the
packages the reference imports but the image lacks (gin, pytorch_lightning, piqa, imageio: decorators / base classes /
IO, never hot-path arithmetic - SURVEY.md 8c).  Only ``bench.py --impl reference`` (the reference arm) and tests use this.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/1st_State-Conditional_Scene"
REF_DST = os.path.join(HERE, "_ref", "S1")
FILES = ["src/model/mipnerf360/helper.py", "src/model/mipnerf360/model.py", "src/model/interface.py", "utils/store_image.py"]


def vendor(force: bool = False) -> bool:
    """Copy the reference files (unmodified) when the reference tree is present.  Returns True if oracle/_ref is usable."""
    if os.path.isdir(REF_SRC):
        for rel in FILES:
            src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
            if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
    return available()


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DST, rel)) for rel in FILES)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    import torch.nn as nn

    def configurable(*a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return lambda f: f

    if "gin" not in sys.modules:
        _mod("gin", configurable=configurable, query_parameter=lambda name: 1)
    if "pytorch_lightning" not in sys.modules:
        _mod("pytorch_lightning", LightningModule=nn.Module).__path__ = []
    if "piqa" not in sys.modules:
        _mod("piqa").__path__ = []
        _mod("piqa.lpips", LPIPS=object)
        _mod("piqa.ssim", SSIM=object)
    if "imageio" not in sys.modules:
        _mod("imageio")


def load_s1():
    """-> (helper module, model module) of the unmodified stage-1 reference (oracle/_ref/S1)."""
    if not available():
        raise RuntimeError("oracle/_ref/S1 is missing: run __graft_entry__.build() where /root/reference exists")
    _install_stubs()
    for k in list(sys.modules):
        if k.split(".")[0] in ("src", "utils"):
            del sys.modules[k]
    sys.path.insert(0, REF_DST)
    try:
        helper = importlib.import_module("src.model.mipnerf360.helper")
        model = importlib.import_module("src.model.mipnerf360.model")
    finally:
        sys.path.remove(REF_DST)
    return helper, model


def build_reference_net(model_mod, nerf_netwidth=None, **kw):
    """``MipNeRF360`` of the reference with the gin binding ``NeRFMLP.netwidth`` applied through the constructor default."""
    saved = model_mod.NeRFMLP.__init__.__defaults__
    if nerf_netwidth is not None:
        model_mod.NeRFMLP.__init__.__defaults__ = (saved[0], nerf_netwidth)
    try:
        return model_mod.MipNeRF360("/nonexistent", **kw)
    finally:
        model_mod.NeRFMLP.__init__.__defaults__ = saved
