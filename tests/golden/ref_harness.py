"""Import the UNMODIFIED reference (/root/reference) in the authoring container.

Only ``make_golden.py`` uses this file; nothing under tests/ that runs on the
GPU box imports it (``/root/reference`` does not exist there).  The reference
needs a handful of packages that are not installed (gin, pytorch_lightning,
piqa, imageio, skimage, termcolor, torch_efficient_distloss) and the stdlib
``imp`` module that Python 3.12 removed.  None of them touches hot-path
arithmetic, so they are replaced with inert in-memory stubs.
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.machinery
import importlib.util
import os
import sys
import types

REF_ROOT = "/root/reference"
S1 = os.path.join(REF_ROOT, "1st_State-Conditional_Scene")
S2 = os.path.join(REF_ROOT, "2nd_State_Conditional_Human-Object")
S3 = os.path.join(REF_ROOT, "3rd_Complete_HOSNeRF")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    import torch.nn as nn

    def configurable(*a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return lambda f: f

    _mod("gin", configurable=configurable, query_parameter=lambda name: 1,
         parse_config_files_and_bindings=lambda *a, **k: None)
    pl = _mod("pytorch_lightning", LightningModule=nn.Module)
    pl.__path__ = []
    _mod("piqa")
    _mod("piqa.lpips", LPIPS=object)
    _mod("piqa.ssim", SSIM=object)
    sys.modules["piqa"].__path__ = []
    _mod("imageio")
    sk = _mod("skimage")
    sk.__path__ = []
    _mod("skimage.metrics", structural_similarity=lambda *a, **k: 0.0)
    _mod("termcolor", colored=lambda s, *a, **k: s)
    _mod("torch_efficient_distloss", eff_distloss=None)
    _mod("tqdm", tqdm=lambda x, *a, **k: x) if "tqdm" not in sys.modules else None

    def load_source(name, path):
        loader = importlib.machinery.SourceFileLoader(name, path)
        spec = importlib.util.spec_from_loader(name, loader)
        module = importlib.util.module_from_spec(spec)
        sys.modules[name] = module
        loader.exec_module(module)
        return module

    _mod("imp", load_source=load_source)


_STAGE_PKGS = ("src", "utils", "core", "third_parties", "configs")


@contextlib.contextmanager
def stage(stage_dir):
    """Make ``stage_dir`` the import root *and* CWD (the reference resolves
    cfg.*.module paths relative to CWD, S3/core/nets/human_nerf/component_factory.py:12-40).
    The three stages share top-level package names, so cached modules are purged
    on entry and exit."""
    def purge():
        for k in list(sys.modules):
            if k.split(".")[0] in _STAGE_PKGS:
                del sys.modules[k]
    purge()
    old_cwd = os.getcwd()
    sys.path.insert(0, stage_dir)
    os.chdir(stage_dir)
    try:
        yield
    finally:
        os.chdir(old_cwd)
        sys.path.remove(stage_dir)
        purge()


def human_cfg(stage_dir, basedir="/nonexistent", **overrides):
    """yacs cfg exactly as S2/S3 run.py builds it: defaults -> configs/default.yaml."""
    from third_parties.yacs import CfgNode as CN
    cfg = CN()
    # run.py defaults (S3/run.py:35-52)
    cfg.resume = False
    cfg.eval_iter = 10000000
    cfg.render_folder_name = ""
    cfg.ignore_non_rigid_motions = False
    cfg.render_skip = 1
    cfg.render_frames = 100
    cfg.num_workers = 4
    cfg.merge_from_file(os.path.join(stage_dir, "configs", "default.yaml"))
    adv = os.path.join(stage_dir, "configs", "human_nerf", "wild", "monocular", "adventure.yaml")
    if os.path.exists(adv):
        cfg.merge_from_file(adv)
    cfg.basedir = basedir
    for k, v in overrides.items():
        node = cfg
        parts = k.split(".")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = v
    return cfg
