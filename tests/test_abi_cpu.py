"""CPU checks of the C-ABI boundary: the shared library builds/loads, exports every symbol that
include/hosnerf_b200.h declares, the ctypes table covers the header, and the product path fails
loudly (no silent CPU fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "hosnerf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hos_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from hosnerf_b200 import _lib
    return _lib.load()


def test_header_symbols_exported(lib):
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/hosnerf_b200.h but not exported"


def test_ctypes_table_matches_header(lib):
    from hosnerf_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_symbols()


def test_version_and_error_string(lib):
    assert lib.hos_version() >= 100
    assert isinstance(lib.hos_last_error(), bytes)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(lib):
    from hosnerf_b200 import MipNeRF360, ops, synth
    assert lib.hos_device_check(0) != 0          # no device here: must report, not crash
    net = MipNeRF360("/nonexistent", num_levels=2, nerf_netwidth=256)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(synth.make_bkg_batch(4), 1.0, False, False, 0.1, 1e6)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.composite_mip360(torch.zeros(2, 4), torch.zeros(2, 5), torch.zeros(2, 3))


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "hosnerf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_state_dict_contract():
    """Checkpoint key names are an external contract."""
    from hosnerf_b200 import MipNeRF360, Network, default_cfg
    k = set(MipNeRF360("/nonexistent").state_dict())
    for name in ("mlps.0.pos_basis_t", "mlps.0.bkgd_stateembeds.0", "mlps.0.pts_linear.3.weight",
                 "mlps.0.density_layer.bias", "mlps.2.pts_linear.7.weight", "mlps.2.bottleneck_layer.weight",
                 "mlps.2.views_linear.0.weight", "mlps.2.rgb_layer.bias"):
        assert name in k, name
    net = MipNeRF360("/nonexistent")
    assert net.mlps[2].pts_linear[5].weight.shape == (1024, 1024 + 568)
    assert net.mlps[2].views_linear[0].weight.shape == (128, 283)
    assert sum(p.numel() for p in net.parameters()) == 9498438
    h = Network(default_cfg())
    k = set(h.state_dict())
    for name in ("cnl_mlp.pts_linears.14.weight", "cnl_mlp.output_linear.0.bias", "non_rigid_mlp.block_mlps.12.weight",
                 "non_rigid_forward_mlp.block_mlps.0.weight", "human_stateembeds.0", "mweight_vol_decoder.const_embedding",
                 "mweight_vol_decoder.decoder.block_conv.8.weight", "pose_decoder.block_mlps_dstR.2.bias"):
        assert name in k, name
    assert h.cnl_mlp.pts_linears[10].weight.shape == (256, 383)
    assert h.non_rigid_mlp.block_mlps[8].weight.shape == (128, 164)
    assert sum(p.numel() for p in h.parameters()) == 64673787


def test_header_is_plain_c_and_links(tmp_path):
    """include/hosnerf_b200.h is the drop-in boundary: it must compile as C11 (no C++ or torch types in the
    signatures) and a C program must link against the shared library."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "hosnerf_b200.h"\n'
                   'int main(void) { hos_bkg_config cfg; (void)cfg; printf("%d\\n", hos_version()); return 0; }\n')
    inc = os.path.join(root, "include")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
    exe = tmp_path / "abi"
    libdir = os.path.join(root, "hosnerf_b200")
    subprocess.check_call(["gcc", "-std=c11", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-l:libhosnerf_b200.so",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and int(out.stdout.strip()) >= 1


def test_render_bkg_workspace_query_needs_no_gpu(lib):
    """hos_render_bkg_workspace validates the level configuration and sizes the caller's workspace on the host:
    errors come back as status + hos_last_error, sizes follow the documented sub-buffers."""
    import ctypes
    from hosnerf_b200 import _lib
    cfg = _lib.BkgConfig()
    n = ctypes.c_size_t(0)
    assert lib.hos_render_bkg_workspace(ctypes.byref(cfg), 128, ctypes.byref(n)) != 0
    assert b"n_levels" in lib.hos_last_error()
    basis = (ctypes.c_float * 63)()
    cfg.n_levels, cfg.deg_view, cfg.basis_host = 2, 4, basis
    dummy = ctypes.c_void_p(0x1000)            # never dereferenced by the size query
    for i, s in enumerate((64, 32)):
        cfg.levels[i].mlp, cfg.levels[i].n_samples, cfg.levels[i].u_base = dummy, s, dummy
    assert lib.hos_render_bkg_workspace(ctypes.byref(cfg), 128, ctypes.byref(n)) != 0      # final level without view term
    cfg.levels[1].view_W, cfg.levels[1].view_b, cfg.levels[1].view_dim = dummy, dummy, 128
    cfg.levels[0].dilate = 1
    assert lib.hos_render_bkg_workspace(ctypes.byref(cfg), 128, ctypes.byref(n)) != 0      # level 0 cannot dilate
    cfg.levels[0].dilate, cfg.levels[1].dilate = 0, 1
    sizes = []
    for rays in (0, 1, 128, 4096):
        assert lib.hos_render_bkg_workspace(ctypes.byref(cfg), rays, ctypes.byref(n)) == 0
        sizes.append(n.value)
    assert sizes == sorted(sizes) and sizes[0] >= 0

    def expect(N, smax=64, slast=32, vdim=128, de=27):
        al = lambda f: (f * 4 + 255) & ~255
        return (al(N * 2) + al(N) + 2 * (al(N * (smax + 1)) + al(N * smax)) + al(N * (smax + 1)) + al(N * smax)
                + al(N * slast * 3) + al(N * de) + al(N * vdim))
    assert sizes[3] == expect(4096) and sizes[2] == expect(128)
    cfg.levels[1].n_samples = 400                                                          # 3 S + 1 > 1024 knots
    assert lib.hos_render_bkg_workspace(ctypes.byref(cfg), 128, ctypes.byref(n)) != 0


def test_host_entry_points_refuse_cpu_inputs():
    """The reference-facing modules have no CPU path: CPU tensors / a module left on the CPU raise instead of falling back."""
    from hosnerf_b200 import LitMipNeRF360, synth
    lit = LitMipNeRF360("/nonexistent", num_levels=2, num_prop_samples=8, num_nerf_samples=8, nerf_netwidth=64)
    b = synth.make_bkg_batch(4, seed=0)
    with pytest.raises(RuntimeError):
        lit.render_rays(b, 0)
    with pytest.raises(RuntimeError):
        list(lit.render_rays_stream(iter([b])))
    with pytest.raises(RuntimeError):
        lit.model.render_fused(b, 1.0, False, 0.1, 1e6)
    b["target"] = torch.zeros(b["rays_o"].shape[0], 3)
    with pytest.raises(RuntimeError, match="CUDA"):       # the training step has no CPU path either
        lit.training_step(b, 0)
