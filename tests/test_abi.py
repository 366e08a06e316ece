"""CPU: the C-ABI library loads and exports every symbol include/b2unet.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import util  # noqa: F401  (sys.path)


def _declared():
    hdr = open(os.path.join(util.ROOT, "include", "b2unet.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from b200unet import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with `python -c 'import __graft_entry__ as g; g.build()'`"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert set(names) == set(_lib.SIGNATURES.keys()), set(names) ^ set(_lib.SIGNATURES.keys())


def test_plan_metadata_matches_reference_module_tree():
    """plan parameter table == state_dict keys/shapes of the oracle's Generic_UNet (host-only calls)"""
    from b200unet.configs import CONFIGS
    from b200unet.generic_UNet import Generic_UNet, _Plan
    from b200unet import _lib
    import torch
    for name in ("tiny", "tiny3", "cfg1"):
        geom = CONFIGS[name]
        onet = util.oracle_net(geom)
        net = Generic_UNet(geom.in_channels, geom.base_features, geom.num_classes, geom.num_pool,
                           pool_op_kernel_sizes=[list(k) for k in geom.pool], max_num_features=geom.max_features)
        osd, sd = onet.state_dict(), net.state_dict()
        assert list(osd.keys()) == list(sd.keys())
        assert all(osd[k].shape == sd[k].shape for k in sd)
        lib = _lib.load()
        g = _lib.Geometry()
        g.batch, g.in_channels, g.num_classes = geom.batch, geom.in_channels, geom.num_classes
        g.base_features, g.max_features, g.num_pool = geom.base_features, geom.max_features, geom.num_pool
        for i in range(3):
            g.patch[i] = geom.patch[i]
        for l, k in enumerate(geom.pool):
            for i in range(3):
                g.pool[l][i] = k[i]
        g.act_dtype, g.lrelu_slope, g.norm_eps = 0, 1e-2, 1e-5
        h = ctypes.c_void_p()
        _lib.check(lib.b2_unet_plan_create(ctypes.byref(g), ctypes.byref(h)))
        info = _lib.ParamInfo()
        names = {}
        for i in range(lib.b2_unet_num_params(h)):
            _lib.check(lib.b2_unet_param_info(h, i, ctypes.byref(info)))
            names[info.name.decode()] = tuple(info.shape[j] for j in range(info.ndim))
        assert set(names) == set(dict(onet.named_parameters()).keys())
        for k, s in names.items():
            assert tuple(osd[k].shape) == s, k
        assert lib.b2_unet_workspace_bytes(h) > 0
        assert lib.b2_unet_num_convs(h) == 2 * (geom.num_pool + 1) + 3 * geom.num_pool
        lib.b2_unet_plan_destroy(h)


def test_invalid_geometry_is_rejected():
    from b200unet import _lib
    lib = _lib.load()
    g = _lib.Geometry()
    h = ctypes.c_void_p()
    assert lib.b2_unet_plan_create(ctypes.byref(g), ctypes.byref(h)) != 0
    assert b"invalid argument" in lib.b2_last_error()


def test_every_kernel_is_pdl_safe():
    """b2_set_option("pdl", 1) launches every kernel with programmatic stream serialization: that is only correct if every
    __global__ function starts with pdl_grid_sync() (griddepcontrol.wait) -- checked on the sources."""
    import glob
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lifelong-nnunet_b200", "csrc")
    missing = []
    for path in sorted(glob.glob(os.path.join(root, "*.cu"))):
        src = open(path).read()
        for m in re.finditer(r"__global__", src):
            i, depth = m.end(), 0
            while True:
                ch = src[i]
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                elif ch == "{" and depth == 0:
                    break
                elif ch == ";" and depth == 0:
                    i = -1
                    break
                i += 1
            if i < 0:
                continue
            if not src[i + 1:i + 60].lstrip().startswith("pdl_grid_sync();"):
                missing.append("%s: %s" % (os.path.basename(path), src[m.start():m.start() + 120].split("(")[0].replace("\n", " ")))
    assert not missing, missing


def test_fastdiv_exact(tmp_path):
    """the magic-number division of the kernels' tile decode (tc_common.cuh: make_fastdiv / fd_div) is exact for n < 2^31:
    a host mirror of the device expression is checked on ~27 M (n, d) pairs (built with nvcc, no GPU needed)"""
    import os
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        import pytest
        pytest.skip("nvcc not available")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host", "fastdiv_host.cu")
    exe = str(tmp_path / "fastdiv_host")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, src])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr


def test_gradient_buckets_follow_the_backward_order_of_the_cfg2_plan():
    """data-parallel overlap: the buckets b2_unet_backward_buckets signals are suffixes of the plan's parameter
    order cut at layer starts -- contiguous, covering every parameter exactly once, each >= the target size except the last,
    first_param strictly descending (the C entry point checks the same), the full-resolution encoder layers (finished last) in
    the last bucket"""
    from b200unet.configs import CONFIGS
    from b200unet.fused_step import bucket_bounds
    from b200unet import _lib
    geom = CONFIGS["cfg2"]
    lib = _lib.load()
    g = _lib.Geometry()
    g.batch, g.in_channels, g.num_classes = geom.batch, geom.in_channels, geom.num_classes
    g.base_features, g.max_features, g.num_pool = geom.base_features, geom.max_features, geom.num_pool
    for i in range(3):
        g.patch[i] = geom.patch[i]
    for l, k in enumerate(geom.pool):
        for i in range(3):
            g.pool[l][i] = k[i]
    g.act_dtype, g.lrelu_slope, g.norm_eps = 1, 1e-2, 1e-5
    h = ctypes.c_void_p()
    _lib.check(lib.b2_unet_plan_create(ctypes.byref(g), ctypes.byref(h)))
    info, names, numels = _lib.ParamInfo(), [], []
    for i in range(lib.b2_unet_num_params(h)):
        _lib.check(lib.b2_unet_param_info(h, i, ctypes.byref(info)))
        names.append(info.name.decode())
        numels.append(int(info.numel))
    lib.b2_unet_plan_destroy(h)
    assert sum(numels) == 30_787_840 or sum(numels) > 30e6          # 30.8 M parameters: the 123 MB arena of DESIGN section 5
    for target_mb in (8, 48, 10 ** 6):
        bounds = bucket_bounds(names, numels, target_mb << 20)
        assert bounds[0][1] == len(names) and bounds[-1][0] == 0
        assert all(bounds[k][0] == bounds[k + 1][1] for k in range(len(bounds) - 1))
        assert all(bounds[k][0] > bounds[k + 1][0] for k in range(len(bounds) - 1))
        sizes = [4 * sum(numels[lo:hi]) for lo, hi in bounds]
        assert sum(sizes) == 4 * sum(numels) and all(s >= (target_mb << 20) for s in sizes[:-1])
        for lo, _ in bounds:
            assert names[lo].endswith("conv.weight") or names[lo].startswith(("tu.", "seg_outputs."))
    b48 = bucket_bounds(names, numels, 48 << 20)
    assert len(b48) == 3 and "conv_blocks_context.0.blocks.0.conv.weight" in names[b48[-1][0]:b48[-1][1]]
    assert len(bucket_bounds(names, numels, 10 ** 12)) == 1
