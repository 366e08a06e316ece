"""GPU parity of the 3x3x3 convolution kernels (SIMT fp32 parity path and the tcgen05/TMA bf16 path), through the
C-ABI building blocks, against torch.nn.functional.conv3d evaluated in fp32 on the CPU (the reference's own op).
Tolerances: fp32 path 1e-4 relative (pure fp32 FMA, different summation order); bf16 tensor-core path: operands are
rounded to bf16 first (so the comparison isolates the kernel), fp32 accumulation, bf16 output rounding -> 1e-2."""
import pytest
import torch
import torch.nn.functional as F

from util import rel_err

pytestmark = pytest.mark.gpu


def _case(N, D, H, W, cin, cout, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((N, cin, D, H, W), generator=g)
    w = torch.randn((cout, cin, 3, 3, 3), generator=g) * (2.0 / (27 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    return x, w, b


def _ndhwc(t, dtype):
    return t.permute(0, 2, 3, 4, 1).contiguous().to(dtype).cuda()


SHAPES = [
    # N, D, H, W, cin, cout, stride
    (2, 8, 16, 16, 32, 32, (1, 1, 1)),
    (1, 4, 16, 24, 64, 32, (1, 1, 1)),        # W not a multiple of the 8-wide box
    (2, 8, 8, 8, 32, 64, (2, 2, 2)),
    (2, 4, 8, 8, 64, 128, (1, 2, 2)),
    (2, 4, 4, 4, 320, 320, (1, 1, 1)),        # 64 voxels/sample: the box spans the batch axis; N split in 2 x 160
    (2, 6, 10, 12, 128, 64, (1, 1, 1)),       # ragged everywhere
    (1, 8, 16, 16, 1, 32, (1, 1, 1)),         # first layer: SIMT path even in bf16
    (2, 8, 16, 16, 32, 16, (1, 1, 1)),        # small-channel decoder block of the test geometries
    (2, 4, 8, 8, 16, 32, (2, 2, 2)),
    (2, 8, 16, 16, 8, 8, (1, 1, 1)),
    # halo-reuse kernel (thin layers, H >= 16, W >= 8): ragged d / h, resident and streamed weights
    (2, 20, 32, 40, 32, 64, (1, 1, 1)),
    (1, 9, 24, 16, 64, 64, (1, 1, 1)),
    (2, 16, 16, 8, 32, 128, (1, 1, 1)),
    (1, 12, 48, 24, 64, 32, (1, 1, 1)),
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mode", ["fp32", "bf16_tc", "bf16_tc_nohalo", "bf16_simt"])
def test_conv_forward_and_stats(shape, mode):
    from b200unet import ops
    N, D, H, W, cin, cout, stride = shape
    x, w, b = _case(N, D, H, W, cin, cout, 1)
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    ops.set_option("tensor_cores", 0 if mode == "bf16_simt" else 1)
    ops.set_option("tc_halo", 0 if mode == "bf16_tc_nohalo" else 1)
    try:
        xr = x.to(dt).float()
        wr = w.to(dt).float() if mode.startswith("bf16_tc") and cin % 32 == 0 else w
        ref = F.conv3d(xr, wr, b, stride=stride, padding=1)
        z, stats = ops.conv3d_fwd(_ndhwc(x, dt), w.cuda(), b.cuda(), stride)
        torch.cuda.synchronize()
        got = z.float().cpu().permute(0, 4, 1, 2, 3)
        tol = 1e-4 if mode == "fp32" else 1e-2
        assert got.shape == ref.shape
        assert rel_err(got, ref) < tol, rel_err(got, ref)
        mean = ref.mean(dim=(2, 3, 4))
        rstd = 1.0 / torch.sqrt(ref.var(dim=(2, 3, 4), unbiased=False) + 1e-5)
        assert rel_err(stats[..., 0], mean) < max(tol, 2e-3)
        assert rel_err(stats[..., 1], rstd) < max(tol, 2e-3)
    finally:
        ops.set_option("tensor_cores", 1)
        ops.set_option("tc_halo", 1)


@pytest.mark.parametrize("shape", SHAPES[:6] + SHAPES[7:])  # all but the 1-channel first layer (no dgrad there)
@pytest.mark.parametrize("mode", ["fp32", "bf16_tc"])
def test_conv_backward(shape, mode):
    from b200unet import ops
    N, D, H, W, cin, cout, stride = shape
    x, w, b = _case(N, D, H, W, cin, cout, 2)
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    xr = x.to(dt).float().requires_grad_()
    wr = w.clone().requires_grad_()
    br = b.clone().requires_grad_()
    out = F.conv3d(xr, wr, br, stride=stride, padding=1)
    g = torch.Generator().manual_seed(3)
    dz = torch.randn(out.shape, generator=g)
    dzr = dz.to(dt).float()
    out.backward(dzr)
    dx, dw, db = ops.conv3d_bwd(_ndhwc(x, dt), _ndhwc(dz, dt), w.cuda(), stride)
    torch.cuda.synchronize()
    tol = 1e-4 if mode == "fp32" else 1e-2
    assert rel_err(dx.float().cpu().permute(0, 4, 1, 2, 3), xr.grad) < tol
    assert rel_err(dw, wr.grad) < tol
    assert rel_err(db, br.grad) < tol
    # accumulate flag (skip-connection gradients)
    base = torch.randn(xr.shape, generator=g)
    acc = _ndhwc(base, dt).clone()
    ops.conv3d_bwd(_ndhwc(x, dt), _ndhwc(dz, dt), w.cuda(), stride, accumulate_into=acc)
    assert rel_err(acc.float().cpu().permute(0, 4, 1, 2, 3), base.to(dt).float() + xr.grad) < max(tol, 2e-2 if dt == torch.bfloat16 else tol)


def test_conv_tc_sass_is_blackwell_native():
    """the shipped library contains tcgen05 / TMA SASS (UTCHMMA, UTMALDG, LDTM) -- not a legacy mma.sync path"""
    import shutil
    import subprocess
    from b200unet import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass


# Larger volumes: every CTA walks several items, so the TMEM accumulator ring (16 slots) and the slab ring wrap around,
# which the small shapes above never reach.  The per-tap kernel / the plain wgrad plan of the same library is the checker:
# both sides accumulate the same bf16 products in fp32 (different order) and round the result to bf16 once.
BIG = [
    # N, D, H, W, cin, cout
    (2, 40, 64, 64, 32, 32),
    (1, 24, 64, 64, 64, 64),      # weights do not fit: output channels split over two CTA classes
    (1, 24, 48, 64, 64, 32),
    (1, 20, 32, 64, 32, 64),
    (1, 12, 32, 32, 64, 128),     # four CTA classes
]


@pytest.mark.parametrize("shape", BIG)
def test_halo_kernel_variants_agree_on_large_volumes(shape):
    from b200unet import ops
    N, D, H, W, cin, cout = shape
    x, w, b = _case(N, D, H, W, cin, cout, 5)
    xd, wd_, bd = _ndhwc(x, torch.bfloat16), w.cuda(), b.cuda()
    try:
        z_merged, _ = ops.conv3d_fwd(xd, wd_, bd, (1, 1, 1))
        ops.set_option("halo_merge", 0)
        z_plain, _ = ops.conv3d_fwd(xd, wd_, bd, (1, 1, 1))
        ops.set_option("tc_halo", 0)
        z_tap, _ = ops.conv3d_fwd(xd, wd_, bd, (1, 1, 1))
        torch.cuda.synchronize()
    finally:
        ops.set_option("halo_merge", 1)
        ops.set_option("tc_halo", 1)
    assert rel_err(z_merged.float(), z_tap.float()) < 1e-2
    assert rel_err(z_plain.float(), z_tap.float()) < 1e-2
    # same products, fp32 accumulation: all but a few outputs round to the same bf16 value
    same = (z_merged == z_tap).float().mean().item()
    assert same > 0.9, same


@pytest.mark.parametrize("shape", BIG[:4])
def test_wgrad_variants_agree_on_large_volumes(shape):
    """halo-reuse wgrad kernel == d-merged per-tap plan == plain per-tap plan (fp32 sums of the same bf16 products)"""
    from b200unet import ops
    N, D, H, W, cin, cout = shape
    x, w, b = _case(N, D, H, W, cin, cout, 6)
    g = torch.Generator().manual_seed(7)
    dz = torch.randn((N, cout, D, H, W), generator=g)
    xd, dzd, wd_ = _ndhwc(x, torch.bfloat16), _ndhwc(dz, torch.bfloat16), w.cuda()
    try:
        _, dw_halo, _ = ops.conv3d_bwd(xd, dzd, wd_, (1, 1, 1))
        ops.set_option("wgrad_halo", 0)
        _, dw1, _ = ops.conv3d_bwd(xd, dzd, wd_, (1, 1, 1))
        ops.set_option("wgrad_dmerge", 0)
        _, dw0, _ = ops.conv3d_bwd(xd, dzd, wd_, (1, 1, 1))
        torch.cuda.synchronize()
    finally:
        ops.set_option("wgrad_dmerge", 1)
        ops.set_option("wgrad_halo", 1)
    assert rel_err(dw1, dw0) < 1e-4, rel_err(dw1, dw0)
    assert rel_err(dw_halo, dw0) < 1e-4, rel_err(dw_halo, dw0)


@pytest.mark.parametrize("shape", [(1, 16, 48, 64, 32, 64, (2, 2, 2)), (2, 8, 32, 32, 64, 128, (1, 2, 2)), (1, 9, 30, 36, 32, 64, (2, 2, 2))])
@pytest.mark.parametrize("accumulate", [False, True])
def test_strided_dgrad_merged_classes_agree_with_class_tiles(shape, accumulate):
    """merged-class strided dgrad (all parity classes of a coarse tile in one accumulator set) == one tile class per parity
    class == one launch per class; many tiles per CTA, odd extents"""
    from b200unet import ops
    N, D, H, W, cin, cout, stride = shape
    x, w, b = _case(N, D, H, W, cin, cout, 8)
    od, oh, ow = [(s - 1) // st + 1 for s, st in zip((D, H, W), stride)]
    g = torch.Generator().manual_seed(9)
    dz = torch.randn((N, cout, od, oh, ow), generator=g)
    base = torch.randn((N, cin, D, H, W), generator=g)
    xd, dzd, wd_ = _ndhwc(x, torch.bfloat16), _ndhwc(dz, torch.bfloat16), w.cuda()

    def run():
        acc = _ndhwc(base, torch.bfloat16).clone() if accumulate else None
        dx, _, _ = ops.conv3d_bwd(xd, dzd, wd_, stride, accumulate_into=acc)
        return (acc if accumulate else dx).float()
    try:
        a = run()
        ops.set_option("dgrad_mes", 0)
        b1 = run()
        ops.set_option("dgrad_one_launch", 0)
        c = run()
        torch.cuda.synchronize()
    finally:
        ops.set_option("dgrad_mes", 1)
        ops.set_option("dgrad_one_launch", 1)
    assert rel_err(b1, c) < 1e-2
    assert rel_err(a, c) < 1e-2
    assert (a == c).float().mean().item() > 0.9
