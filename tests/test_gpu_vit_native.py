"""GPU: the hand-written ViT (csrc/vit.cu: tcgen05 GEMMs, shared-memory attention, LayerNorm / GELU kernels; forward and
backward, bf16 operands) against the SAME module evaluated through ATen in fp32 -- the path that tests/test_gpu_vit_unet.py
holds to the oracle (reference vision_transformer.py:218-458) at 1e-3.  Tolerances: bf16 operand rounding (2^-9 relative per
product, fp32 accumulation): outputs within 3e-2 of the max-norm, gradients by cosine similarity > 0.99 and norm ratio
within 5 %."""
import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))


CASES = [
    # B, C, (D, H, W), patch, embed, heads, depth, out (c, d, h, w)
    (2, 8, (16, 32, 32), 16, 128, 2, 2, (16, 2, 4, 4)),        # 4 + 1 tokens (the tiny ViT-U-Net geometry)
    (2, 32, (16, 64, 64), 8, 128, 2, 2, (8, 4, 4, 4)),         # 128 + 1 tokens: several key chunks per warp in attention
    (1, 16, (8, 24, 40), 8, 64, 1, 1, (4, 1, 3, 5)),           # ragged volume (floor semantics of the k = s Conv3d), batch 1
    (2, 32, (24, 96, 96), 8, 128, 2, 1, (8, 3, 4, 4)),         # 432 + 1 tokens (the cfg4 token count): 7 key blocks, ragged last block
]


@pytest.mark.parametrize("lsa", [False, True])
@pytest.mark.parametrize("case", CASES)
def test_native_vit_forward_backward_vs_aten_fp32(case, lsa):
    """lsa: Locality Self-Attention (learnable per-head temperature -- also its gradient --, masked diagonal, bias-free qkv)"""
    from b200unet.vision_transformer import VisionTransformer
    B, Cc, (D, H, W), patch, E, heads, depth, oshape = case
    F_ = oshape[0] * oshape[1] * oshape[2] * oshape[3]
    torch.manual_seed(0)
    vit = VisionTransformer(ViT_2d=False, img_size=[D, H, W], patch_size=(patch, patch), img_depth=[D], in_chans=Cc, num_classes=F_,
                            embed_dim=E, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True, is_LSA=lsa).cuda()
    with torch.no_grad():      # non-trivial values everywhere (the reference leaves pos_embed / biases at zero)
        for n, p in vit.named_parameters():
            if n.endswith('attn.scale'):
                p.mul_(1.0 + 0.3 * torch.randn_like(p))
            elif p.dim() == 1 or 'pos_embed' in n or 'cls' in n:
                p.copy_(0.1 * torch.randn_like(p) + (1.0 if 'norm' in n and n.endswith('weight') else 0.0))
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((B, Cc, D, H, W), generator=g, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last_3d)
    u = torch.randn((B, F_), generator=g, device="cuda")
    # reference: ATen fp32 on the bf16-rounded input
    xr = x.float().requires_grad_()
    out_ref = vit(xr)
    (out_ref * u).sum().backward()
    ref_grads = {n: p.grad.clone() for n, p in vit.named_parameters() if p.grad is not None}
    dx_ref = xr.grad.clone()
    vit.zero_grad(set_to_none=True)
    # native
    assert vit.native_supported(x)
    dskip = torch.zeros_like(x)
    out = vit.forward_native(x.requires_grad_(), dskip, oshape)
    assert rel_err(out, out_ref) < 3e-2, rel_err(out, out_ref)
    (out * u).sum().backward()
    torch.cuda.synchronize()
    bad = []
    for n, p in vit.named_parameters():
        if lsa and '.attn.proj.' in n:          # timm's proj stays registered but unused under LSA (no gradient in either path)
            assert p.grad is None and n not in ref_grads
            continue
        assert p.grad is not None, n
        c = _cos(p.grad, ref_grads[n])
        ratio = float(p.grad.double().norm() / (ref_grads[n].double().norm() + 1e-30))
        if not (c > 0.99 and 0.95 < ratio < 1.05):
            bad.append("%-40s cos %.4f ratio %.3f" % (n, c, ratio))
    assert not bad, "\n".join(bad)
    gd, gh, gw = D // patch, H // patch, W // patch
    covered = dx_ref[:, :, :gd * patch, :gh * patch, :gw * patch]
    got = dskip.float()[:, :, :gd * patch, :gh * patch, :gw * patch]
    assert _cos(got, covered) > 0.99 and rel_err(got, covered) < 5e-2
    # deterministic: a second forward / backward gives bit-identical results
    vit.zero_grad(set_to_none=True)
    g1 = {n: None for n, _ in vit.named_parameters()}
    out2 = vit.forward_native(x, torch.zeros_like(x), oshape)
    assert torch.equal(out2, out)


def test_native_vit_no_library_kernels_in_bf16_mode():
    """bf16 mode launches only kernels of libb2unet for the ViT: the kernel trace of a ViT-U-Net step holds no ATen / cuBLAS /
    SDPA kernel between the encoder and the decoder"""
    from torch.profiler import ProfilerActivity, profile
    from b200unet.vision_transformer import VisionTransformer
    torch.manual_seed(0)
    vit = VisionTransformer(ViT_2d=False, img_size=[16, 32, 32], patch_size=(16, 16), img_depth=[16], in_chans=8, num_classes=512,
                            embed_dim=128, depth=1, num_heads=2, mlp_ratio=4, qkv_bias=True).cuda()
    x = torch.randn((2, 8, 16, 32, 32), device="cuda").bfloat16().contiguous(memory_format=torch.channels_last_3d).requires_grad_()
    dskip = torch.zeros_like(x)
    out = vit.forward_native(x, dskip, (16, 2, 4, 4))       # warm-up (plan creation, attribute setting)
    out.sum().backward()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        out = vit.forward_native(x, dskip, (16, 2, 4, 4))
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.key and "Memset" not in e.key]
    foreign = [n for n in names if "b2::" not in n and "b2_" not in n]
    assert names and not foreign, foreign
