"""Seeded synthetic inputs of SURVEY.md section 8(d) (device-agnostic torch code; generated on the CPU generator so
that the oracle and the CUDA path see bit-identical inputs)."""
import torch
import torch.nn.functional as F


def make_batch(geom, batch=None, seed=1234):
    """data ~ N(0,1) fp32 (B,C,D,H,W); target[0] = block-wise random labels (4^3 blocks) so Dice is non-degenerate;
    deep-supervision targets = strided sub-sampling with the cumulative pool strides."""
    B = geom.batch if batch is None else batch
    g = torch.Generator().manual_seed(seed)
    D, H, W = geom.patch
    data = torch.randn((B, geom.in_channels, D, H, W), generator=g, dtype=torch.float32)
    bd, bh, bw = max(D // 4, 1), max(H // 4, 1), max(W // 4, 1)
    coarse = torch.randint(0, geom.num_classes, (B, 1, bd, bh, bw), generator=g).float()
    t0 = F.interpolate(coarse, size=(D, H, W), mode="nearest")
    targets, sd, sh, sw = [t0], 1, 1, 1
    for k in geom.pool[:-1]:
        sd, sh, sw = sd * k[0], sh * k[1], sw * k[2]
        targets.append(t0[..., ::sd, ::sh, ::sw].contiguous())
    return data, targets


def make_ewc_state(named_params, seed=7, with_scores=False):
    """theta* = theta + 0.01 N(0,1); F = 0.01 N(0,1)^2; S = 0.01 U(0,1)  (SURVEY.md 8(d))."""
    g = torch.Generator().manual_seed(seed)
    fisher, params, scores = {}, {}, {}
    for name, p in named_params:
        pc = p.detach().cpu().float()
        params[name] = pc + 0.01 * torch.randn(pc.shape, generator=g)
        fisher[name] = 0.01 * torch.randn(pc.shape, generator=g).pow(2)
        if with_scores:
            scores[name] = 0.01 * torch.rand(pc.shape, generator=g)
    return (fisher, params, scores) if with_scores else (fisher, params)
