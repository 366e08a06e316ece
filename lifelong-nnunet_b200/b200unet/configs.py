"""Network geometries of the BASELINE.json configs.

nnU-Net planner conventions: widths min(base * 2^d, 320), two 3x3x3 convs per stage, stride-conv pooling,
kernel==stride transposed-conv upsampling, one 1x1x1 head per decoder level.
"""
from dataclasses import dataclass
from typing import List, Tuple


@dataclass(frozen=True)
class UNetGeometry:
    name: str
    patch: Tuple[int, int, int]
    pool: Tuple[Tuple[int, int, int], ...]
    in_channels: int = 1
    num_classes: int = 3
    base_features: int = 32
    max_features: int = 320
    batch: int = 2

    @property
    def num_pool(self):
        return len(self.pool)

    def stage_features(self) -> List[int]:
        f, out = self.base_features, []
        for _ in range(self.num_pool + 1):
            out.append(min(f, self.max_features))
            f = int(round(f * 2))
        return out

    def stage_sizes(self) -> List[Tuple[int, int, int]]:
        d, h, w = self.patch
        out = [(d, h, w)]
        for k in self.pool:
            assert d % k[0] == 0 and h % k[1] == 0 and w % k[2] == 0, "patch must divide by pool strides"
            d, h, w = d // k[0], h // k[1], w // k[2]
            out.append((d, h, w))
        return out

    def fwd_flops_per_patch(self) -> float:
        """Conv-stack FLOPs, 2*V_out*Cin*Cout*taps."""
        feats, sizes = self.stage_features(), self.stage_sizes()
        vol = lambda s: s[0] * s[1] * s[2]
        fl, cin = 0.0, self.in_channels
        for d in range(self.num_pool + 1):
            fl += 2.0 * vol(sizes[d]) * cin * feats[d] * 27
            fl += 2.0 * vol(sizes[d]) * feats[d] * feats[d] * 27
            cin = feats[d]
        for u in range(self.num_pool):
            lvl = self.num_pool - 1 - u
            cdown, cskip = feats[lvl + 1], feats[lvl]
            fl += 2.0 * vol(sizes[lvl]) * cdown * cskip            # transposed conv, one tap per output voxel
            fl += 2.0 * vol(sizes[lvl]) * (2 * cskip) * cskip * 27
            fl += 2.0 * vol(sizes[lvl]) * cskip * cskip * 27
            fl += 2.0 * vol(sizes[lvl]) * cskip * self.num_classes
        return fl


P2 = (2, 2, 2)
P1 = (1, 2, 2)

CONFIGS = {
    # cfg1: Sequential nnUNetTrainer, 2-stage, 32x64x64, CPU-runnable
    "cfg1": UNetGeometry("cfg1", (32, 64, 64), (P2, P2), 1, 3, 32, 320, 2),
    # cfg2 / cfg5: EWC / RW, 5-stage, 64x128x128 hippocampus-shaped
    "cfg2": UNetGeometry("cfg2", (64, 128, 128), (P2, P2, P2, P2, P1), 1, 3, 32, 320, 2),
    "cfg5": UNetGeometry("cfg5", (64, 128, 128), (P2, P2, P2, P2, P1), 1, 3, 32, 320, 2),
    # cfg3: LwF, 64x160x160 prostate-shaped
    "cfg3": UNetGeometry("cfg3", (64, 160, 160), (P2, P2, P2, P2, P1), 1, 2, 32, 320, 2),
    # cfg4 (U-Net part): PLOP, 48x192x192
    "cfg4": UNetGeometry("cfg4", (48, 192, 192), (P2, P2, P2, P1, P1), 1, 2, 32, 320, 2),
    # tiny geometry for parity tests / smoke
    "tiny": UNetGeometry("tiny", (16, 32, 32), (P2, P2), 1, 3, 8, 320, 2),
    # tensor-core capable small geometry (channels multiples of 32; a (1,2,2) pool)
    "tiny32": UNetGeometry("tiny32", (8, 32, 32), (P2, P1), 1, 3, 32, 320, 2),
    "tiny3": UNetGeometry("tiny3", (8, 32, 32), (P2, P2, P1), 2, 3, 8, 24, 2),
}
