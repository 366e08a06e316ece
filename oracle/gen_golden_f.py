"""Golden fixtures of the rows around the step: the patch pipeline and the sliding-window prediction, written from
the ORACLE restatements (oracle/augment.py -- scipy.ndimage --, oracle/sliding_window.py) on seeded inputs.  nnunet /
batchgenerators are absent, so these pin the oracles (and, through tests/test_gpu_*, the CUDA path) to today's behaviour, not to
the un-vendored packages ("parity unpinned" at that boundary).  Test infrastructure.

  python oracle/gen_golden_f.py   ->  tests/golden/augment_tiny.npz, tests/golden/sliding_window_tiny.npz
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lifelong-nnunet_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

PATCH, STRIDES = (16, 32, 32), [(1, 1, 1), (2, 2, 2), (4, 4, 4)]
AUG_PARAMS = dict(p_rot=1.0, p_scale=1.0, p_noise=1.0, p_blur=1.0, p_blur_per_channel=1.0, p_brightness=1.0, p_contrast=1.0,
                  p_lowres=1.0, p_lowres_per_channel=1.0, p_gamma_inverted=1.0, p_gamma=1.0)


def augment_cases(seed=21):
    rs = np.random.RandomState(seed)
    out = []
    for i, sh in enumerate(((30, 44, 40), (22, 50, 46))):
        d = rs.randn(3, *sh).astype(np.float32)
        seg = np.zeros(sh, np.float32)
        seg[sh[0] // 4: sh[0] // 2, sh[1] // 3: sh[1] // 2, sh[2] // 4: sh[2] // 2] = 1
        seg[sh[0] // 2: sh[0] // 2 + 4, 5:20, 8:30] = 2
        seg[:2] = -1
        d[-1] = seg
        out.append({"key": "g%d" % i, "data": d})
    return out


def augment_values():
    from b200unet import augment            # host-only use: the plan (random parameters) comes from the product's draw_plan
    from oracle import augment as oaug
    cases = augment_cases()
    pipe = augment.GPUPatchPipeline(cases, PATCH, 2, STRIDES, params=AUG_PARAMS, seed=8, plan_only=True)
    plan = pipe.draw_plan()
    data, targets, margin = oaug.apply_plan([c["data"] for c in cases], plan, PATCH, pipe.gen_patch, STRIDES)
    vals = {"plan_json": np.frombuffer(json.dumps(plan).encode(), dtype=np.uint8), "data": data, "margin": margin.astype(np.float32),
            "gen_patch": np.array(pipe.gen_patch)}
    for k, t in enumerate(targets):
        vals["target%d" % k] = t.astype(np.int8)
    return vals


def sliding_values():
    from b200unet.configs import CONFIGS
    from oracle import sliding_window, step
    geom = CONFIGS["tiny"]
    net = step.build_network(geom.in_channels, geom.base_features, geom.num_classes, [list(k) for k in geom.pool])
    net.eval()
    x = torch.randn((1, 18, 40, 36), generator=torch.Generator().manual_seed(17))
    seg, prob = sliding_window.predict_3D(net, x, geom.patch, True, (0, 1, 2), 0.5, True)
    return {"x": x.numpy(), "prob": prob.numpy().astype(np.float32), "seg": seg.numpy().astype(np.int8)}


if __name__ == "__main__":
    torch.set_num_threads(4)
    gold = os.path.join(ROOT, "tests", "golden")
    np.savez_compressed(os.path.join(gold, "augment_tiny.npz"), **augment_values())
    np.savez_compressed(os.path.join(gold, "sliding_window_tiny.npz"), **sliding_values())
    print({f: os.path.getsize(os.path.join(gold, f)) for f in ("augment_tiny.npz", "sliding_window_tiny.npz")})
