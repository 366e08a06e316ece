"""GPU parity of the validation / inference sweep: tiled sliding-window prediction with Gaussian weighting
and test-time mirroring (b200unet/inference.py: plan forward + b2_sliding_accumulate / b2_sliding_finalize) against the CPU
restatement of nnunet's predict_3D (oracle/sliding_window.py), and the per-subject Dice / IoU sweep of the trainer
(reference MultiHead:678-901, 963-1049) against the same counts taken from the oracle network.
Tolerance: class probabilities within 1e-3 absolute (fp32 path); segmentation labels equal wherever the oracle's top-two
probabilities differ by more than 2e-3."""
import numpy as np
import pytest
import torch

from util import cuda_net, oracle_net

pytestmark = pytest.mark.gpu


def _case(shape, seed=5):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((1,) + tuple(shape), generator=g)


@pytest.mark.parametrize("shape,mirror,axes", [((20, 45, 50), True, (0, 1, 2)), ((16, 40, 32), True, (1, 2)),
                                               ((24, 32, 61), False, (0, 1, 2)), ((10, 20, 30), True, (0, 1, 2))])
def test_sliding_window_prediction_matches_oracle(shape, mirror, axes):
    from b200unet import inference
    from b200unet.configs import CONFIGS
    from oracle import sliding_window
    geom = CONFIGS["tiny"]
    onet = oracle_net(geom)
    onet.eval()
    cnet = cuda_net(geom, onet.state_dict())
    cnet.eval()
    x = _case(shape)
    oseg, oprob = sliding_window.predict_3D(onet, x, geom.patch, mirror, axes, 0.5, True)
    cseg, cprob = inference.predict_3D(cnet, x, geom.patch, mirror, axes, 0.5, True)
    assert tuple(cprob.shape) == tuple(oprob.shape) == (geom.num_classes,) + tuple(shape)
    assert float((cprob.cpu() - oprob).abs().max()) < 1e-3
    top2 = oprob.topk(2, 0).values
    decided = (top2[0] - top2[1]) > 2e-3
    assert bool((cseg.cpu()[decided] == oseg[decided]).all())
    assert float(decided.float().mean()) > 0.9
    assert abs(float(cprob.sum(0).mean()) - 1.0) < 1e-5


def test_sliding_window_bf16_network():
    """the production precision: probabilities within 3e-2 of the fp32 oracle"""
    from b200unet import inference
    from b200unet.configs import CONFIGS
    from oracle import sliding_window
    geom = CONFIGS["tiny32"]
    onet = oracle_net(geom)
    cnet = cuda_net(geom, onet.state_dict(), precision="bf16")
    x = _case((12, 48, 40))
    _, oprob = sliding_window.predict_3D(onet, x, geom.patch, True, (0, 1, 2), 0.5, True)
    _, cprob = inference.predict_3D(cnet, x, geom.patch, True, (0, 1, 2), 0.5, True)
    assert float((cprob.cpu() - oprob).abs().max()) < 3e-2


def test_per_subject_validation_sweep_matches_oracle_counts():
    from b200unet import synth
    from b200unet.configs import CONFIGS
    from b200unet.trainers import nnUNetTrainerSequential
    from oracle import cl_losses
    geom = CONFIGS["tiny"]
    onet = oracle_net(geom)
    tr = nnUNetTrainerSequential(geom, precision="fp32", task="A")
    tr.initialize()
    tr.network.load_state_dict(onet.state_dict())
    batches, keys = [], [["s0", "s1"], ["s1", "s2"], ["s0", "s2"]]
    for i, k in enumerate(keys):
        data, targets = synth.make_batch(geom, seed=100 + i)
        batches.append({"data": data, "target": targets, "keys": k})
    res = tr._perform_validation({"A": iter(batches)}, nr_batches=3)["epoch_0"]["A"]
    # oracle: per-sample hard counts of the oracle network on the same batches (MultiHead:938-951), summed per subject
    acc = {}
    with torch.no_grad():
        for b in batches:
            tp, fp, fn = cl_losses.hard_tp_fp_fn(onet(b["data"])[0], b["target"][0])
            for s, name in enumerate(b["keys"]):
                a = acc.setdefault(name, np.zeros((3, geom.num_classes - 1)))
                a += np.stack([tp[s].numpy(), fp[s].numpy(), fn[s].numpy()])
    assert sorted(res) == sorted(acc)
    for name, (tp, fp, fn) in acc.items():
        for c in range(geom.num_classes - 1):
            assert abs(res[name]['mask_%d' % (c + 1)]['Dice'] - 2 * tp[c] / (2 * tp[c] + fp[c] + fn[c])) < 1e-3
            assert abs(res[name]['mask_%d' % (c + 1)]['IoU'] - tp[c] / (tp[c] + fp[c] + fn[c])) < 1e-3
    assert tr.eval_batch is True and tr.online_eval_tp == [] and tr.network.training


def test_sliding_window_matches_the_committed_fixture():
    """tests/golden/sliding_window_tiny.npz (oracle/gen_golden_f.py): tiled prediction of the seeded tiny network with Gaussian
    weighting and full mirroring; probabilities within 1e-3, labels equal wherever the fixture's top two differ by > 2e-3"""
    import os
    import util
    from b200unet import inference
    from b200unet.configs import CONFIGS
    z = np.load(os.path.join(util.ROOT, "tests", "golden", "sliding_window_tiny.npz"))
    geom = CONFIGS["tiny"]
    cnet = cuda_net(geom, oracle_net(geom).state_dict())
    cnet.eval()
    cseg, cprob = inference.predict_3D(cnet, torch.from_numpy(z["x"]), geom.patch, True, (0, 1, 2), 0.5, True)
    prob = torch.from_numpy(z["prob"])
    assert float((cprob.cpu() - prob).abs().max()) < 1e-3
    top2 = prob.topk(2, 0).values
    decided = (top2[0] - top2[1]) > 2e-3
    assert bool((cseg.cpu()[decided] == torch.from_numpy(z["seg"]).long()[decided]).all())
