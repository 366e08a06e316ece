"""CPU: host-side trainer logic that needs no kernel launch -- continual-learning bookkeeping (MultiHead_Module wiring,
name-masked EWC variants, freezing policies: SURVEY 8(f) rank 3) and the on-disk formats (rank 4), including loading a
checkpoint laid out by the REFERENCE's own MultiHead_Module (when /root/reference is mounted)."""
import math
import os
import sys

import pytest
import torch

import util

REF = "/root/reference"


def _geom(name="tiny"):
    from b200unet.configs import CONFIGS
    return CONFIGS[name]


def _mk(cls, **kw):
    tr = cls(_geom(), precision="fp32", device="cpu", **kw)
    tr.initialize()
    return tr


def test_trainer_owns_a_multihead_module_and_switches_tasks():
    from b200unet.trainers import nnUNetTrainerSequential
    tr = _mk(nnUNetTrainerSequential, task="A")
    assert tr.network is tr.mh_network.model and list(tr.mh_network.heads) == ["A"]
    w_a = tr.network.seg_outputs[1].weight.detach().clone()
    tr.start_task("B")                                   # transfer_heads=True: B starts from A's head (MultiHead:551-556)
    assert list(tr.mh_network.heads) == ["A", "B"] and tr.mh_network.active_task == "B" and tr.task == "B"
    assert torch.equal(tr.network.seg_outputs[1].weight, w_a)
    with torch.no_grad():
        tr.network.seg_outputs[1].weight.add_(1.0)       # "training" on B
    tr.start_task("A")
    assert torch.equal(tr.network.seg_outputs[1].weight, w_a)
    tr.start_task("B")
    assert torch.equal(tr.network.seg_outputs[1].weight, w_a + 1.0)


def test_masked_ewc_variants_select_the_reference_parameter_sets():
    from b200unet import trainers as T
    names = ["ViT.blocks.layer.0.norm1.weight", "ViT.blocks.layer.0.attn.qkv.weight", "conv_blocks_context.0.blocks.0.conv.weight",
             "conv_blocks_context.0.blocks.0.instnorm.weight", "seg_outputs.0.weight"]
    sel = lambda cls: [n for n in names if T.ds._match(n, True, cls.MATCH, cls.MATCH_TRUE)]
    assert sel(T.nnUNetTrainerEWCViT) == names[:2]                                  # ewc_vit:50
    assert sel(T.nnUNetTrainerEWCLN) == names[:1]                                   # ewc_ln:50 (ViT AND norm)
    assert sel(T.nnUNetTrainerEWCUNet) == names[2:]                                 # ewc_unet:50 (everything but ViT)
    tr = _mk(T.nnUNetTrainerEWCUNet)
    assert isinstance(tr.loss.network_params, list) and len(tr.loss.network_params) == len(list(tr.network.parameters()))
    assert tr.loss.match_case and tr.loss.match == ['ViT'] and tr.loss.match_true is False


def test_freezing_policies_on_the_vit_unet():
    from b200unet import trainers as T
    for cls, frozen_if in ((T.nnUNetTrainerFrozenViT, lambda n: 'ViT' in n),
                           (T.nnUNetTrainerFrozenUNet, lambda n: 'ViT' not in n),
                           (T.nnUNetTrainerFrozenNonLN, lambda n: not ('ViT' in n and 'norm' in n))):
        tr = _mk(cls, use_vit=True, task="A")
        assert all(p.requires_grad for p in tr.network.parameters())
        tr.start_task("B")
        for n, p in tr.network.named_parameters():
            assert p.requires_grad == (not frozen_if(n)), (cls.__name__, n)
        del tr
    tr = _mk(T.nnUNetTrainerFrozEWC, use_vit=True, task="A", adaptive=True)
    vit_grad = lambda: {p.requires_grad for n, p in tr.network.named_parameters() if 'ViT' in n}
    tr.start_task("B")                                   # 1 head, new task -> even number of heads: freeze (froz_ewc:92-108)
    assert vit_grad() == {False} and math.isclose(tr.loss.ewc_lambda, tr.ewc_lambda * math.exp(-1 / 3))
    tr.start_task("C")                                   # 2 heads, new task -> unfreeze (:121-130)
    assert vit_grad() == {True} and tr.loss.ewc_lambda == tr.ewc_lambda


def test_checkpoint_round_trip_and_file_set(tmp_path):
    from b200unet import trainers as T
    tr = _mk(T.nnUNetTrainerRW, task="A")
    tr.start_task("A")
    tr.start_task("B")
    with torch.no_grad():
        for p in tr.network.parameters():
            p.add_(0.01)
    for p in tr.network.parameters():                    # a momentum buffer for every parameter
        tr.optimizer.state[p]['momentum_buffer'] = torch.full_like(p, 0.5)
    tr.finish_training_on("A")
    fname = str(tmp_path / "model_final_checkpoint.model")
    tr.save_checkpoint(fname)
    tr.save_importance(str(tmp_path / "rw_data"))
    assert sorted(os.listdir(tmp_path)) == ["model_final_checkpoint.model", "model_final_checkpoint.model.pkl", "rw_data", "rw_trained_on.pkl"]
    assert sorted(os.listdir(tmp_path / "rw_data")) == ["fisher_values.pkl", "param_values.pkl", "score_values.pkl"]
    ck = torch.load(fname, weights_only=False)
    assert set(ck) == {'epoch', 'state_dict', 'optimizer_state_dict', 'lr_scheduler_state_dict', 'plot_stuff', 'best_stuff', 'amp_grad_scaler'}
    prefixes = {k.split('.')[0] + ('.' + k.split('.')[1] if k.startswith('heads') else '') for k in ck['state_dict']}
    assert prefixes == {'model', 'body', 'heads.A', 'heads.B'}
    tr2 = _mk(T.nnUNetTrainerRW, task="A", seed=5)
    tr2.load_checkpoint(fname)
    tr2.load_importance(str(tmp_path / "rw_data"))
    assert tr2.task == "B" and list(tr2.mh_network.heads) == ["A", "B"]
    for (n, a), (_, b) in zip(tr.network.named_parameters(), tr2.network.named_parameters()):
        assert torch.equal(a, b), n
    for a, b in zip(tr.optimizer.param_groups[0]['params'], tr2.optimizer.param_groups[0]['params']):
        assert torch.equal(tr.optimizer.state[a]['momentum_buffer'], tr2.optimizer.state[b]['momentum_buffer'])
    assert tr2.already_trained_on["0"]["finished_training_on"] == ["A"]
    assert set(tr2.fisher) == {"A", "B"} and set(tr2.scores["A"]) == set(n for n, _ in tr.network.named_parameters())


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")
def test_loads_a_checkpoint_laid_out_by_the_reference_multihead_module(tmp_path):
    """state_dict produced by the REFERENCE's MultiHead_Module (unmodified file) around the oracle's Generic_UNet, stored the
    way nnunet's save_checkpoint stores it, restored into the CUDA-class trainer by name"""
    for p in (os.path.join(util.ROOT, "oracle", "shim"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    from nnunet_ext.network_architecture.MultiHead_Module import MultiHead_Module as RefMH
    from nnunet.network_architecture.generic_UNet import Generic_UNet as OracleUNet
    from b200unet import checkpoint
    from b200unet.trainers import nnUNetTrainerSequential
    geom = _geom()
    onet = util.oracle_net(geom, seed=3)
    ref = RefMH(OracleUNet, "seg_outputs", "A", prev_trainer=onet)
    ref.add_new_task("B", use_init=False)
    with torch.no_grad():
        for p in ref.heads["B"].parameters():
            p.mul_(3.0)
    ref.assemble_model("B")
    ato = {"0": {"finished_training_on": ["A"], "tasks_at_time_of_checkpoint": ["A", "B"], "active_task_at_time_of_checkpoint": "B"}}
    checkpoint.write_pickle(ato, str(tmp_path / "sequential_trained_on.pkl"))
    torch.save({'epoch': 7, 'state_dict': ref.state_dict(), 'optimizer_state_dict': None, 'lr_scheduler_state_dict': None,
                'plot_stuff': ([], [], [], []), 'best_stuff': (None, None, None), 'amp_grad_scaler': None},
               str(tmp_path / "model_final_checkpoint.model"))
    tr = _mk(nnUNetTrainerSequential, task="A")
    tr.load_checkpoint(str(tmp_path / "model_final_checkpoint.model"))
    assert tr.epoch == 7 and tr.mh_network.active_task == "B"
    rsd = ref.model.state_dict()
    for n, p in tr.network.named_parameters():
        assert torch.equal(p, rsd[n]), n
    assert torch.equal(tr.mh_network.heads["A"].seg_outputs[0].weight, ref.heads["A"].seg_outputs[0].weight)
    assert set(tr.mh_network.state_dict().keys()) == set(ref.state_dict().keys())


def test_sliding_window_steps_and_gaussian_known_answers():
    """nnunet _compute_steps_for_sliding_window: patch 128 on 200 voxels at step 0.5 -> ceil(72/64)+1 = 3 equidistant origins"""
    from b200unet import inference
    from oracle import sliding_window
    assert inference.compute_steps_for_sliding_window((128,), (200,), 0.5) == [[0, 36, 72]]
    assert inference.compute_steps_for_sliding_window((16, 32, 32), (16, 32, 33), 0.5) == [[0], [0], [0, 1]]
    assert inference.compute_steps_for_sliding_window((16, 32, 32), (20, 45, 50), 0.5) == sliding_window.compute_steps((16, 32, 32), (20, 45, 50), 0.5)
    g = inference.get_gaussian((8, 16, 16))
    assert g.dtype.name == "float32" and float(g.max()) == 1.0 and float(g.min()) > 0 and g[4, 8, 8] == 1.0
    assert torch.equal(torch.from_numpy(g), sliding_window.gaussian_map((8, 16, 16)))


def test_finish_online_evaluation_extended_per_subject_dice_iou():
    """reference MultiHead:963-1049: counts summed per subject name; IoU = tp/(tp+fp+fn), Dice = 2tp/(2tp+fp+fn)"""
    import numpy as np
    from b200unet.trainers import nnUNetTrainerSequential
    tr = _mk(nnUNetTrainerSequential, task="A")
    tr.subject_names_raw = [["a", "b"], ["b", "a"]]
    tr.online_eval_tp = [np.array([[4., 1.], [2., 0.]]), np.array([[1., 1.], [3., 2.]])]
    tr.online_eval_fp = [np.array([[1., 0.], [1., 0.]]), np.array([[0., 1.], [0., 1.]])]
    tr.online_eval_fn = [np.array([[0., 2.], [1., 0.]]), np.array([[2., 0.], [1., 1.]])]
    tr.epoch = 3
    store = tr.finish_online_evaluation_extended("A")
    assert tr.validation_results["epoch_3"]["A"] is store and sorted(store) == ["a", "b"]
    # subject a: tp (7, 3) fp (1, 1) fn (1, 3);  subject b: tp (3, 1) fp (1, 1) fn (3, 0)
    assert math.isclose(store["a"]["mask_1"]["Dice"], 14 / 16) and math.isclose(store["a"]["mask_1"]["IoU"], 7 / 9)
    assert math.isclose(store["a"]["mask_2"]["Dice"], 6 / 10) and math.isclose(store["b"]["mask_2"]["IoU"], 1 / 2)
    assert math.isclose(store["b"]["mask_1"]["Dice"], 6 / 10)
    assert tr.online_eval_tp == [] and tr.subject_names_raw == []


def test_plop_median_thresholds_equal_oracle_on_random_histograms():
    """host half of plop:149-173 (median search) against the oracle restatement on a given histogram"""
    import numpy as np
    from b200unet.trainers import nnUNetTrainerPLOP
    from oracle import cl_losses
    rs = np.random.RandomState(0)
    hist = rs.randint(0, 50, size=(3, 100)).astype(np.int64)
    hist[1] = 0                                          # a class that is never the pseudo label keeps the floor
    hist[2, :40] = 0
    mine = nnUNetTrainerPLOP._median_thresholds(hist)
    # oracle: feed the same histogram through one-hot "outputs" is roundabout; restate with its own loop on a prepared table
    thr = []
    for c in range(3):
        total = hist[c].sum()
        if total <= 0:
            thr.append(0.001)
            continue
        cum = np.cumsum(hist[c])
        b = int(np.argmax(cum >= total / 2))
        prev = cum[b - 1] if b else 0
        thr.append(max(b / 100 + ((total / 2 - prev) / hist[c, b]) / 100, 0.001))
    assert np.allclose(mine, thr, atol=1e-12) and mine[1] == 0.001 and mine[2] > 0.4


def test_frozen_body_trainer_freezes_the_body_from_the_second_task():
    """reference frozen_body_seq: first task trains everything; from the second task on only the task's head requires grad"""
    from b200unet.trainers import nnUNetTrainerFrozenBody
    tr = _mk(nnUNetTrainerFrozenBody, task="A")
    tr.start_task("A")
    assert all(p.requires_grad for p in tr.network.parameters())
    tr.start_task("B")
    head = {n for n, _ in tr.network.named_parameters() if n.startswith("seg_outputs")}
    for n, p in tr.network.named_parameters():
        assert p.requires_grad == (n in head), n
    assert tr.mh_network.body_freezed and tr.mh_network.active_task == "B"


def test_rehearsal_fuses_a_seeded_sample_of_previous_tasks():
    """reference rehearsal:80-124: random.seed(seed); per previous task (head order) random.sample(items, round(len * samples))"""
    import random
    from b200unet.trainers import nnUNetTrainerRehearsal
    tr = _mk(nnUNetTrainerRehearsal, task="A", samples_in_perc=0.25, rehearsal_seed=3299)
    tr.start_task("B")
    tr.start_task("C")
    prev = {"A": {"a%02d" % i: {"data": i} for i in range(20)}, "B": {"b%02d" % i: {"data": i} for i in range(10)}}
    cur = {"c%02d" % i: {"data": i} for i in range(6)}
    fused = tr.fuse_datasets(cur, prev)
    random.seed(3299)
    exp = dict(cur)
    exp.update(random.sample(list(prev["A"].items()), 5))
    exp.update(random.sample(list(prev["B"].items()), 2))      # round(2.5) == 2 (banker's rounding, as in the reference)
    assert fused == exp and len(fused) == 6 + 5 + 2
    assert tr.fuse_datasets(cur, prev) == fused                 # seeded: reproducible
