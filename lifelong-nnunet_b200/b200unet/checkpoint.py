"""On-disk formats of the reference trainers, written and read compatibly:

* ``<fname>`` (``model_final_checkpoint.model`` / ``model_latest.model`` / ``model_old.model``): ``torch.save`` of
  ``{'epoch', 'state_dict', 'optimizer_state_dict', 'lr_scheduler_state_dict', 'plot_stuff', 'best_stuff',
  'amp_grad_scaler'}`` (nnunet NetworkTrainer.save_checkpoint, reached from reference
  .../multihead/nnUNetTrainerMultiHead.py:1164-1197).  ``state_dict`` is the state of the WHOLE ``MultiHead_Module``
  (keys ``model.*``, ``body.*``, ``heads.<task>.*``), because the reference swaps ``self.network = self.mh_network``
  before saving (:1170); ``model_old.model`` holds the plain teacher network (:1186-1190).
* ``<fname>.pkl``: ``{'init': init args, 'name': class name, 'class': str(class), 'plans': plans}`` (nnunet
  ``save_checkpoint`` / reference ``update_init_args`` :1199-1215).
* ``<ext>_trained_on.pkl``: the ``already_trained_on`` bookkeeping incl. ``tasks_at_time_of_checkpoint`` and
  ``active_task_at_time_of_checkpoint`` (:1173-1178), needed to rebuild the heads before ``load_state_dict`` (:1285-1286).
* ``fisher_values.pkl`` / ``param_values.pkl`` (EWC, reference ewc:205-228) and ``score_values.pkl`` (RW, rw:267-314):
  plain pickles of ``{task: {parameter name: tensor}}``.
"""
import os
import pickle
from collections import OrderedDict

import torch


def write_pickle(obj, path):
    """batchgenerators.utilities.file_and_folder_operations.write_pickle"""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, 'wb') as f:
        pickle.dump(obj, f)


def load_pickle(path):
    with open(path, 'rb') as f:
        return pickle.load(f)


def _cpu_state(sd):
    out = OrderedDict()
    for k, v in sd.items():
        out[k] = v.detach().cpu() if torch.is_tensor(v) else v
    return out


def save_checkpoint(trainer, fname, save_optimizer=True):
    """reference MultiHead:1164-1197 (+ nnunet NetworkTrainer.save_checkpoint)"""
    fold = str(getattr(trainer, 'fold', 0))
    ato = trainer.already_trained_on.setdefault(fold, dict())
    ato['checkpoint_should_exist'] = True
    ato['tasks_at_time_of_checkpoint'] = list(trainer.mh_network.heads.keys())
    ato['active_task_at_time_of_checkpoint'] = trainer.mh_network.active_task
    out_dir = os.path.dirname(os.path.abspath(fname))
    write_pickle(trainer.already_trained_on, os.path.join(out_dir, trainer.extension + '_trained_on.pkl'))
    opt = trainer.optimizer.state_dict() if (save_optimizer and trainer.optimizer is not None) else None
    save_this = {'epoch': trainer.epoch + 1, 'state_dict': _cpu_state(trainer.mh_network.state_dict()),
                 'optimizer_state_dict': opt, 'lr_scheduler_state_dict': None,
                 'plot_stuff': (getattr(trainer, 'all_tr_losses', []), getattr(trainer, 'all_val_losses', []),
                                getattr(trainer, 'all_val_losses_tr_mode', []), getattr(trainer, 'all_val_eval_metrics', [])),
                 'best_stuff': (None, None, None), 'amp_grad_scaler': None}
    torch.save(save_this, fname)
    info = OrderedDict(init=trainer.init_args(), name=type(trainer).__name__, plans=getattr(trainer, 'plans', None))
    info['class'] = str(type(trainer))
    write_pickle(info, fname + ".pkl")
    if getattr(trainer, 'network_old', None) is not None:     # teacher of MiB / PLOP / POD (:1186-1190)
        old = dict(save_this)
        old['state_dict'] = _cpu_state(trainer.network_old.state_dict())
        torch.save(old, os.path.join(out_dir, "model_old.model"))


def load_checkpoint(trainer, fname, train=True, fname_old=None):
    """reference MultiHead:1278-1313: rebuild every head, load the MultiHead_Module state, optionally the teacher"""
    out_dir = os.path.dirname(os.path.abspath(fname))
    tpath = os.path.join(out_dir, trainer.extension + '_trained_on.pkl')
    if os.path.exists(tpath):
        trainer.already_trained_on = load_pickle(tpath)
    fold = str(getattr(trainer, 'fold', 0))
    ato = trainer.already_trained_on[fold]
    ckpt = torch.load(fname, map_location='cpu', weights_only=False)
    trainer.mh_network.add_n_tasks_and_activate(ato['tasks_at_time_of_checkpoint'], ato['active_task_at_time_of_checkpoint'])
    # keys may carry a DataParallel 'module.' prefix (nnunet strips it); map by suffix otherwise exactly
    sd = OrderedDict((k[7:] if k.startswith('module.') else k, v) for k, v in ckpt['state_dict'].items())
    trainer.mh_network.load_state_dict(sd)
    trainer.network = trainer.mh_network.model
    trainer.task = trainer.mh_network.active_task
    trainer.epoch = ckpt['epoch']
    if train and ckpt.get('optimizer_state_dict') is not None and trainer.optimizer is not None:
        trainer.optimizer.load_state_dict(ckpt['optimizer_state_dict'])
    if ckpt.get('plot_stuff') is not None:
        (trainer.all_tr_losses, trainer.all_val_losses, trainer.all_val_losses_tr_mode, trainer.all_val_eval_metrics) = ckpt['plot_stuff']
    if fname_old is None and os.path.exists(os.path.join(out_dir, "model_old.model")) and hasattr(trainer, 'network_old'):
        fname_old = os.path.join(out_dir, "model_old.model")
    if fname_old is not None and hasattr(trainer, 'make_teacher'):
        trainer.make_teacher()
        old = torch.load(fname_old, map_location='cpu', weights_only=False)
        trainer.network_old.load_state_dict(old['state_dict'])
    trainer._steps = {}
    return ckpt


def save_importance(trainer, path):
    """EWC: fisher_values.pkl + param_values.pkl (ewc:205-228); RW additionally score_values.pkl (rw:267-314)"""
    fold = str(getattr(trainer, 'fold', 0))
    ato = trainer.already_trained_on.setdefault(fold, dict())
    cpu = lambda d: {t: {k: v.detach().cpu() for k, v in m.items()} for t, m in d.items()}
    write_pickle(cpu(trainer.fisher), os.path.join(path, 'fisher_values.pkl'))
    write_pickle(cpu(trainer.params), os.path.join(path, 'param_values.pkl'))
    ato['fisher_at'] = os.path.join(path, 'fisher_values.pkl')
    ato['params_at'] = os.path.join(path, 'param_values.pkl')
    if hasattr(trainer, 'scores'):
        write_pickle(cpu(trainer.scores), os.path.join(path, 'score_values.pkl'))
        ato['scores_at'] = os.path.join(path, 'score_values.pkl')


def load_importance(trainer, path):
    dev = trainer.device
    to = lambda d: {t: {k: v.to(dev) for k, v in m.items()} for t, m in d.items()}
    trainer.fisher.clear(); trainer.fisher.update(to(load_pickle(os.path.join(path, 'fisher_values.pkl'))))
    trainer.params.clear(); trainer.params.update(to(load_pickle(os.path.join(path, 'param_values.pkl'))))
    if hasattr(trainer, 'scores') and os.path.exists(os.path.join(path, 'score_values.pkl')):
        trainer.scores.clear(); trainer.scores.update(to(load_pickle(os.path.join(path, 'score_values.pkl'))))
        trainer.loss.update_rw_params(trainer.fisher, trainer.params, trainer.scores)
    else:
        trainer.loss.update_ewc_params(trainer.fisher, trainer.params)
    trainer._steps = {}
