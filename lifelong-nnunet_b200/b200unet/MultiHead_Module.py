"""``MultiHead_Module`` -- body / head bookkeeping of the multi-head network, mirror of the reference's
nnunet_ext/network_architecture/MultiHead_Module.py (ctor :16-125, forward :127-137, update_after_iteration :139-157,
split :159-324, assemble_model :326-377, _set_requires_grad :379-395, add_new_task :435-458,
add_n_tasks_and_activate :460-485, getters / setters / replace_layers :487-571).

Same observable behaviour (names, arguments, asserts, ``state_dict`` keys ``model.* / body.* / heads.<task>.*``):

* a network is split at a dotted path; everything registered BEFORE the split point (depth first, in registration
  order) is the shared ``body``, the split point and everything registered after it is a ``head``;
* ``heads`` is a ``ModuleDict`` task -> head whose module paths are those of the full network
  (``heads[t].seg_outputs[i]`` ...), ``body`` shares its sub-module objects with the running ``model``;
* ``assemble_model(task)`` loads body + head ``task`` into the running model IN PLACE and (un)freezes the body;
* ``forward`` is ``class_object.forward(self.model, x)`` -- for ``b200unet.Generic_UNet`` that is the CUDA plan.

What is different is the cost on the per-iteration path (the reference re-splits the model
recursively and ``copy.deepcopy``s the head after EVERY training iteration, 0.6-4 ms of host time against a ~5 ms GPU
step): here the ACTIVE head is not a copy -- its sub-modules are the running model's own sub-module objects, so
``update_after_iteration()`` has nothing to do; a head becomes a private snapshot (deep copy) at the moment another
task is activated.
"""
import copy
from operator import attrgetter
from typing import Type

from torch import nn


def _get(module, path):
    return attrgetter('.'.join(path))(module) if path else module


class MultiHead_Module(nn.Module):
    def __init__(self, class_object: Type[nn.Module], split_at, task, prev_trainer=None, *args, **kwargs):
        super().__init__()
        self.class_object = class_object
        if prev_trainer is None:
            self.model = self.class_object(*args, **kwargs)
        else:
            assert isinstance(prev_trainer, self.class_object), \
                "This function splits a \'{}\' module class object, but a \'{}\' module is provided.".format(
                    self.class_object.__name__, type(prev_trainer))
            assert len(list(prev_trainer.children())) > 0, \
                "When using a prev_trainer, please ensure that it is not empty or do not specify one."
            self.model = prev_trainer
        assert isinstance(split_at, str), "The provided split needs to be a string.."
        self.split = [x.strip() for x in split_at.split('.')]
        self._check_and_simplify_split(self.model)
        self.heads = nn.ModuleDict()
        assert isinstance(task, (str, int)), "The provided task needs to be an integer (ID) or string, not {}..".format(type(task))
        self.active_task = task
        self.body, head = self._split(self.model, share_head=True)
        self.state_init = {k: v.detach().clone() for k, v in head.state_dict().items()}
        self.heads[self.active_task] = head
        self._set_requires_grad(True)      # the reference's first assemble_model(task, freeze_body=False) unfreezes the body (:120-124)
        self.body_freezed = False

    # ------------------------------------------------------------------------------------------------------------
    def _check_and_simplify_split(self, model):
        """reference :72-97 -- the path must exist; a trailing element that names the FIRST child of its parent is
        redundant and dropped; splitting before the very first layer would leave an empty body"""
        try:
            _ = _get(model, self.split)
        except Exception:
            assert False, "The provided split path \'{}\' does not exist..".format('.'.join(self.split))
        original = self.split[:]
        while len(self.split) > 1:
            first_layer_name, _ = next(_get(model, self.split[:-1]).named_children())
            if self.split[-1] != first_layer_name:
                break
            new_split = self.split[:-1]
            print('Note: The split \'{}\' has been transformed to \'{}\' since it specifies the same split.\n'.format(
                '.'.join(self.split), '.'.join(new_split)))
            self.split = new_split
        first = next(model.named_children())
        assert not (len(self.split) == 1 and self.split[0] == first[0]), \
            "You tried to split before the first layer, so the body would be empty --> body can never be empty.. " \
            "(split \'{}\')".format('.'.join(original))

    def _members(self, model):
        """(body members, head members) as lists of (path tuple, module): depth-first along the split path, children
        registered before the path node belong to the body, the split node and everything after it to the head"""
        body, head = [], []

        def walk(mod, depth, prefix):
            hit = False
            for name, child in mod.named_children():
                if hit:
                    head.append((prefix + (name,), child))
                elif name == self.split[depth]:
                    hit = True
                    if depth == len(self.split) - 1:
                        head.append((prefix + (name,), child))
                    else:
                        walk(child, depth + 1, prefix + (name,))
                else:
                    body.append((prefix + (name,), child))
            assert hit, "The provided split path \'{}\' does not exist..".format('.'.join(self.split))
        walk(model, 0, ())
        return body, head

    @staticmethod
    def _skeleton(members, transform):
        """an nn.Module tree holding `members` at their original paths (plain nn.Module containers in between)"""
        root = nn.Module()
        for path, mod in members:
            cur = root
            for p in path[:-1]:
                if p not in cur._modules:
                    cur.add_module(p, nn.Module())
                cur = cur._modules[p]
            cur.add_module(path[-1], transform(mod))
        return root

    def _split(self, model, share_head=False):
        body_m, head_m = self._members(model)
        body = self._skeleton(body_m, lambda m: m)                  # shared objects (Q13)
        head = self._skeleton(head_m, (lambda m: m) if share_head else copy.deepcopy)
        return body, head

    def _split_model_recursively_into_body_head(self, layer_id, model, body=None, head=None, parent=None,
                                                simplify_split=False):
        """reference :159-324 (kept for callers that use it directly): (body, deep copy of the head, len(split), [])"""
        if simplify_split:
            self._check_and_simplify_split(model)
        b, h = self._split(model, share_head=False)
        return b, h, len(self.split), []

    # ------------------------------------------------------------------------------------------------------------
    def forward(self, x):
        """reference :127-137"""
        return self.class_object.forward(self.model, x)

    def update_after_iteration(self, model=None, update_body=True):
        """reference :139-157.  With the running model itself nothing has to be copied: the active head's sub-modules
        ARE the model's sub-modules.  A foreign `model` is split and copied like the reference does."""
        if model is None or model is self.model:
            return
        body, head = self._split(model, share_head=False)
        if update_body:
            self.body = body
        self.heads[self.active_task] = head

    def _snapshot_active_head(self):
        """the active head stops aliasing the running model: it becomes a private deep copy"""
        if self.active_task in self.heads:
            _, head_m = self._members(self.model)
            shared = {id(m) for _, m in head_m}
            cur = self.heads[self.active_task]
            if any(id(m) in shared for m in cur.modules()):
                self.heads[self.active_task] = copy.deepcopy(cur)

    def assemble_model(self, task, freeze_body=False):
        """reference :326-377"""
        if self.active_task == task and freeze_body == self.body_freezed:
            return self.model
        assert task in self.heads.keys(), \
            "The provided task \'{}\' is not a known head, so either initialize the task or provide one that already exists: {}.".format(
                task, list(self.heads.keys()))
        if task != self.active_task:
            self._snapshot_active_head()
            head = self.heads[task]
            _, head_m = self._members(self.model)
            for path, mod in head_m:                                  # load the stored head into the running model in place
                mod.load_state_dict(_get(head, path).state_dict())
            self.heads[task] = self._skeleton(head_m, lambda m: m)    # ... and alias it
            self.active_task = task
        if freeze_body and not self.body_freezed:
            self._set_requires_grad(False)
            self.body_freezed = True
        if not freeze_body and self.body_freezed:
            self._set_requires_grad(True)
            self.body_freezed = False
        return self.model

    def _set_requires_grad(self, requires_grad):
        """reference :379-395"""
        body_parameters = set(name for name, _ in self.body.named_parameters())
        for name, param in self.model.named_parameters():
            if name in body_parameters:
                param.requires_grad = requires_grad

    def add_new_task(self, task, use_init, model=None):
        """reference :435-458"""
        if model is None:
            self.heads[task] = copy.deepcopy(self.heads[list(self.heads.keys())[-1]])
            if use_init:
                self.heads[task].load_state_dict(self.state_init)
        else:
            self.heads[task] = copy.deepcopy(model)

    def add_n_tasks_and_activate(self, list_of_tasks, activate_with, remove_old_tasks=True):
        """reference :460-485"""
        for task in list_of_tasks:
            if task not in self.heads:
                self.add_new_task(task, use_init=True)
        if remove_old_tasks:
            for task in list(self.heads.keys()):
                if task not in list_of_tasks:
                    if task == self.active_task:
                        self._snapshot_active_head()
                    del self.heads[task]
        self.assemble_model(activate_with)

    # -- getters / setters (reference :487-571) -----------------------------------------------------------------------
    def get_heads(self):
        return copy.deepcopy(self.heads)

    def get_body(self):
        return copy.deepcopy(self.body)

    def set_heads(self, heads, reset=True):
        assert isinstance(heads, nn.ModuleDict), "Provided heads are not a nn.ModuleDict."
        if reset:
            del self.heads
            self.heads = heads
        else:
            self.heads.update(heads)

    def set_body(self, body):
        assert isinstance(body, nn.Module), "Provided body is not a nn.Module.."
        del self.body
        self.body = copy.deepcopy(body)

    def get_model_type(self):
        return self.model.__class__.__name__

    def get_split_path(self):
        return '.'.join(self.split)

    def replace_layers(self, model, old, new):
        assert model is not None and new is not None and old is not None, \
            "To replace a Module, the layers need to be Modules as well as the model.."
        for name, module in model.named_children():
            if len(list(module.children())) > 0:
                self.replace_layers(module, old, new)
            if isinstance(module, old):
                setattr(model, name, new)
        return model
