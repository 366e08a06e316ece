"""ORACLE (test infrastructure). Restates nnunet@77bc485 nnunet/utilities/to_torch.py
(SURVEY.md Appendix A).  ``to_cuda`` is a no-op for CPU runs: the reference passes
``gpu_id=loss.get_device()`` == -1 on CPU (reference deep_supervision.py:74-76)."""
import numpy as np
import torch


def maybe_to_torch(d):
    if isinstance(d, (list, tuple)):
        return [maybe_to_torch(i) for i in d]
    if isinstance(d, np.ndarray):
        return torch.from_numpy(d).float()
    return d


def to_cuda(data, non_blocking=True, gpu_id=0):
    def _mv(t):
        if not torch.cuda.is_available() or (isinstance(gpu_id, int) and gpu_id < 0):
            return t
        return t.cuda(gpu_id, non_blocking=non_blocking)
    if isinstance(data, (list, tuple)):
        return [_mv(i) for i in data]
    return _mv(data)
