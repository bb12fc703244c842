"""DataParallelWithCallback / patch_replication_callback — sync_batchnorm/replicate.py:27-94.

The reference replicates the module onto every visible GPU inside one process and runs the
replicas from Python threads (train.py:220).  The B200 design is one process per GPU
(torchrun), so the wrapper keeps the reference's constructor and ``.module`` attribute
(checkpoints keep their ``module.`` prefix, train.py:308-319) but runs the wrapped module on
the rank's own device only; SynchronizedBatchNorm layers inside reduce their statistics
across ranks over NCCL.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch.nn.parallel.data_parallel import DataParallel

__all__ = ["DataParallelWithCallback", "patch_replication_callback", "execute_replication_callbacks"]


def execute_replication_callbacks(modules):
    """replicate.py:27-47 — kept so that user code calling it keeps working."""
    master = modules[0]
    n = len(list(master.modules()))
    ctxs = [type("CallbackContext", (), {})() for _ in range(n)]
    for i, module in enumerate(modules):
        for j, m in enumerate(module.modules()):
            if hasattr(m, "__data_parallel_replicate__"):
                m.__data_parallel_replicate__(ctxs[j], i)


def _local_device_ids(device_ids):
    if not torch.cuda.is_available():
        return device_ids
    cur = torch.cuda.current_device()
    if dist.is_available() and dist.is_initialized():
        return [cur]
    ids = list(range(torch.cuda.device_count())) if device_ids is None else list(device_ids)
    if len(ids) > 1:
        raise RuntimeError(
            "eegan_b200.sync_batchnorm: single-process multi-GPU replication is not supported; launch one "
            "process per GPU (torchrun) and call torch.distributed.init_process_group('nccl') first, or pass "
            "device_ids=[%d]" % cur)
    return ids


class DataParallelWithCallback(DataParallel):
    def __init__(self, module, device_ids=None, output_device=None, dim=0):
        super().__init__(module, device_ids=_local_device_ids(device_ids), output_device=output_device, dim=dim)

    def replicate(self, module, device_ids):
        modules = super().replicate(module, device_ids)
        execute_replication_callbacks(modules)
        return modules


def patch_replication_callback(data_parallel):
    """replicate.py:70-94."""
    assert isinstance(data_parallel, DataParallel)
    data_parallel.device_ids = _local_device_ids(data_parallel.device_ids)
    old = data_parallel.replicate

    def new_replicate(module, device_ids):
        modules = old(module, device_ids)
        execute_replication_callbacks(modules)
        return modules

    data_parallel.replicate = new_replicate
