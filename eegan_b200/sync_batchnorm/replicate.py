"""DataParallelWithCallback / patch_replication_callback — sync_batchnorm/replicate.py:27-94.

The reference replicates the module onto every visible GPU inside one process and runs the
replicas from Python threads (train.py:220).  The B200 design is one process per GPU
(torchrun), so the wrapper keeps the reference's constructor and ``.module`` attribute
(checkpoints keep their ``module.`` prefix, train.py:308-319) but runs the wrapped module on
the rank's own device only; SynchronizedBatchNorm layers inside reduce their statistics
across ranks over NCCL.

Parameter gradients.  nn.DataParallel SUMS the replicas' parameter gradients into the one
parameter set of the master copy.  With one process per GPU that sum has to cross processes:
when a process group is active the wrapper registers a post-accumulate hook on every
parameter that all-reduces (SUM) its gradient over the group, so every rank ends a backward
with the gradient the reference's single parameter set would hold and the ranks' optimisers
stay in lock-step.  Scale: the sharded losses (eegan_b200.sharded) already hand every rank the
gradient of the GLOBAL-batch loss with respect to its own samples, exactly what a DataParallel
replica receives from the scatter of the gathered outputs' gradient — hence SUM, not the MEAN of
DistributedDataParallel.  A loss that is a per-rank mean over local samples must be divided by
the world size by the caller (``grad_sync="mean"`` does that in the hook).  ``grad_sync=None``
switches the hooks off (the caller then owns gradient synchronisation).  The reduction is ONE
all-reduce of the flattened gradients at the end of each backward pass, applied to ``p.grad`` as it
stands then: zero the gradients before every backward (train.py:499 does), as a gradient carried
over from an earlier pass would be summed over the ranks a second time.
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


def _register_grad_sync(module, mode, group):
    """all-reduce every parameter's gradient over ``group`` right after autograd accumulated it."""
    handles = []
    if mode is None or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return handles
    if mode not in ("sum", "mean"):
        raise ValueError("grad_sync must be 'sum', 'mean' or None")
    world = dist.get_world_size(group)
    params = [p for p in module.parameters() if p.requires_grad]
    state = {"queued": False}

    def flush():
        # end of the backward pass: ONE all-reduce over the flattened gradients (a generator has ~100 parameter tensors; a
        # collective per tensor costs the host more than the whole reduction costs the links)
        state["queued"] = False
        grads = [p.grad for p in params if p.grad is not None]
        if not grads:
            return
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if mode == "mean":
            flat.div_(world)
        torch._foreach_copy_(grads, list(torch._utils._unflatten_dense_tensors(flat, grads)))

    def hook(p):
        if not state["queued"]:  # first gradient of this backward pass: run flush() once the pass is complete
            state["queued"] = True
            torch.autograd.Variable._execution_engine.queue_callback(flush)

    for p in params:
        handles.append(p.register_post_accumulate_grad_hook(hook))
    return handles


class DataParallelWithCallback(DataParallel):
    def __init__(self, module, device_ids=None, output_device=None, dim=0, grad_sync="sum", process_group=None):
        super().__init__(module, device_ids=_local_device_ids(device_ids), output_device=output_device, dim=dim)
        self._grad_sync_handles = _register_grad_sync(module, grad_sync, process_group)

    def forward(self, *inputs, **kwargs):
        if not torch.cuda.is_available() or not self.device_ids:
            return self.module(*inputs, **kwargs)  # CPU (tests over gloo): DataParallel's own pass-through
        return super().forward(*inputs, **kwargs)

    def replicate(self, module, device_ids):
        modules = super().replicate(module, device_ids)
        execute_replication_callbacks(modules)
        return modules


def patch_replication_callback(data_parallel, grad_sync="sum", process_group=None):
    """replicate.py:70-94."""
    assert isinstance(data_parallel, DataParallel)
    data_parallel.device_ids = _local_device_ids(data_parallel.device_ids)
    if not getattr(data_parallel, "_grad_sync_handles", None):
        data_parallel._grad_sync_handles = _register_grad_sync(data_parallel.module, grad_sync, process_group)
    old = data_parallel.replicate

    def new_replicate(module, device_ids):
        modules = old(module, device_ids)
        execute_replication_callbacks(modules)
        return modules

    data_parallel.replicate = new_replicate
