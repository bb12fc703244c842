"""Drop-in for the reference's vendored ``sync_batchnorm`` package (its __init__.py:11-12
exports exactly these five names), with the cross-replica statistics reduction moved from
the Python thread rendezvous (comm.py:18-137) onto an NCCL all-reduce."""
from .batchnorm import SynchronizedBatchNorm1d, SynchronizedBatchNorm2d, SynchronizedBatchNorm3d
from .replicate import DataParallelWithCallback, patch_replication_callback

__all__ = ["SynchronizedBatchNorm1d", "SynchronizedBatchNorm2d", "SynchronizedBatchNorm3d",
           "DataParallelWithCallback", "patch_replication_callback"]
