"""Rank / world helpers with the reference's names (dinov3_jax/distributed/__init__.py:12-21), backed by
torch.distributed (one process per GPU) instead of jax.device_count()."""
import torch.distributed as dist


def is_enabled() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_rank() -> int:
    return dist.get_rank() if is_enabled() else 0


def get_world_size() -> int:
    return dist.get_world_size() if is_enabled() else 1


def is_main_process() -> bool:
    return get_rank() == 0
