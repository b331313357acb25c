"""Rank / world helpers (mirror of the reference's utils/dist_util.py)."""
import torch.distributed as dist


def _on():
    return dist.is_available() and dist.is_initialized()


def get_rank():
    return dist.get_rank() if _on() else 0


def get_world_size():
    return dist.get_world_size() if _on() else 1


def is_main_process():
    return get_rank() == 0
