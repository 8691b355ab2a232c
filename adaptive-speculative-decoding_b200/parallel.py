"""Tensor-parallel plumbing: one process per GPU, ``torch.distributed`` for rendezvous, and a raw
NCCL communicator (created through ctypes on the NCCL library torch already loaded) whose
``ncclAllReduce`` address is handed to the C engine (``asd_engine_set_allreduce``), which calls it on
the two row-parallel boundaries of every layer.

Mirrors the reference's ``tensor_parallel_size`` knob, which it passes to vLLM
(/root/reference/src/serving/real_model_pipeline.py:100; configs/qwen3_models.yaml:10,22,34,46)."""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist


class _UniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_byte * 128)]


class NcclComm:
    def __init__(self, rank: int, world: int, group=None):
        self.rank, self.world = rank, world
        self.lib = ctypes.CDLL("libnccl.so.2")
        self.lib.ncclGetUniqueId.restype = ctypes.c_int
        self.lib.ncclGetUniqueId.argtypes = [ctypes.POINTER(_UniqueId)]
        self.lib.ncclCommInitRank.restype = ctypes.c_int
        self.lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _UniqueId, ctypes.c_int]
        self.lib.ncclCommDestroy.restype = ctypes.c_int
        self.lib.ncclCommDestroy.argtypes = [ctypes.c_void_p]
        uid = _UniqueId()
        if rank == 0:
            rc = self.lib.ncclGetUniqueId(ctypes.byref(uid))
            if rc != 0:
                raise RuntimeError(f"ncclGetUniqueId failed: {rc}")
        box = [bytes(uid.internal) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        ctypes.memmove(ctypes.byref(uid), box[0], 128)
        self.comm = ctypes.c_void_p()
        rc = self.lib.ncclCommInitRank(ctypes.byref(self.comm), world, uid, rank)
        if rc != 0:
            raise RuntimeError(f"ncclCommInitRank failed: {rc}")

    @property
    def comm_ptr(self) -> int:
        return self.comm.value

    @property
    def allreduce_fn_ptr(self) -> int:
        return ctypes.cast(self.lib.ncclAllReduce, ctypes.c_void_p).value

    def destroy(self):
        if self.comm:
            self.lib.ncclCommDestroy(self.comm)
            self.comm = ctypes.c_void_p()


def init_distributed(backend: str = "nccl"):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun); returns (rank, local_rank, world)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def shard_ranges(n: int, world: int):
    """contiguous Megatron-style split of n units over the ranks"""
    assert n % world == 0, (n, world)
    per = n // world
    return [(r * per, (r + 1) * per) for r in range(world)]
