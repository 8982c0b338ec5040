"""NCCL communicator for `skyjo_stats_allreduce` (include/skyjo_b200.h), the library's only collective (SURVEY 8e:
one all-reduce of the int64[32] episode statistics per iteration, never on the step path).

torch.distributed does not hand out its ncclComm_t, so a host that wants the in-library all-reduce creates its own
communicator with the NCCL the process already has loaded (torch's bundled libnccl.so.2): rank 0 draws the unique id,
`torch.distributed` (any backend) broadcasts its 128 bytes, every rank calls ncclCommInitRank.  Without
torch.distributed (world size 1) the communicator is local.  `BatchedSkyjoEnv.stats(all_reduce=True)` keeps using
`torch.distributed.all_reduce` unless a communicator is passed."""
import ctypes as C

import torch


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]       # nccl.h: NCCL_UNIQUE_ID_BYTES


_nccl = None


def _lib():
    global _nccl
    if _nccl is None:
        L = C.CDLL("libnccl.so.2", mode=C.RTLD_GLOBAL)    # the copy torch loaded (same soname), made visible to dlsym
        L.ncclGetUniqueId.argtypes = [C.POINTER(_UniqueId)]
        L.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        L.ncclCommDestroy.argtypes = [C.c_void_p]
        L.ncclGetErrorString.restype = C.c_char_p
        L.ncclGetErrorString.argtypes = [C.c_int]
        _nccl = L
    return _nccl


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"NCCL error {rc}: {_lib().ncclGetErrorString(rc).decode()}")


class StatsComm:
    """One NCCL communicator over the ranks of `group` (default: the world), on `device`."""

    def __init__(self, device, group=None):
        import torch.distributed as dist
        L = _lib()
        self.device = torch.device(device)
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        uid = _UniqueId()
        if self.rank == 0:
            _check(L.ncclGetUniqueId(C.byref(uid)))
        if self.world > 1:
            box = [bytes(uid.internal)]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            C.memmove(uid.internal, box[0], 128)
        self._comm = C.c_void_p()
        with torch.cuda.device(self.device):
            _check(L.ncclCommInitRank(C.byref(self._comm), self.world, uid, self.rank))

    @property
    def handle(self):
        return self._comm.value

    def close(self):
        if self._comm:
            _lib().ncclCommDestroy(self._comm)
            self._comm = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
