"""One context per device per process (include/orc_b200.h: orc_ctx_*)."""
import ctypes as C
import os

from . import _lib


class Context:
    def __init__(self, device=0, stream=None):
        """`stream`: a cudaStream_t as an int (e.g. torch.cuda.current_stream().cuda_stream); None = library-owned stream."""
        self._h = C.c_void_p()
        _lib.check(_lib.lib().orc_ctx_create(C.c_int32(device), C.c_void_p(stream or 0), C.byref(self._h)))
        self.device = device

    @property
    def handle(self):
        return self._h

    def launch_count(self):
        return int(_lib.lib().orc_ctx_launch_count(self._h))

    def synchronize(self):
        _lib.check(_lib.lib().orc_ctx_synchronize(self._h))

    def comm_init(self, rank=None, world_size=None):
        """Multi-GPU: create this context's NCCL communicator. The 128-byte NCCL id is made on rank 0 and broadcast with
        torch.distributed (any backend). One process per GPU, launched by torchrun."""
        import torch
        import torch.distributed as dist
        rank = dist.get_rank() if rank is None else rank
        world_size = dist.get_world_size() if world_size is None else world_size
        buf = (C.c_char * 128)()
        if world_size > 1:
            if rank == 0:
                _lib.check(_lib.lib().orc_comm_unique_id(buf))
            dev = torch.device("cuda", self.device) if dist.get_backend() == "nccl" else torch.device("cpu")
            t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
            dist.broadcast(t, 0)
            raw = bytes(t.cpu().tolist())
            buf = (C.c_char * 128).from_buffer_copy(raw)
        _lib.check(_lib.lib().orc_ctx_comm_init(self._h, C.c_int32(rank), C.c_int32(world_size), buf))
        self.rank, self.world_size = rank, world_size
        self.peer = False
        if world_size > 1 and os.environ.get("ORC_B200_PEER", "1") != "0":
            self._peer_init(dist, torch)

    def _peer_init(self, dist, torch):
        """Symmetric peer windows over CUDA IPC (include/orc_b200.h: orc_ctx_peer_*): handles travel over torch.distributed. All ranks
        end in the same state: the peer path is on only if every rank could map every window."""
        dev = torch.device("cuda", self.device) if dist.get_backend() == "nccl" else torch.device("cpu")
        handle = (C.c_char * 64)()
        ok = _lib.lib().orc_ctx_peer_window(self._h, handle) == _lib.OK
        mine = torch.tensor(list(bytes(handle)) + [1 if ok else 0], dtype=torch.uint8, device=dev)
        allh = [torch.zeros_like(mine) for _ in range(self.world_size)]
        dist.all_gather(allh, mine)
        rows = [bytes(t.cpu().tolist()) for t in allh]
        ok = all(r[64] == 1 for r in rows)
        if ok:
            blob = b"".join(r[:64] for r in rows)
            ok = _lib.lib().orc_ctx_peer_open(self._h, C.c_char_p(blob)) == _lib.OK
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            self.peer = True
        else:
            _lib.lib().orc_ctx_peer_disable(self._h)

    PROF_CLASSES = ("spmv", "vector", "assembly", "restriction", "galerkin", "scaling", "other", "bicgstab")

    def prof_enable(self, on=True):
        _lib.check(_lib.lib().orc_prof_enable(self._h, C.c_int32(1 if on else 0)))

    def prof_config(self, classes=None, sample_every=1):
        """Time only `classes` (names from PROF_CLASSES; None = all), and only every `sample_every`-th launch of each."""
        mask = 0xFFFFFFFF if classes is None else sum(1 << self.PROF_CLASSES.index(k) for k in classes)
        _lib.check(_lib.lib().orc_prof_config(self._h, C.c_uint32(mask), C.c_uint32(sample_every)))

    def prof_get(self):
        """{class: (ms, algorithmic bytes, launches)} accumulated since prof_enable."""
        n = len(self.PROF_CLASSES)
        ms, by, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_uint64 * n)()
        _lib.check(_lib.lib().orc_prof_get(self._h, ms, by, cnt, C.c_int32(n)))
        return {k: (ms[i], by[i], int(cnt[i])) for i, k in enumerate(self.PROF_CLASSES)}

    def prof_ref_bytes(self, cls="spmv"):
        """Bytes of the timed launches of `cls` counted in the reference's units (a lockstep SpMV = 3 reference SpMVs)."""
        n = len(self.PROF_CLASSES)
        by = (C.c_double * n)()
        _lib.check(_lib.lib().orc_prof_get_ref_bytes(self._h, by, C.c_int32(n)))
        return by[self.PROF_CLASSES.index(cls)]

    def prof_spmv_detail(self):
        """[(rows, entries, systems per launch, ms, algorithmic bytes, timed launches)] of the timed SpMV launches, per matrix.
        Records with systems code 5 / 7 are whole BiCGSTAB calls (class "bicgstab") with 1 / 3 systems."""
        import numpy as np
        cap = 4096
        rows, nnz, sysn = np.zeros(cap, np.int64), np.zeros(cap, np.int64), np.zeros(cap, np.int32)
        ms, by, cnt = np.zeros(cap), np.zeros(cap), np.zeros(cap, np.uint64)
        n = C.c_int32()
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(_lib.lib().orc_prof_get_spmv_detail(self._h, C.c_int32(cap), P(rows), P(nnz), P(sysn), P(ms), P(by), P(cnt), C.byref(n)))
        return [(int(rows[k]), int(nnz[k]), int(sysn[k]), float(ms[k]), float(by[k]), int(cnt[k])) for k in range(n.value)]

    def close(self):
        if self._h:
            _lib.lib().orc_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default = {}


def default_context(device=0):
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]
