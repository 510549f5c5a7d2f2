"""One context per device per process (include/orc_b200.h: orc_ctx_*)."""
import ctypes as C

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

    PROF_CLASSES = ("spmv", "vector", "assembly", "restriction", "galerkin", "scaling", "other")

    def prof_enable(self, on=True):
        _lib.check(_lib.lib().orc_prof_enable(self._h, C.c_int32(1 if on else 0)))

    def prof_get(self):
        """{class: (ms, algorithmic bytes, launches)} accumulated since prof_enable."""
        n = len(self.PROF_CLASSES)
        ms, by, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_uint64 * n)()
        _lib.check(_lib.lib().orc_prof_get(self._h, ms, by, cnt, C.c_int32(n)))
        return {k: (ms[i], by[i], int(cnt[i])) for i, k in enumerate(self.PROF_CLASSES)}

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
