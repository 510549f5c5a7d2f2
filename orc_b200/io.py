"""Mirror of the reference's io::read_mesh (src/io.rs:32-515): TGRID/Fluent ASCII .msh reader + geometry."""
import ctypes as C

from . import _lib
from .mesh import Mesh


def read_mesh(mesh_path):
    out = C.c_void_p()
    _lib.check(_lib.lib().orc_mesh_read(str(mesh_path).encode(), C.byref(out)))
    return Mesh(out)
