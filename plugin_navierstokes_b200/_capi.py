"""ctypes binding of libnsb200.so (include/nsb200.h). No compute happens in Python.

The library is built in-tree by build.py (nvcc, sm_100a). Importing never falls back to a CPU path: if the
shared object is missing or cannot be loaded, LibraryMissing is raised at first use.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnsb200.so")

TRI, QUAD, TET, HEX, PRISM = 0, 1, 2, 3, 4
DISC_FV1, DISC_FVCR = 0, 1
JAC_A, DEF_A, JAC_M, DEF_M, RHS = 1, 2, 4, 8, 16
PHASE_PRIORITY, PHASE_REST = 256, 512
SCATTER_GATHER, SCATTER_COLORED, SCATTER_ATOMIC = 0, 1, 2
HOST, DEVICE, HOST_ASYNC = 0, 1, 2
Q_DEVICE_BYTES, Q_SETUP_SECONDS, Q_FUSED, Q_PATCHES, Q_SCVF_EVALS, Q_PATCH_TABLE_BYTES, Q_LAST_SCATTER = range(7)
OK, ERR_INVALID, ERR_SETUP, ERR_CUDA, ERR_GEOMETRY, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5

SYMBOLS = [
    "nsb_create", "nsb_destroy", "nsb_last_error", "nsb_set_stream", "nsb_params_default", "nsb_set_params",
    "nsb_upload_mesh", "nsb_upload_mesh_fvcr", "nsb_num_dofs", "nsb_nnz", "nsb_num_colors", "nsb_get_csr",
    "nsb_prep_elem_loop", "nsb_assemble", "nsb_local_contributions", "nsb_pack", "nsb_unpack_add",
    "nsb_launch_count", "nsb_synchronize", "nsb_version", "nsb_check_errors", "nsb_query",
    "nsb_assemble_resident", "nsb_resident_jacobian", "nsb_apply_jacobian", "nsb_set_dirichlet", "nsb_adjust_jacobian",
    "nsb_adjust_vector", "nsb_set_ip_data", "nsb_set_boundary_faces", "nsb_assemble_boundary",
    "nsb_turbulent_viscosity", "nsb_diagnostic", "nsb_fvcr_constraint_defect", "nsb_set_priority_nodes",
]
BND_OUTFLOW, BND_INFLOW, BND_TURB_ZERO = 0, 1, 2
TURB_OFF, TURB_SMAGORINSKY = -1, 0
DIAG_VORTICITY, DIAG_KINETIC_ENERGY, DIAG_CFL = 0, 1, 2
IP_KIN_VISC_SCVF, IP_DENSITY_SCVF, IP_DENSITY_SCV, IP_SOURCE_SCVF, IP_SOURCE_SCV = range(5)


class Params(C.Structure):
    _fields_ = [("disc", C.c_int32), ("conv_upwind", C.c_int32), ("stab", C.c_int32), ("stab_upwind", C.c_int32),
                ("diff_length", C.c_int32), ("stokes", C.c_int32), ("laplace", C.c_int32),
                ("peclet_blend", C.c_int32), ("pac_upwind", C.c_int32), ("defect_upwind", C.c_int32),
                ("has_source", C.c_int32), ("kin_visc_set", C.c_int32), ("density_set", C.c_int32),
                ("reserved", C.c_int32), ("exact_jacobian", C.c_double), ("grad_div", C.c_double),
                ("kin_visc", C.c_double), ("density", C.c_double), ("source", C.c_double * 3)]


class TimeSeries(C.Structure):
    _fields_ = [("sol0", C.c_void_p), ("sol1", C.c_void_p), ("dt", C.c_double)]


class LibraryMissing(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing("libnsb200.so is not built (run `python build.py`); there is no CPU fallback")
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as e:
        raise LibraryMissing("cannot load libnsb200.so: %s (there is no CPU fallback)" % e) from e
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.nsb_create.argtypes = [i32, C.POINTER(vp)]
    L.nsb_destroy.argtypes = [vp]
    L.nsb_destroy.restype = None
    L.nsb_last_error.argtypes = [vp]
    L.nsb_last_error.restype = C.c_char_p
    L.nsb_set_stream.argtypes = [vp, vp]
    L.nsb_params_default.argtypes = [C.POINTER(Params)]
    L.nsb_params_default.restype = None
    L.nsb_set_params.argtypes = [vp, C.POINTER(Params)]
    L.nsb_upload_mesh.argtypes = [vp, i32, i64, i64, vp, vp]
    L.nsb_upload_mesh_fvcr.argtypes = [vp, i32, i64, i64, i64, vp, vp, vp]
    L.nsb_num_dofs.argtypes = [vp]
    L.nsb_num_dofs.restype = i64
    L.nsb_nnz.argtypes = [vp]
    L.nsb_nnz.restype = i64
    L.nsb_num_colors.argtypes = [vp]
    L.nsb_get_csr.argtypes = [vp, vp, vp]
    L.nsb_prep_elem_loop.argtypes = [vp]
    L.nsb_assemble.argtypes = [vp, i32, i32, vp, C.POINTER(TimeSeries), C.c_double, C.c_double, C.c_double, vp, vp, i32]
    L.nsb_local_contributions.argtypes = [vp, i32, vp, C.POINTER(TimeSeries), vp, vp, i32]
    L.nsb_pack.argtypes = [vp, i64, vp, vp, vp]
    L.nsb_unpack_add.argtypes = [vp, i64, vp, vp, vp]
    L.nsb_launch_count.argtypes = [vp]
    L.nsb_launch_count.restype = i64
    L.nsb_synchronize.argtypes = [vp]
    L.nsb_check_errors.argtypes = [vp]
    L.nsb_version.restype = C.c_char_p
    L.nsb_query.argtypes = [vp, i32, C.POINTER(C.c_double)]
    L.nsb_assemble_resident.argtypes = [vp, i32, i32, vp, C.POINTER(TimeSeries), C.c_double, C.c_double, C.c_double, vp, i32]
    L.nsb_resident_jacobian.argtypes = [vp, C.POINTER(vp)]
    L.nsb_apply_jacobian.argtypes = [vp, vp, C.c_double, vp, C.c_double, vp, i32]
    L.nsb_set_dirichlet.argtypes = [vp, i64, vp]
    L.nsb_adjust_jacobian.argtypes = [vp, vp]
    L.nsb_adjust_vector.argtypes = [vp, vp, vp, i32]
    L.nsb_set_ip_data.argtypes = [vp, i32, vp, i32]
    L.nsb_set_boundary_faces.argtypes = [vp, i32, i64, vp, vp, vp]
    L.nsb_assemble_boundary.argtypes = [vp, i32, vp, C.c_double, vp, vp, i32]
    L.nsb_turbulent_viscosity.argtypes = [vp, i32, C.c_double, vp, i64, vp, vp, i32]
    L.nsb_diagnostic.argtypes = [vp, i32, vp, C.c_double, vp, i32]
    L.nsb_set_priority_nodes.argtypes = [vp, i64, vp]
    L.nsb_fvcr_constraint_defect.argtypes = [vp, vp, C.c_double, i32, i32, i64, vp, vp, i32]
    _lib = L
    return L
