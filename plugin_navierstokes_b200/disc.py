"""Host-side mirror of the reference's element-disc classes for the assembly path.

Same names, setters, defaults and error behaviour as the UG4 plugin's Lua-visible classes (SURVEY.md
App. D): NavierStokesFV1 / NavierStokesFVCR (+ the `NavierStokes(fcts, subsets, discType)` factory of
lua/lua-include.lua:36-47), upwind names of upwind_interface.cpp:43-62, stabilisation names of
fv1/stabilization.cpp:46-57,86-100.  Everything numerical happens in libnsb200.so (CUDA); this module
only keeps the disc state and passes pointers.  ugcore's Domain / ApproximationSpace are replaced by
`set_grid(elem, conn, coords)`; `assemble_jacobian` / `assemble_defect` play the role of ugcore's
DomainDiscretization::assemble_* restricted to this disc.
"""
import ctypes as C

import numpy as np

from . import _capi as capi

_UPWIND_NAMES = {  # upwind_interface.cpp:46-61 (trimmed, case-insensitive)
    "no": 1, "full": 2, "skewed": 3, "linearprofileskewed": 4, "lps": 4, "positive": 5, "pos": 5,
}
_STAB_NAMES = {"fields": 0, "flow": 1}                      # stabilization.cpp:52-53
_DIFF_NAMES = {"raw": 0, "fivepoint": 1, "cor": 2}          # stabilization.cpp:94-96
_ELEMS = {"tri": capi.TRI, "quad": capi.QUAD, "tet": capi.TET, "hex": capi.HEX, "prism": capi.PRISM,
          "triangle": capi.TRI, "quadrilateral": capi.QUAD, "tetrahedron": capi.TET, "hexahedron": capi.HEX}
_NSH = {capi.TRI: 3, capi.QUAD: 4, capi.TET: 4, capi.HEX: 8, capi.PRISM: 6}
_DIM = {capi.TRI: 2, capi.QUAD: 2, capi.TET: 3, capi.HEX: 3, capi.PRISM: 3}
_NSIDE = {capi.TRI: 3, capi.QUAD: 4, capi.TET: 4, capi.HEX: 6, capi.PRISM: 5}


class UGError(RuntimeError):
    """the counterpart of UG_THROW on this path"""


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class INavierStokesUpwind:
    """marker objects mirroring the registered upwind classes (register_navier_stokes.cpp:156-228)"""
    _id = 0

    def __eq__(self, other):
        return isinstance(other, INavierStokesUpwind) and other._id == self._id


class NavierStokesNoUpwind(INavierStokesUpwind):
    _id = 1


class NavierStokesFullUpwind(INavierStokesUpwind):
    _id = 2


class NavierStokesSkewedUpwind(INavierStokesUpwind):
    _id = 3


class NavierStokesLinearProfileSkewedUpwind(INavierStokesUpwind):
    _id = 4


class NavierStokesPositiveUpwind(INavierStokesUpwind):
    _id = 5


class NavierStokesRegularUpwind(INavierStokesUpwind):
    """2-D only in the reference (upwind.cpp:806); not provided on the device path."""
    _id = 6


def CreateNavierStokesUpwind(name):
    n = name.strip().lower()
    if n not in _UPWIND_NAMES and n not in ("regular", "reg"):
        raise UGError("NavierStokes: upwind type '%s' not found. Options are: no, full, skewed, "
                      "linearprofileskewed (lps), positive (pos), regular (reg)" % name)
    if n in ("regular", "reg"):
        return NavierStokesRegularUpwind()
    cls = {1: NavierStokesNoUpwind, 2: NavierStokesFullUpwind, 3: NavierStokesSkewedUpwind,
           4: NavierStokesLinearProfileSkewedUpwind, 5: NavierStokesPositiveUpwind}[_UPWIND_NAMES[n]]
    return cls()


class INavierStokesFV1Stabilization:
    _id = -1

    def __init__(self):
        self._upwind = None
        self._diff = 0          # RAW, stabilization.h:324

    def set_upwind(self, upwind):
        self._upwind = upwind

    def upwind(self):
        return self._upwind


class INavierStokesSRFV1Stabilization(INavierStokesFV1Stabilization):
    def set_diffusion_length(self, name):
        n = name.strip().lower()
        if n not in _DIFF_NAMES:
            raise UGError("Diffusion Length calculation method not found. Use one of [Raw, Fivepoint, Cor].")
        self._diff = _DIFF_NAMES[n]


class NavierStokesFIELDSStabilization(INavierStokesSRFV1Stabilization):
    _id = 0


class NavierStokesFLOWStabilization(INavierStokesSRFV1Stabilization):
    _id = 1


class NavierStokesFV1WithoutStabilization(INavierStokesFV1Stabilization):
    _id = 2


def CreateNavierStokesStabilization(name):
    n = name.strip().lower()
    if n == "fields":
        return NavierStokesFIELDSStabilization()
    if n == "flow":
        return NavierStokesFLOWStabilization()
    raise UGError("NavierStokes: stabilization type '%s' not a valid name of a Schneider-Raw stabilization."
                  " Options are: fields, flow" % name)


class _DeviceDisc:
    """shared machinery: context ownership, grid upload, the element loop on the device"""
    _disc = capi.DISC_FV1

    def __init__(self, fcts, subsets="", device=0):
        if isinstance(fcts, str):
            fcts = [f.strip() for f in fcts.split(",") if f.strip()]
        self._fcts = list(fcts)
        self._subsets = subsets
        self._device = device
        self._ctx = None
        self._elem = None
        self._time_dependent = False
        # navier_stokes_base.cpp:53-67, incompressible_navier_stokes_base.cpp:53-66
        self._exact_jac = 0.0
        self._stokes = self._laplace = self._peclet = False
        self._visc = None
        self._density = 1.0
        self._source = None
        self._grad_div = 0.0
        self._conv_upwind = None
        self.scatter_mode = capi.SCATTER_GATHER

    # ---- NavierStokesBase (register_navier_stokes.cpp:105-126) ----
    # UserData imports: a number (constant), or a callable f(x) -> value evaluated by the HOST at the integration points the
    # reference evaluates its imports at (fv1/navier_stokes_fv1.cpp:184-197) and handed to the device as per-ip arrays
    # (nsb_set_ip_data); Lua callback NAMES (the const char* overloads) cannot be resolved outside a Lua state.
    def set_kinematic_viscosity(self, v):
        if isinstance(v, str):
            raise UGError("device path: Lua callback names cannot be resolved (pass a number or a Python callable)")
        self._ip_fn = getattr(self, "_ip_fn", {})
        if callable(v):
            self._ip_fn["visc"] = v
            self._visc = float("nan") if self._visc is None else self._visc      # "data given"
        else:
            self._ip_fn.pop("visc", None)
            self._visc = float(v)
        self._ip_dirty = True

    def set_source(self, v):
        if isinstance(v, str):
            raise UGError("device path: Lua callback names cannot be resolved (pass a vector or a Python callable)")
        self._ip_fn = getattr(self, "_ip_fn", {})
        if callable(v):
            self._ip_fn["source"] = v
            self._source = None
        else:
            self._ip_fn.pop("source", None)
            self._source = [float(x) for x in v]
        self._ip_dirty = True

    def set_exact_jacobian(self, v):
        # bool overload -> 1.0/0.0, number overload -> factor (navier_stokes_base.h)
        self._exact_jac = float(v)

    # ---- IncompressibleNavierStokesBase (incompressible_navier_stokes_plugin.cpp:244-266) ----
    def set_density(self, v):
        if isinstance(v, str):
            raise UGError("device path: Lua callback names cannot be resolved (pass a number or a Python callable)")
        self._ip_fn = getattr(self, "_ip_fn", {})
        if callable(v):
            self._ip_fn["density"] = v
        else:
            self._ip_fn.pop("density", None)
            self._density = float(v)
        self._ip_dirty = True

    def set_peclet_blend(self, b):
        self._peclet = bool(b)

    def set_grad_div(self, f):
        self._grad_div = float(f)

    def set_laplace(self, b):
        self._laplace = bool(b)

    def set_stokes(self, b):
        self._stokes = bool(b)

    def requests_local_time_series(self):
        return True                                   # navier_stokes_base.h:200

    # ---- grid / context ----
    def _context(self):
        if self._ctx is None:
            L = capi.lib()
            ctx = C.c_void_p()
            rc = L.nsb_create(self._device, C.byref(ctx))
            if rc != 0:
                raise UGError(L.nsb_last_error(None).decode())
            self._ctx = ctx
        return self._ctx

    def _check(self, rc):
        if rc != 0:
            raise UGError(capi.lib().nsb_last_error(self._ctx).decode())

    def __del__(self):
        try:
            if self._ctx is not None:
                capi.lib().nsb_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    close = __del__

    def _num_fct_check(self, dim):
        if len(self._fcts) != dim + 1:           # fv1/navier_stokes_fv1.cpp:66, fvcr/navier_stokes_fvcr.cpp:67-68
            raise UGError("Wrong number of functions: The ElemDisc 'NavierStokes' needs exactly %d symbolic function." % (dim + 1))

    @property
    def num_dofs(self):
        return capi.lib().nsb_num_dofs(self._ctx)

    @property
    def nnz(self):
        return capi.lib().nsb_nnz(self._ctx)

    @property
    def num_colors(self):
        return capi.lib().nsb_num_colors(self._ctx)

    @property
    def launch_count(self):
        return capi.lib().nsb_launch_count(self._ctx)

    def query(self, what):
        """instrumentation (capi.Q_*): device bytes, set-up seconds, fused-kernel statistics"""
        out = C.c_double(0.0)
        self._check(capi.lib().nsb_query(self._ctx, int(what), C.byref(out)))
        return out.value

    def csr(self):
        """(rowptr int64 [ndof+1], colind int32 [nnz]) of the global Jacobian (sorted rows)"""
        rowptr = np.empty(self.num_dofs + 1, dtype=np.int64)
        colind = np.empty(self.nnz, dtype=np.int32)
        self._check(capi.lib().nsb_get_csr(self._ctx, rowptr.ctypes.data, colind.ctypes.data))
        return rowptr, colind

    def use_stream(self, cuda_stream):
        self._check(capi.lib().nsb_set_stream(self._context(), C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(capi.lib().nsb_synchronize(self._ctx))
        self._async_keep = []                                   # buffers of asynchronous calls are complete now

    def check_errors(self):
        self._check(capi.lib().nsb_check_errors(self._ctx))

    def _params(self):
        raise NotImplementedError

    def prep_elem_loop(self):
        """validation of prep_elem_loop (throws like the reference)"""
        p = self._params()
        self._check(capi.lib().nsb_set_params(self._context(), C.byref(p)))
        self._check(capi.lib().nsb_prep_elem_loop(self._ctx))

    # ---- the element loop ----
    @staticmethod
    def _ptr(a):
        if a is None:
            return None
        if _is_torch(a):
            return C.c_void_p(a.data_ptr())
        return C.c_void_p(a.ctypes.data)

    def assemble(self, what, u, values=None, defect=None, time_series=None, scale_a=1.0, scale_m=1.0, beta=0.0,
                 scatter_mode=None):
        """values/defect := beta*old + scale_a*A-part + scale_m*M-part (see nsb_assemble).
        u (and the optional (sol0, sol1, dt) time series) are numpy arrays (host path, copies inside the
        call) or torch CUDA tensors (device path, asynchronous). Returns (values, defect)."""
        L = capi.lib()
        p = self._params()
        self._check(L.nsb_set_params(self._context(), C.byref(p)))
        on_dev = _is_torch(u)
        jac = bool(what & (capi.JAC_A | capi.JAC_M))
        dfc = bool(what & (capi.DEF_A | capi.DEF_M | capi.RHS))
        if on_dev:
            import torch
            assert u.is_cuda and u.dtype == torch.float64 and u.is_contiguous()
            # stream contract: device-pointer calls run on torch's CURRENT stream of u's device, so they are
            # ordered with the producers of u and with the consumers of the returned tensors like any torch op
            self.use_stream(torch.cuda.current_stream(u.device).cuda_stream)
            # a missing output is allocated zero-filled and the caller's beta is kept (same as the host path)
            if jac and values is None:
                values = torch.zeros(self.nnz, dtype=torch.float64, device=u.device)
            if dfc and defect is None:
                defect = torch.zeros(self.num_dofs, dtype=torch.float64, device=u.device)
            for name, t, n in (("values", values if jac else None, self.nnz), ("defect", defect if dfc else None, self.num_dofs)):
                if t is not None and not (_is_torch(t) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() == n):
                    raise UGError("assemble: %s must be a contiguous float64 CUDA tensor with %d entries" % (name, n))
        else:
            u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
            if u.shape[0] != self.num_dofs:
                raise UGError("assemble: u has %d entries, the grid has %d dofs" % (u.shape[0], self.num_dofs))
            if jac and values is None:
                values = np.zeros(self.nnz)
            if dfc and defect is None:
                defect = np.zeros(self.num_dofs)
            for name, a, n in (("values", values if jac else None, self.nnz), ("defect", defect if dfc else None, self.num_dofs)):
                if a is not None and not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size == n):
                    raise UGError("assemble: %s must be a C-contiguous float64 numpy array with %d entries" % (name, n))
        ts = None
        keep = []
        if time_series is not None:
            s0, s1, dt = time_series
            if not on_dev:
                s0 = np.ascontiguousarray(s0, dtype=np.float64).reshape(-1)
                s1 = np.ascontiguousarray(s1, dtype=np.float64).reshape(-1)
            keep = [s0, s1]
            ts = capi.TimeSeries(self._ptr(s0), self._ptr(s1), float(dt))
        mode = self.scatter_mode if scatter_mode is None else scatter_mode
        rc = L.nsb_assemble(self._ctx, what, mode, self._ptr(u), C.byref(ts) if ts is not None else None,
                            float(scale_a), float(scale_m), float(beta), self._ptr(values) if jac else None,
                            self._ptr(defect) if dfc else None, capi.DEVICE if on_dev else capi.HOST)
        del keep
        self._check(rc)
        return values, defect

    # ---- GPU-resident Jacobian (nsb_assemble_resident / nsb_apply_jacobian) and Dirichlet post-pass ----
    def assemble_resident(self, what, u, defect=None, time_series=None, scale_a=1.0, scale_m=1.0, beta=0.0, scatter_mode=None,
                          asynchronous=False):
        """like assemble(), but the CSR values stay in a context-owned device buffer: only u (and the time series) go to the
        device and the defect comes back. Returns the defect.
        asynchronous=True (host arrays, NSB_HOST_ASYNC): the call returns once the work is queued; the copies run on the context's
        copy streams. Pass C-contiguous float64 arrays (pinned for real overlap); they are complete after synchronize()."""
        L = capi.lib()
        p = self._params()
        self._check(L.nsb_set_params(self._context(), C.byref(p)))
        on_dev = _is_torch(u)
        dfc = bool(what & (capi.DEF_A | capi.DEF_M | capi.RHS))
        if on_dev:
            import torch
            self.use_stream(torch.cuda.current_stream(u.device).cuda_stream)
            if dfc and defect is None:
                defect = torch.zeros(self.num_dofs, dtype=torch.float64, device=u.device)
        else:
            if asynchronous:
                for a in (u, defect) + (tuple(time_series[:2]) if time_series is not None else ()):
                    if a is not None and not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]):
                        raise UGError("assemble_resident(asynchronous=True): pass C-contiguous float64 numpy arrays")
                if dfc and defect is None:
                    raise UGError("assemble_resident(asynchronous=True): pass the defect array")
                self._async_keep = getattr(self, "_async_keep", []) + [u, defect]
            u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
            if dfc and defect is None:
                defect = np.zeros(self.num_dofs)
        ts = None
        keep = []
        if time_series is not None:
            s0, s1, dt = time_series
            if not on_dev:
                s0 = np.ascontiguousarray(s0, dtype=np.float64).reshape(-1)
                s1 = np.ascontiguousarray(s1, dtype=np.float64).reshape(-1)
            keep = [s0, s1]
            ts = capi.TimeSeries(self._ptr(s0), self._ptr(s1), float(dt))
        mode = self.scatter_mode if scatter_mode is None else scatter_mode
        rc = L.nsb_assemble_resident(self._ctx, what, mode, self._ptr(u), C.byref(ts) if ts is not None else None,
                                     float(scale_a), float(scale_m), float(beta), self._ptr(defect) if dfc else None,
                                     capi.DEVICE if on_dev else (capi.HOST_ASYNC if asynchronous else capi.HOST))
        if asynchronous and not on_dev:
            self._async_keep += keep
        del keep
        self._check(rc)
        return defect

    def resident_jacobian_ptr(self):
        """device pointer (int) of the resident CSR values, pattern = csr()"""
        out = C.c_void_p()
        self._check(capi.lib().nsb_resident_jacobian(self._ctx, C.byref(out)))
        return out.value

    def apply_jacobian(self, x, y=None, alpha=1.0, beta=0.0, values=None, asynchronous=False):
        """y = alpha * J x + beta * y with the resident Jacobian (values=None) or a CUDA tensor of CSR values.
        asynchronous=True (host arrays, NSB_HOST_ASYNC): x goes up on the H2D stream while the context stream is still busy,
        y is complete after synchronize()."""
        on_dev = _is_torch(x)
        if on_dev:
            import torch
            self.use_stream(torch.cuda.current_stream(x.device).cuda_stream)
            if y is None:
                y = torch.zeros(self.num_dofs, dtype=torch.float64, device=x.device)
        else:
            if asynchronous:
                for a in (x, y):
                    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]):
                        raise UGError("apply_jacobian(asynchronous=True): pass C-contiguous float64 numpy arrays for x and y")
                self._async_keep = getattr(self, "_async_keep", []) + [x, y]
            x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
            if y is None:
                y = np.zeros(self.num_dofs)
        self._check(capi.lib().nsb_apply_jacobian(self._ctx, self._ptr(values), float(alpha), self._ptr(x), float(beta), self._ptr(y),
                                                  capi.DEVICE if on_dev else (capi.HOST_ASYNC if asynchronous else capi.HOST)))
        return y

    # ---- per-ip data imports ----
    _IP_KINDS = {"visc": capi.IP_KIN_VISC_SCVF, "rho_scvf": capi.IP_DENSITY_SCVF, "rho_scv": capi.IP_DENSITY_SCV,
                 "src_scvf": capi.IP_SOURCE_SCVF, "src_scv": capi.IP_SOURCE_SCV}

    def set_ip_data(self, kind, data):
        """per-ip array of one import (kind: visc | rho_scvf | rho_scv | src_scvf | src_scv; layouts of nsb_set_ip_data), None clears it"""
        k = self._IP_KINDS[kind]
        if data is None:
            self._check(capi.lib().nsb_set_ip_data(self._context(), k, None, capi.HOST))
            return
        if _is_torch(data):
            self._check(capi.lib().nsb_set_ip_data(self._context(), k, C.c_void_p(data.data_ptr()), capi.DEVICE))
        else:
            a = np.ascontiguousarray(data, dtype=np.float64)
            self._check(capi.lib().nsb_set_ip_data(self._context(), k, a.ctypes.data, capi.HOST))

    def _push_ip_data(self):
        """evaluates callable UserData at the integration points of the uploaded grid (once per change)"""
        fns = getattr(self, "_ip_fn", {})
        if not getattr(self, "_ip_dirty", False) or getattr(self, "_grid_host", None) is None:
            return
        from . import meshgen
        elem, conn, coords = self._grid_host
        name = ("tri", "quad", "tet", "hex", "prism")[elem]
        xf = meshgen.fv1_scvf_ips(name, conn, coords)
        xv = meshgen.fv1_scv_ips(name, conn, coords)

        def ev(fn, x, ncomp):
            flat = x.reshape(-1, x.shape[-1])
            out = np.array([np.atleast_1d(fn(*pt)) for pt in flat], dtype=np.float64)
            return out.reshape(x.shape[:-1] + ((ncomp,) if ncomp > 1 else ()))

        dim = coords.shape[1]
        self.set_ip_data("visc", ev(fns["visc"], xf, 1) if "visc" in fns else None)
        self.set_ip_data("rho_scvf", ev(fns["density"], xf, 1) if "density" in fns else None)
        self.set_ip_data("rho_scv", ev(fns["density"], xv, 1) if "density" in fns else None)
        self.set_ip_data("src_scvf", ev(fns["source"], xf, dim) if "source" in fns else None)
        self.set_ip_data("src_scv", ev(fns["source"], xv, dim) if "source" in fns else None)
        self._ip_dirty = False

    def set_priority_nodes(self, nodes):
        """grid nodes whose rows are assembled first (nsb_set_priority_nodes); assemble(what | capi.PHASE_PRIORITY, ...) then
        assemble(what | capi.PHASE_REST, ...) split the pass behind them (device tensors only). Empty list clears."""
        nodes = np.ascontiguousarray(nodes, dtype=np.int64).reshape(-1)
        self._check(capi.lib().nsb_set_priority_nodes(self._context(), nodes.size, self._ptr(nodes) if nodes.size else None))

    def set_dirichlet(self, dofs):
        dofs = np.ascontiguousarray(dofs, dtype=np.int64).reshape(-1)
        self._check(capi.lib().nsb_set_dirichlet(self._context(), dofs.size, dofs.ctypes.data))
        self._n_dirichlet = int(dofs.size)

    def adjust_jacobian(self, values=None):
        """Dirichlet rows := unit rows, in the resident Jacobian (values=None) or a CUDA tensor of CSR values"""
        self._check(capi.lib().nsb_adjust_jacobian(self._ctx, self._ptr(values)))

    def adjust_vector(self, vec, g=None):
        """vec[dirichlet dofs] := g (host array, one value per Dirichlet dof) or 0: adjust_solution / adjust_defect"""
        if g is not None:
            g = np.ascontiguousarray(g, dtype=np.float64).reshape(-1)
            if g.size != getattr(self, "_n_dirichlet", -1):
                raise UGError("adjust_vector: one value per Dirichlet dof expected")
        self._check(capi.lib().nsb_adjust_vector(self._ctx, self._ptr(vec), self._ptr(g), capi.DEVICE if _is_torch(vec) else capi.HOST))
        return vec

    # ---- boundary element discs on the FV1 boundary faces (nsb_set_boundary_faces / nsb_assemble_boundary) ----
    def set_boundary_faces(self, kind, elems, sides, data=None):
        elems = np.ascontiguousarray(elems, dtype=np.int32).reshape(-1)
        sides = np.ascontiguousarray(sides, dtype=np.int32).reshape(-1)
        if elems.size != sides.size:
            raise UGError("set_boundary_faces: one local side index per element expected")
        if data is not None:
            data = np.ascontiguousarray(data, dtype=np.float64)
            if data.size != elems.size * 4 * (len(self._fcts) - 1):
                raise UGError("set_boundary_faces: data must be [n_side][4][dim]")
        self._check(capi.lib().nsb_set_boundary_faces(self._context(), kind, elems.size, self._ptr(elems), self._ptr(sides), self._ptr(data)))

    def assemble_boundary(self, what, u, values=None, defect=None, scale_a=1.0):
        """ADDS the contributions of the registered boundary discs (outflow, inflow continuity term) to the resident Jacobian
        (values=None) or a CUDA tensor of CSR values, and to `defect`; call after assemble*/assemble_resident and before the
        Dirichlet post-pass. Returns defect."""
        L = capi.lib()
        p = self._params()
        self._check(L.nsb_set_params(self._context(), C.byref(p)))
        on_dev = _is_torch(u)
        dfc = bool(what & (capi.DEF_A | capi.RHS))
        if dfc and defect is None:
            raise UGError("assemble_boundary: the defect to add to is missing")
        if on_dev:
            import torch
            self.use_stream(torch.cuda.current_stream(u.device).cuda_stream)
        else:
            u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
            if dfc and (defect.dtype != np.float64 or not defect.flags.c_contiguous):
                raise UGError("assemble_boundary: defect must be a contiguous float64 array")
        self._check(L.nsb_assemble_boundary(self._ctx, what, self._ptr(u), float(scale_a), self._ptr(values), self._ptr(defect) if dfc else None,
                                            capi.DEVICE if on_dev else capi.HOST))
        return defect

    # ---- diagnostics of navier_stokes_tools.h on the device (nsb_diagnostic) ----
    def _diagnostic(self, kind, u, dt, n_out):
        on_dev = _is_torch(u)
        if on_dev:
            import torch
            self.use_stream(torch.cuda.current_stream(u.device).cuda_stream)
            out = torch.zeros(n_out, dtype=torch.float64, device=u.device)
        else:
            u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
            out = np.zeros(n_out)
        self._context()
        self._check(capi.lib().nsb_diagnostic(self._ctx, kind, self._ptr(u), float(dt), self._ptr(out), capi.DEVICE if on_dev else capi.HOST))
        return out

    def vorticity(self, u):
        """vorticityFV1 (navier_stokes_tools.h:386-525): d_x v - d_y u per vertex (FV1 grids)"""
        return self._diagnostic(capi.DIAG_VORTICITY, u, 0.0, self.num_dofs // len(self._fcts))

    def kinetic_energy(self, u):
        """kineticEnergy (navier_stokes_tools.h:850-965), Crouzeix-Raviart velocity (FVCR grids)"""
        r = self._diagnostic(capi.DIAG_KINETIC_ENERGY, u, 0.0, 1)
        return float(r[0])

    def cfl_number(self, u, dt):
        """cflNumber (navier_stokes_tools.h:731-848), Crouzeix-Raviart velocity (FVCR grids)"""
        r = self._diagnostic(capi.DIAG_CFL, u, dt, 1)
        return float(r[0])

    def assemble_jacobian(self, u, **kw):
        return self.assemble(capi.JAC_A, u, **kw)[0]

    def assemble_defect(self, u, **kw):
        return self.assemble(capi.DEF_A | capi.RHS, u, **kw)[1]

    def assemble_jacobian_defect(self, u, **kw):
        return self.assemble(capi.JAC_A | capi.DEF_A | capi.RHS, u, **kw)


class NavierStokesFV1(_DeviceDisc):
    """fv1/navier_stokes_fv1.h -- registered at fv1/register_fv1.cpp:166-184"""
    _disc = capi.DISC_FV1

    def __init__(self, fcts, subsets="", device=0):
        super().__init__(fcts, subsets, device)
        self._stab = None
        self._conv_stab = None

    def disc_type(self):
        return "fv1"

    def use_hanging(self):
        return False

    # fv1/navier_stokes_fv1.h:185-225
    def set_stabilization(self, stab, diff_length=None):
        if isinstance(stab, str):
            s = CreateNavierStokesStabilization(stab)
            if diff_length is not None:
                s.set_diffusion_length(diff_length)
            self._stab = s
            if self._conv_upwind is not None:
                self._stab.set_upwind(self._conv_upwind)
        else:
            self._stab = stab

    def stabilization(self):
        return self._stab

    def set_upwind(self, up):
        if isinstance(up, str):
            self._conv_stab = None
            self._conv_upwind = CreateNavierStokesUpwind(up)
            if self._stab is not None and self._stab.upwind() is None:
                self._stab.set_upwind(self._conv_upwind)
        elif isinstance(up, INavierStokesFV1Stabilization):
            self._conv_stab = up
            self._conv_upwind = None
        else:
            self._conv_stab = None
            self._conv_upwind = up

    def set_pac_upwind(self, b):
        if b:
            if self._conv_upwind is None:
                raise UGError("Upwind must be specified previously.\n")
            if self._stab is None:
                raise UGError("Stabilization must be specified previously.\n")
            self._stab.set_upwind(self._conv_upwind)
            self.set_upwind(self._stab)

    def set_grid(self, elem, conn, coords):
        """upload grid connectivity + coordinates (replaces the Domain / DoFDistribution of ugcore)"""
        e = _ELEMS[elem] if isinstance(elem, str) else int(elem)
        self._num_fct_check(_DIM[e])
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        assert conn.shape[1] == _NSH[e] and coords.shape[1] == _DIM[e]
        self._elem = e
        self._check(capi.lib().nsb_upload_mesh(self._context(), e, conn.shape[0], coords.shape[0],
                                               conn.ctypes.data, coords.ctypes.data))
        self._n_elem = conn.shape[0]
        self._grid_host = (e, conn, coords) if getattr(self, "_ip_fn", None) else (e, conn, coords)
        self._ip_dirty = bool(getattr(self, "_ip_fn", {}))

    def _params(self):
        self._push_ip_data()
        p = capi.Params()
        capi.lib().nsb_params_default(C.byref(p))
        p.disc = capi.DISC_FV1
        pac = self._conv_stab is not None
        if pac and self._conv_stab is not self._stab:
            raise UGError("device path: a convective stabilisation different from the continuity stabilisation is not supported")
        cu = self._stab.upwind() if (pac and self._stab is not None) else self._conv_upwind
        if cu is not None and cu._id == 6:
            raise UGError("device path: RegularUpwind is not provided")
        p.conv_upwind = cu._id if cu is not None else 0
        p.pac_upwind = int(pac)
        if self._stab is not None:
            p.stab = self._stab._id
            su = self._stab.upwind()
            p.stab_upwind = su._id if su is not None else 0
            p.diff_length = self._stab._diff
        p.stokes, p.laplace, p.peclet_blend = int(self._stokes), int(self._laplace), int(self._peclet)
        p.exact_jacobian = self._exact_jac
        p.kin_visc_set = int(self._visc is not None)
        p.kin_visc = self._visc if self._visc is not None else 0.0
        p.density_set = int(self._density is not None)
        p.density = self._density if self._density is not None else 0.0
        if self._source is not None:
            p.has_source = 1
            for d, v in enumerate(self._source[:3]):
                p.source[d] = v
        return p

    def local_contributions(self, what, u, time_series=None):
        """compat mode: per-element LocalMatrix / LocalVector blocks [n_elem, L, L], [n_elem, L]"""
        L = capi.lib()
        p = self._params()
        self._check(L.nsb_set_params(self._context(), C.byref(p)))
        nsh, nf = _NSH[self._elem], _DIM[self._elem] + 1
        Ls = nsh * nf
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
        J = np.zeros((self._n_elem, Ls, Ls))
        d = np.zeros((self._n_elem, Ls))
        ts = None
        if time_series is not None:
            s0 = np.ascontiguousarray(time_series[0], dtype=np.float64).reshape(-1)
            s1 = np.ascontiguousarray(time_series[1], dtype=np.float64).reshape(-1)
            ts = capi.TimeSeries(self._ptr(s0), self._ptr(s1), float(time_series[2]))
        self._check(L.nsb_local_contributions(self._ctx, what, self._ptr(u), C.byref(ts) if ts is not None else None,
                                              self._ptr(J), self._ptr(d), capi.HOST))
        return J, d


class NavierStokesFVCR(_DeviceDisc):
    """fvcr/navier_stokes_fvcr.h -- registered at fvcr/register_fvcr.cpp:289-303"""
    _disc = capi.DISC_FVCR

    def __init__(self, fcts, subsets="", device=0):
        super().__init__(fcts, subsets, device)
        self._defect_upwind = True                     # fvcr/navier_stokes_fvcr.cpp:82
        # GATHER = the deterministic default: one launch of order-free reductions for beta == 0 (a CR entry has at most two
        # contributions), coloured sweeps otherwise (nsb200.cu: assemble_fvcr)
        self.scatter_mode = capi.SCATTER_GATHER

    def disc_type(self):
        return "fvcr"

    def use_hanging(self):
        return True                                    # fvcr/navier_stokes_fvcr.cpp:111-116

    def set_upwind(self, up):
        self._conv_upwind = CreateNavierStokesUpwind(up) if isinstance(up, str) else up

    def set_defect_upwind(self, b):
        self._defect_upwind = bool(b)

    def set_grid(self, elem, conn, coords, elem_sides=None, n_side=None):
        from . import meshgen
        e = _ELEMS[elem] if isinstance(elem, str) else int(elem)
        self._num_fct_check(_DIM[e])
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        if elem_sides is None:
            name = {capi.TRI: "tri", capi.QUAD: "quad", capi.TET: "tet", capi.HEX: "hex", capi.PRISM: "prism"}[e]
            elem_sides, n_side = meshgen.element_sides(name, conn)
        es = np.ascontiguousarray(elem_sides, dtype=np.int32)
        self._elem = e
        self._check(capi.lib().nsb_upload_mesh_fvcr(self._context(), e, conn.shape[0], coords.shape[0], int(n_side),
                                                    conn.ctypes.data, es.ctypes.data, coords.ctypes.data))
        self._n_elem = conn.shape[0]
        self.elem_sides, self.n_side = es, int(n_side)

    def _params(self):
        p = capi.Params()
        capi.lib().nsb_params_default(C.byref(p))
        p.disc = capi.DISC_FVCR
        cu = self._conv_upwind
        if cu is not None and cu._id == 6:
            raise UGError("device path: RegularUpwind is not provided")
        p.conv_upwind = cu._id if cu is not None else 0
        p.defect_upwind = int(self._defect_upwind)
        p.stokes, p.laplace, p.peclet_blend = int(self._stokes), int(self._laplace), int(self._peclet)
        p.exact_jacobian = self._exact_jac
        p.grad_div = self._grad_div
        p.kin_visc_set = int(self._visc is not None)
        p.kin_visc = self._visc if self._visc is not None else 0.0
        p.density_set = 1
        p.density = self._density
        if self._source is not None:
            p.has_source = 1
            for d, v in enumerate(self._source[:3]):
                p.source[d] = v
        return p


def NavierStokes(fcts, subsets, disc_type=None, device=0):
    """lua/lua-include.lua:36-47"""
    if disc_type is None or disc_type == "fv1":
        return NavierStokesFV1(fcts, subsets, device)
    if disc_type == "fvcr":
        return NavierStokesFVCR(fcts, subsets, device)
    if disc_type in ("fv", "fe", "fecr"):
        raise UGError("NavierStokes: disc type '%s' is outside the device assembly path (fv1, fvcr)" % disc_type)
    raise UGError("NavierStokes: no disc type '%s' available. Use 'fv1', 'fv', 'fvcr', 'fe' or 'fecr'." % disc_type)


# ------------------------------------------------------------------------------------------------
# boundary conditions that are Dirichlet constraints on the velocity
# ------------------------------------------------------------------------------------------------
class _DirichletVelocity:
    """collects (node, value) pairs of the velocity components and applies ugcore's DirichletBoundary post-pass
    (adjust_jacobian / adjust_defect / adjust_solution) through the C ABI (nsb_set_dirichlet, nsb_adjust_*)."""

    def __init__(self, master):
        fcts = master.symb_fcts() if hasattr(master, "symb_fcts") else master._fcts
        self._master = master
        self._dim = len(fcts) - 1
        if len(fcts) != self._dim + 1 or self._dim not in (2, 3):
            raise UGError("This Boundary Condition works on exactly dim+1 (velocity+pressure) components, but %d components given." % len(fcts))
        self._nodes, self._vals = [], []

    def _add(self, nodes, values):
        nodes = np.asarray(nodes, dtype=np.int64).reshape(-1)
        values = np.broadcast_to(np.asarray(values, dtype=np.float64), (nodes.size, self._dim))
        self._nodes.append(nodes)
        self._vals.append(np.array(values))

    def dirichlet(self):
        """(dofs int64, values float64) of all constrained velocity dofs (FV1 numbering node*(dim+1)+d)"""
        if not self._nodes:
            return np.zeros(0, dtype=np.int64), np.zeros(0)
        nodes = np.concatenate(self._nodes)
        vals = np.concatenate(self._vals, axis=0)
        nf = self._dim + 1
        dofs = (nodes[:, None] * nf + np.arange(self._dim)[None, :]).reshape(-1)
        v = vals.reshape(-1)
        dofs, first = np.unique(dofs, return_index=True)     # a node in two boundary subsets: first registration wins
        return dofs, v[first]

    def apply(self, disc=None):
        """register the constraint with the device context of `disc` (default: the master disc)"""
        d = disc if disc is not None else self._master
        dofs, vals = self.dirichlet()
        d.set_dirichlet(dofs)
        self._applied_vals = vals
        return dofs, vals


class NavierStokesWall(_DirichletVelocity):
    """bnd/wall_impl.h:44-70: velocity = 0 on the wall subsets (registered at register_navier_stokes.cpp)"""

    def add(self, boundary_nodes):
        self._add(boundary_nodes, 0.0)


class NavierStokesInflowFV1(_DirichletVelocity):
    """fv1/bnd/inflow_fv1_impl.h:42-82: velocity = user data on the inflow subsets. Two parts, as in the reference (:69, :77):
    the Dirichlet rows of the velocity (nsb_set_dirichlet) and, when the boundary SIDES are given, the NeumannBoundaryFV1 term of
    the continuity equation over their boundary faces (nsb_set_boundary_faces(NSB_BND_INFLOW): defect(p) += user . n)."""

    def __init__(self, master):
        super().__init__(master)
        self._be, self._bs, self._bdata = [], [], []

    def add(self, user, boundary_nodes, coords=None, sides=None, conn=None, elem=None):
        """sides = (elements, local sides) of the inflow boundary (e.g. meshgen.boundary_sides) adds the continuity term; conn / elem
        (and coords) are then needed to locate the boundary-face ips where `user` is evaluated"""
        if callable(user):
            if coords is None:
                raise UGError("NavierStokesInflow::add: coordinates needed to evaluate the user data")
            vals = np.array([user(*coords[n]) for n in np.asarray(boundary_nodes).reshape(-1)], dtype=np.float64)
        else:
            vals = user
        self._add(boundary_nodes, vals)
        if sides is not None:
            from . import meshgen
            be, bs = sides
            if conn is None or elem is None or coords is None:
                raise UGError("NavierStokesInflow::add: conn, elem and coords needed for the boundary faces")
            xip = meshgen.fv1_bf_ips(elem, conn, coords, be, bs)
            if callable(user):
                data = np.array([[user(*xip[q, j]) for j in range(4)] for q in range(len(be))], dtype=np.float64)
            else:
                data = np.broadcast_to(np.asarray(user, dtype=np.float64).reshape(-1)[: self._dim], xip.shape).copy()
            self._be.append(np.asarray(be)); self._bs.append(np.asarray(bs)); self._bdata.append(data.reshape(len(be), 4, self._dim))

    def apply(self, disc=None):
        d = disc if disc is not None else self._master
        if self._be:
            d.set_boundary_faces(capi.BND_INFLOW, np.concatenate(self._be), np.concatenate(self._bs), np.concatenate(self._bdata, axis=0))
        return super().apply(disc)


class NavierStokesNoNormalStressOutflowFV1:
    """fv1/bnd/no_normal_stress_outflow_fv1.cpp:192-427 ("NavierStokesNoNormalStressOutflow" with the master's disc scheme,
    register_navier_stokes.cpp): zero normal stress on the outflow boundary -- tangential diffusive flux, convective flux without
    back-flow and the continuity flux over the boundary faces of the given sides, added to the Jacobian / defect of the master
    disc by disc.assemble_boundary()."""

    def __init__(self, master):
        fcts = master.symb_fcts() if hasattr(master, "symb_fcts") else master._fcts
        if len(fcts) not in (3, 4):
            raise UGError("NavierStokesNoNormalStressOutflow::set_functions: This Boundary Condition works on exactly dim+1 "
                          "(velocity+pressure) components, but %d components given." % len(fcts))
        self._master = master
        self._be, self._bs = [], []

    def add(self, elems, sides):
        """boundary sides as (element, local side) pairs (the reference names boundary subsets)"""
        self._be.append(np.asarray(elems, dtype=np.int32).reshape(-1))
        self._bs.append(np.asarray(sides, dtype=np.int32).reshape(-1))

    def apply(self, disc=None):
        d = disc if disc is not None else self._master
        if self._be:
            d.set_boundary_faces(capi.BND_OUTFLOW, np.concatenate(self._be), np.concatenate(self._bs))


NavierStokesNoNormalStressOutflow = NavierStokesNoNormalStressOutflowFV1


class FV1SmagorinskyTurbViscData:
    """fv1/turbulent_viscosity_fv1.h:200-383 ("NavierStokesFV1SmagorinskyTurbViscData", register_fv1.cpp): Smagorinsky eddy
    viscosity nu_t = c delta^2 |S| at the vertices from the deformation tensor of the current solution, interpolated to the SCVF
    ips and added to the kinematic viscosity. update(u) runs on the device and fills the per-ip viscosity import of the master
    disc (nsb_turbulent_viscosity) -- the role of passing this object to set_kinematic_viscosity() in the reference."""

    def __init__(self, master, c=0.05):
        self._master, self._c = master, float(c)
        self._be, self._bs, self._zero = [], [], []
        self.nu_t = None

    def set_model_parameter(self, c):
        self._c = float(c)

    def set_kinematic_viscosity(self, v):
        self._master.set_kinematic_viscosity(v)

    def set_turbulence_zero_bnd(self, elems, sides, nodes):
        """setTurbulenceZeroBoundaries: boundary sides (BF closure of the deformation tensor) and the vertices of those subsets
        (nu_t = 0)"""
        self._be.append(np.asarray(elems, dtype=np.int32).reshape(-1))
        self._bs.append(np.asarray(sides, dtype=np.int32).reshape(-1))
        self._zero.append(np.asarray(nodes, dtype=np.int64).reshape(-1))
        self._dirty = True

    def update(self, u, want_nodal=False):
        d = self._master
        L = capi.lib()
        p = d._params()                                          # flushes pending constant / callable imports first
        d._check(L.nsb_set_params(d._context(), C.byref(p)))
        if getattr(self, "_dirty", False):
            d.set_boundary_faces(capi.BND_TURB_ZERO, np.concatenate(self._be), np.concatenate(self._bs))
            self._dirty = False
        zero = np.ascontiguousarray(np.concatenate(self._zero), dtype=np.int64) if self._zero else np.zeros(0, dtype=np.int64)
        on_dev = _is_torch(u)
        out = None
        if on_dev:
            import torch
            d.use_stream(torch.cuda.current_stream(u.device).cuda_stream)
            if want_nodal:
                out = torch.zeros(d.num_dofs // len(d._fcts), dtype=torch.float64, device=u.device)
        else:
            u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
            if want_nodal:
                out = np.zeros(d.num_dofs // len(d._fcts))
        d._check(L.nsb_turbulent_viscosity(d._ctx, capi.TURB_SMAGORINSKY, self._c, d._ptr(u), zero.size, d._ptr(zero) if zero.size else None,
                                           d._ptr(out), capi.DEVICE if on_dev else capi.HOST))
        self.nu_t = out
        return out

    def disable(self):
        self._master._check(capi.lib().nsb_turbulent_viscosity(self._master._context(), capi.TURB_OFF, 0.0, None, 0, None, None, capi.HOST))


class DiscConstraintFVCR:
    """fvcr/disc_constraint_fvcr.h:164-1198 ("DiscConstraintFVCR", register_fvcr.cpp): post-assembly correction of the FVCR defect
    by a linear upwind reconstruction of the convected velocity (side gradients) and a linear pressure reconstruction. Same
    constructor flags as the reference; available on the device: the defect variants of the default configuration."""

    def __init__(self, disc, bLinUpConvDefect=True, bLinUpConvJacobian=False, bLinPressureDefect=True, bLinPressureJacobian=False,
                 bAdaptive=False, bLimiter=False, zero_grad_sides=None):
        if disc.disc_type() != "fvcr":
            raise UGError("DiscConstraintFVCR: works on the Crouzeix-Raviart discretisation")
        if bLinUpConvJacobian or bLinPressureJacobian or bAdaptive or bLimiter:
            raise UGError("DiscConstraintFVCR: device path provides the defect corrections only (no Jacobian variants, hanging nodes or limiter)")
        self._disc, self._up, self._pr = disc, bool(bLinUpConvDefect), bool(bLinPressureDefect)
        self._zero = np.zeros(0, dtype=np.int64)
        if zero_grad_sides is not None:
            self.set_zero_grad_bnd(zero_grad_sides)

    def set_zero_grad_bnd(self, sides):
        self._zero = np.ascontiguousarray(sides, dtype=np.int64).reshape(-1)

    def set_limiter(self, b):
        if b:
            raise UGError("DiscConstraintFVCR: the limiter is not available on the device path")

    def adjust_defect(self, d, u, scale_stiff=1.0):
        """d += correction(u) (adjust_defect :1149-1171; with a time series call once per time point with its stiffness scale)"""
        dsc = self._disc
        on_dev = _is_torch(u)
        if on_dev:
            import torch
            dsc.use_stream(torch.cuda.current_stream(u.device).cuda_stream)
        else:
            u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
            if d.dtype != np.float64 or not d.flags.c_contiguous:
                raise UGError("adjust_defect: defect must be a contiguous float64 array")
        if not self._up and not self._pr:
            return d
        z = self._zero
        dsc._check(capi.lib().nsb_fvcr_constraint_defect(dsc._context(), dsc._ptr(u), float(scale_stiff), int(self._up), int(self._pr), z.size,
                                                         dsc._ptr(z) if z.size else None, dsc._ptr(d), capi.DEVICE if on_dev else capi.HOST))
        return d


class ThetaTimeStep:
    """instationary combination (ugcore ThetaTimeStep drives add_jac_A/M, add_def_A/M with the scales
    s_m = 1, s_a = theta dt at the new time point and s_m = -1, s_a = (1 - theta) dt at the old one,
    fv1/navier_stokes_fv1.cpp:268-280,617-629): J = M + theta dt A(u_new),
    d = M u_new - M u_old + dt [theta A(u_new) + (1 - theta) A(u_old)]. Two passes over the mesh, the second accumulates
    (beta = 1); the Jacobian stays on the device (assemble_resident)."""

    def __init__(self, disc, theta=1.0):
        self.disc, self.theta = disc, float(theta)

    def assemble(self, u_new, u_old, dt, scatter_mode=None):
        d, th = self.disc, self.theta
        full = capi.JAC_A | capi.JAC_M | capi.DEF_A | capi.DEF_M | capi.RHS
        ts = (u_new, u_old, dt)
        defect = d.assemble_resident(full, u_new, time_series=ts, scale_a=th * dt, scale_m=1.0, scatter_mode=scatter_mode)
        a_old = (capi.DEF_A | capi.RHS) if th < 1.0 else 0
        defect = d.assemble_resident(capi.DEF_M | a_old, u_old, defect=defect, time_series=ts, scale_a=(1.0 - th) * dt, scale_m=-1.0,
                                     beta=1.0, scatter_mode=scatter_mode)
        return defect
