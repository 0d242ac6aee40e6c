"""ctypes binding of libsphbvf.so (include/sphbvf.h).  No arithmetic here."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("SPHBVF_LIB") or os.path.join(HERE, "libsphbvf.so")   # override: A/B builds while tuning


class SphbvfError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("dim", C.c_int), ("periodic", C.c_int * 3), ("boxlo", C.c_double * 3), ("boxhi", C.c_double * 3),
                ("ntypes", C.c_int), ("nspecies", C.c_int), ("variant", C.c_int), ("skin", C.c_double),
                ("neigh_every", C.c_int), ("neigh_delay", C.c_int), ("neigh_check", C.c_int), ("dt", C.c_double),
                ("integrate_groupbit", C.c_int), ("device", C.c_int), ("procgrid", C.c_int * 3),
                ("rank", C.c_int), ("nranks", C.c_int)]


# enum sphbvf_field, in header order
_FIELD_NAMES = ["tag", "type", "mask", "solid_tag", "fixed_tag", "x", "v", "vest", "f", "rho", "rhoI", "drho", "e",
                "phi", "number_density", "nw", "ddv", "ddx", "rhoAux1", "rhoAux2", "Pnew", "dev", "ddev", "C", "Q"]
FIELDS = {n: i for i, n in enumerate(_FIELD_NAMES)}
_INT_FIELDS = {"tag", "type", "mask", "solid_tag", "fixed_tag"}
_NCOLS = {"x": 3, "v": 3, "vest": 3, "f": 3, "nw": 3, "ddv": 3, "ddx": 3, "dev": 9, "ddev": 9}

_lib = None

# every symbol include/sphbvf.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = ["sphbvf_version", "sphbvf_device_count", "sphbvf_create", "sphbvf_destroy", "sphbvf_last_error",
           "sphbvf_set_type", "sphbvf_set_pair", "sphbvf_set_dt", "sphbvf_set_random", "sphbvf_set_timestep", "sphbvf_set_run_length",
           "sphbvf_set_atoms", "sphbvf_upload", "sphbvf_download", "sphbvf_download_local", "sphbvf_upload_local", "sphbvf_add_buoyancy",
           "sphbvf_add_forcing", "sphbvf_add_buffer", "sphbvf_add_setforce", "sphbvf_add_chem_rxn", "sphbvf_max_vsq", "sphbvf_ke_tensor", "sphbvf_setup", "sphbvf_run",
           "sphbvf_setup_neighbors",
           "sphbvf_initial_integrate", "sphbvf_post_integrate", "sphbvf_neighbor", "sphbvf_pair_compute",
           "sphbvf_virial", "sphbvf_post_force", "sphbvf_setup_post_force", "sphbvf_final_integrate", "sphbvf_end_of_step", "sphbvf_build_neighbors",
           "sphbvf_nlocal", "sphbvf_nghost", "sphbvf_ntimestep", "sphbvf_nbuilds", "sphbvf_ndanger", "sphbvf_pair_mode",
           "sphbvf_get_pairs", "sphbvf_sync", "sphbvf_launch_count", "sphbvf_set_profiling", "sphbvf_kernel_ms",
           "sphbvf_stream", "sphbvf_comm_unique_id", "sphbvf_comm_init", "sphbvf_brick_bounds", "sphbvf_proc_grid",
           "sphbvf_comm_plan"]


def lib():
    """Load libsphbvf.so.  Raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise SphbvfError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C sph-bvf_b200`" % LIBPATH)
    L = C.CDLL(LIBPATH)
    vp, ci, cd, cl = C.c_void_p, C.c_int, C.c_double, C.c_long
    L.sphbvf_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.sphbvf_destroy.argtypes = [vp]
    L.sphbvf_destroy.restype = None
    L.sphbvf_last_error.argtypes = [vp]
    L.sphbvf_last_error.restype = C.c_char_p
    L.sphbvf_set_type.argtypes = [vp, ci, cd, cd, cd, cd]
    L.sphbvf_set_pair.argtypes = [vp, ci, ci, cd, cd, cd, vp]
    L.sphbvf_set_dt.argtypes = [vp, cd]
    L.sphbvf_set_random.argtypes = [vp, cd, C.c_ulonglong]
    L.sphbvf_set_timestep.argtypes = [vp, cl]
    L.sphbvf_set_run_length.argtypes = [vp, cl]
    L.sphbvf_set_atoms.argtypes = [vp, ci] + [vp] * 11
    L.sphbvf_upload.argtypes = [vp, ci, vp]
    L.sphbvf_download.argtypes = [vp, ci, vp]
    L.sphbvf_download_local.argtypes = [vp, ci, vp, ci]
    L.sphbvf_upload_local.argtypes = [vp, ci, vp, ci]
    L.sphbvf_add_buoyancy.argtypes = [vp, ci, ci, cd, ci, ci, cd]
    L.sphbvf_add_forcing.argtypes = [vp, ci, ci, cl, ci, ci, cd, cd, cd, cd, cd]
    L.sphbvf_add_buffer.argtypes = [vp, ci, ci, ci, cl, ci, cd, cd, cd, cd, cd]
    L.sphbvf_add_setforce.argtypes = [vp, ci, cd, cd, cd]
    L.sphbvf_add_chem_rxn.argtypes = [vp, ci, cd, ci, vp, ci, vp]
    L.sphbvf_max_vsq.argtypes = [vp, ci, C.POINTER(cd)]
    L.sphbvf_ke_tensor.argtypes = [vp, ci, vp]
    for f in ("setup", "setup_neighbors", "setup_post_force", "initial_integrate", "post_integrate", "pair_compute", "post_force", "final_integrate",
              "end_of_step", "build_neighbors", "nlocal", "nghost", "nbuilds", "ndanger", "sync", "pair_mode"):
        getattr(L, "sphbvf_" + f).argtypes = [vp]
    L.sphbvf_run.argtypes = [vp, ci]
    L.sphbvf_virial.argtypes = [vp, vp]
    L.sphbvf_neighbor.argtypes = [vp, C.POINTER(ci)]
    L.sphbvf_ntimestep.argtypes = [vp]
    L.sphbvf_ntimestep.restype = cl
    L.sphbvf_get_pairs.argtypes = [vp, vp, cl]
    L.sphbvf_get_pairs.restype = cl
    L.sphbvf_launch_count.argtypes = [vp]
    L.sphbvf_launch_count.restype = cl
    L.sphbvf_set_profiling.argtypes = [vp, ci]
    L.sphbvf_kernel_ms.argtypes = [vp, ci, C.POINTER(cl)]
    L.sphbvf_kernel_ms.restype = cd
    L.sphbvf_stream.argtypes = [vp]
    L.sphbvf_stream.restype = vp
    L.sphbvf_comm_unique_id.argtypes = [vp]
    L.sphbvf_comm_init.argtypes = [vp, vp]
    L.sphbvf_brick_bounds.argtypes = [C.POINTER(Config), ci, C.POINTER(cd * 3), C.POINTER(cd * 3)]
    L.sphbvf_proc_grid.argtypes = [ci, ci, C.POINTER(cd * 3), C.POINTER(ci * 3)]
    L.sphbvf_comm_plan.argtypes = [C.POINTER(Config), ci, C.POINTER(ci * 27), C.POINTER(cd * 81)]
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def config_from_meta(meta, device=0, procgrid=(1, 1, 1), rank=0, nranks=1):
    cfg = Config()
    cfg.dim = meta["dim"]
    cfg.periodic[:] = meta["periodic"]
    cfg.boxlo[:] = meta["boxlo"]
    cfg.boxhi[:] = meta["boxhi"]
    cfg.ntypes = meta["ntypes"]
    cfg.nspecies = meta["S"]
    cfg.variant = meta["variant"]
    cfg.skin = meta["skin"]
    cfg.neigh_every, cfg.neigh_delay, cfg.neigh_check = meta["every"], meta["delay"], meta["check"]
    cfg.dt = meta["dt"]
    cfg.integrate_groupbit = meta.get("integrate_groupbit", 1)
    cfg.device = device
    cfg.procgrid[:] = list(procgrid)
    cfg.rank, cfg.nranks = rank, nranks
    return cfg


class Engine:
    """One sphbvf context; the method names mirror tests/oracle_api.Oracle so parity tests drive both."""

    def __init__(self, meta, device=0, procgrid=(1, 1, 1), rank=0, nranks=1):
        L = lib()
        self.S = meta["S"]
        self.cfg = config_from_meta(meta, device, procgrid, rank, nranks)
        h = C.c_void_p()
        rc = L.sphbvf_create(C.byref(self.cfg), C.byref(h))
        if rc != 0:
            raise SphbvfError("sphbvf_create failed with %d (no CUDA device? this library has no CPU path)" % rc)
        self.h = h
        for t, tp in enumerate(meta["types"], start=1):
            self._ck(L.sphbvf_set_type(self.h, t, tp["mass"], tp["rho0"], tp["c0"], tp["G0"]))
        for p in meta["pairs"]:
            kap = np.asarray(p["kappa"] + [0.0], dtype=np.float64)
            self._ck(L.sphbvf_set_pair(self.h, p["i"], p["j"], p["eta"], p["h"], p["cutc"], _p(kap)))
        for fx in meta.get("fixes", []):
            k = fx["kind"]
            if k == "buoyancy":
                self._ck(L.sphbvf_add_buoyancy(self.h, fx["groupbit"], fx["gravity"], fx["accel"], fx["coord"], fx["k"], fx["Cref"]))
            elif k == "forcing":
                self._ck(L.sphbvf_add_forcing(self.h, fx["groupbit"], fx["what"], fx["step"], fx["idx"], fx["shape"],
                                              fx["cx"], fx["cy"], fx["a"], fx["b"], fx["value"]))
            elif k == "buffer":
                self._ck(L.sphbvf_add_buffer(self.h, fx["groupbit"], fx["what"], fx["axis"], fx["step"], fx["idx"],
                                             fx["cx"], fx["cy"], fx["length"], fx["width"], fx["value"]))
            elif k == "setforce":
                self._ck(L.sphbvf_add_setforce(self.h, fx["groupbit"], fx["fx"], fx["fy"], fx["fz"]))
            elif k == "chem_rxn":
                r = np.asarray(fx["reactants"], dtype=np.int32)
                p = np.asarray(fx["products"], dtype=np.int32)
                self._ck(L.sphbvf_add_chem_rxn(self.h, fx["groupbit"], fx["k"], len(r), _p(r), len(p), _p(p)))

    def _ck(self, rc):
        if rc != 0:
            raise SphbvfError("sphbvf error %d: %s" % (rc, lib().sphbvf_last_error(self.h).decode()))

    def set_atoms(self, tag, type_, mask, solid, fixed, x, v, rho, e, Cc=None, dev=None):
        i32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        keep = [i32(tag), i32(type_), i32(mask), i32(solid), i32(fixed), f64(x), f64(v), f64(rho), f64(e),
                f64(Cc) if self.S else None, f64(dev)]
        self.n = len(keep[0])
        self._ck(lib().sphbvf_set_atoms(self.h, self.n, *[_p(a) for a in keep]))

    def set_random(self, kboltz, seed):
        self._ck(lib().sphbvf_set_random(self.h, kboltz, seed))

    def set_run_length(self, n):
        self._ck(lib().sphbvf_set_run_length(self.h, n))

    def set_timestep(self, n):
        self._ck(lib().sphbvf_set_timestep(self.h, n))

    def setup(self):
        self._ck(lib().sphbvf_setup(self.h))

    def run(self, n):
        self._ck(lib().sphbvf_run(self.h, n))

    def build_neighbors(self):
        self._ck(lib().sphbvf_build_neighbors(self.h))

    def pair_compute(self):
        self._ck(lib().sphbvf_pair_compute(self.h))
        self.sync()

    def virial(self):
        out = np.zeros(6)
        self._ck(lib().sphbvf_virial(self.h, _p(out)))
        return out

    def ke_tensor(self, groupbit=1):
        out = np.zeros(6)
        self._ck(lib().sphbvf_ke_tensor(self.h, groupbit, _p(out)))
        return out

    def step_pieces(self):
        """One timestep through the fine-grained entry points the /cuda host classes use."""
        L = lib()
        self.set_timestep(self.ntimestep + 1)
        self._ck(L.sphbvf_initial_integrate(self.h))
        self._ck(L.sphbvf_post_integrate(self.h))
        rebuilt = C.c_int(0)
        self._ck(L.sphbvf_neighbor(self.h, C.byref(rebuilt)))
        self._ck(L.sphbvf_pair_compute(self.h))
        self._ck(L.sphbvf_post_force(self.h))
        self._ck(L.sphbvf_final_integrate(self.h))
        self._ck(L.sphbvf_end_of_step(self.h))
        return rebuilt.value

    def sync(self):
        self._ck(lib().sphbvf_sync(self.h))

    def get(self, name, local=False):
        nc = self.S if name in ("C", "Q") else _NCOLS.get(name, 1)
        dt = np.int32 if name in _INT_FIELDS else np.float64
        n = self.nlocal if local else self.n
        out = np.zeros((n, nc) if nc != 1 else (n,), dtype=dt)
        if nc:
            if local:
                self._ck(lib().sphbvf_download_local(self.h, FIELDS[name], _p(out), n))
            else:
                self._ck(lib().sphbvf_download(self.h, FIELDS[name], _p(out)))
        return out

    def put(self, name, arr):
        dt = np.int32 if name in _INT_FIELDS else np.float64
        a = np.ascontiguousarray(arr, dtype=dt)
        self._ck(lib().sphbvf_upload(self.h, FIELDS[name], _p(a)))

    def upload_ptr(self, field_id, ptr):
        self._ck(lib().sphbvf_upload(self.h, field_id, ptr))

    def download_ptr(self, field_id, ptr):
        self._ck(lib().sphbvf_download(self.h, field_id, ptr))

    def upload_local_ptr(self, field_id, ptr, nrows):
        self._ck(lib().sphbvf_upload_local(self.h, field_id, ptr, nrows))

    def download_local_ptr(self, field_id, ptr, cap_rows):
        self._ck(lib().sphbvf_download_local(self.h, field_id, ptr, cap_rows))

    def pairs(self):
        n = lib().sphbvf_get_pairs(self.h, None, 0)
        out = np.zeros((n, 2), dtype=np.int32)
        if n:
            lib().sphbvf_get_pairs(self.h, _p(out), n)
        return out

    def profiling(self, on=True):
        self._ck(lib().sphbvf_set_profiling(self.h, int(on)))

    def kernel_ms(self, which):
        n = C.c_long(0)
        ms = lib().sphbvf_kernel_ms(self.h, which, C.byref(n))
        return ms, n.value

    def comm_init(self, id_bytes):
        buf = C.create_string_buffer(bytes(id_bytes), 128)
        self._ck(lib().sphbvf_comm_init(self.h, buf))

    @property
    def launch_count(self):
        return lib().sphbvf_launch_count(self.h)

    @property
    def nlocal(self):
        return lib().sphbvf_nlocal(self.h)

    @property
    def nghost(self):
        return lib().sphbvf_nghost(self.h)

    @property
    def nbuilds(self):
        return lib().sphbvf_nbuilds(self.h)

    def pair_mode(self):
        """'tile' (16-bit slot list, candidates staged in shared memory) or 'gather' (32-bit list, records through L1)"""
        return "tile" if lib().sphbvf_pair_mode(self.h) else "gather"

    @property
    def ntimestep(self):
        return lib().sphbvf_ntimestep(self.h)

    def close(self):
        if getattr(self, "h", None):
            lib().sphbvf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id():
    buf = C.create_string_buffer(128)
    rc = lib().sphbvf_comm_unique_id(buf)
    if rc != 0:
        raise SphbvfError("sphbvf_comm_unique_id failed: %d" % rc)
    return buf.raw
