"""ctypes binding of oracle/_ref/liboracle.so (the plain-C CPU restatement).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "liboracle.so")


class OrcConfig(C.Structure):
    _fields_ = [("dim", C.c_int), ("periodic", C.c_int * 3), ("boxlo", C.c_double * 3),
                ("boxhi", C.c_double * 3), ("ntypes", C.c_int), ("nspecies", C.c_int),
                ("variant", C.c_int), ("skin", C.c_double), ("every", C.c_int), ("delay", C.c_int),
                ("check", C.c_int), ("dt", C.c_double), ("integrate_groupbit", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "sphbvf_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True,
                           stdout=subprocess.DEVNULL)
        L = C.CDLL(LIB)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcConfig)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_type.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 4
        L.orc_set_pair.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.orc_set_atoms.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 11
        L.orc_add_buoyancy.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double]
        L.orc_add_forcing.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, C.c_int, C.c_int] + [C.c_double] * 5
        L.orc_add_buffer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_long, C.c_int] + [C.c_double] * 5
        L.orc_add_setforce.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 3
        L.orc_add_chem_rxn.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        for f in ("orc_setup", "orc_build_neighbors", "orc_pair_compute", "orc_nlocal", "orc_nghost", "orc_nbuilds"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_run.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_run_length.argtypes = [C.c_void_p, C.c_long]
        L.orc_set_consistent_ghosts.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_symmetric_switch.argtypes = [C.c_void_p, C.c_int]
        L.orc_ntimestep.argtypes = [C.c_void_p]
        L.orc_ntimestep.restype = C.c_long
        L.orc_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.orc_get_int.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.orc_get_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
        L.orc_get_pairs.restype = C.c_long
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Same surface as the CUDA engine wrapper (tests/common.py drives both from one fixture)."""

    def __init__(self, meta):
        L = lib()
        cfg = OrcConfig()
        cfg.dim = meta["dim"]
        cfg.periodic[:] = meta["periodic"]
        cfg.boxlo[:] = meta["boxlo"]
        cfg.boxhi[:] = meta["boxhi"]
        cfg.ntypes = meta["ntypes"]
        cfg.nspecies = meta["S"]
        cfg.variant = meta["variant"]
        cfg.skin = meta["skin"]
        cfg.every, cfg.delay, cfg.check = meta["every"], meta["delay"], meta["check"]
        cfg.dt = meta["dt"]
        cfg.integrate_groupbit = meta.get("integrate_groupbit", 1)
        self.h = L.orc_create(C.byref(cfg))
        if not self.h:
            raise RuntimeError("orc_create failed")
        self.S = meta["S"]
        for t, tp in enumerate(meta["types"], start=1):
            self._ck(L.orc_set_type(self.h, t, tp["mass"], tp["rho0"], tp["c0"], tp["G0"]))
        for p in meta["pairs"]:
            kap = np.asarray(p["kappa"], dtype=np.float64)
            self._ck(L.orc_set_pair(self.h, p["i"], p["j"], p["eta"], p["h"], p["cutc"], _p(kap)))
        for fx in meta.get("fixes", []):
            k = fx["kind"]
            if k == "buoyancy":
                self._ck(L.orc_add_buoyancy(self.h, fx["groupbit"], fx["gravity"], fx["accel"], fx["coord"], fx["k"], fx["Cref"]))
            elif k == "forcing":
                self._ck(L.orc_add_forcing(self.h, fx["groupbit"], fx["what"], fx["step"], fx["idx"], fx["shape"],
                                           fx["cx"], fx["cy"], fx["a"], fx["b"], fx["value"]))
            elif k == "buffer":
                self._ck(L.orc_add_buffer(self.h, fx["groupbit"], fx["what"], fx["axis"], fx["step"], fx["idx"],
                                          fx["cx"], fx["cy"], fx["length"], fx["width"], fx["value"]))
            elif k == "setforce":
                self._ck(L.orc_add_setforce(self.h, fx["groupbit"], fx["fx"], fx["fy"], fx["fz"]))
            elif k == "chem_rxn":
                r = np.asarray(fx["reactants"], dtype=np.int32)
                p = np.asarray(fx["products"], dtype=np.int32)
                self._ck(L.orc_add_chem_rxn(self.h, fx["groupbit"], fx["k"], len(r), _p(r), len(p), _p(p)))

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(lib().orc_last_error(self.h).decode())

    def set_atoms(self, tag, type_, mask, solid, fixed, x, v, rho, e, Cc=None, dev=None):
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        self._keep = [i32(tag), i32(type_), i32(mask), i32(solid), i32(fixed), f64(x), f64(v), f64(rho), f64(e),
                      f64(Cc) if self.S else None, f64(dev)]
        self.n = len(self._keep[0])
        self._ck(lib().orc_set_atoms(self.h, self.n, *[_p(a) for a in self._keep]))

    def setup(self):
        self._ck(lib().orc_setup(self.h))

    def run(self, n):
        self._ck(lib().orc_run(self.h, n))

    def set_run_length(self, n):
        lib().orc_set_run_length(self.h, n)

    def set_symmetric_switch(self, on=True):
        lib().orc_set_symmetric_switch(self.h, int(on))

    def set_consistent_ghosts(self, on=True):
        lib().orc_set_consistent_ghosts(self.h, int(on))

    def build_neighbors(self):
        self._ck(lib().orc_build_neighbors(self.h))

    def pair_compute(self):
        self._ck(lib().orc_pair_compute(self.h))

    def get(self, name):
        nc = lib().orc_get(self.h, name.encode(), None)
        if nc < 0:
            raise KeyError(name)
        out = np.zeros((self.n, nc) if nc != 1 else (self.n,), dtype=np.float64)
        if nc:
            lib().orc_get(self.h, name.encode(), _p(out))
        return out

    def tags(self):
        out = np.zeros(self.n, dtype=np.int32)
        lib().orc_get_int(self.h, b"tag", _p(out))
        return out

    def pairs(self):
        n = lib().orc_get_pairs(self.h, None, 0)
        out = np.zeros((n, 2), dtype=np.int32)
        lib().orc_get_pairs(self.h, _p(out), n)
        return out

    @property
    def nbuilds(self):
        return lib().orc_nbuilds(self.h)

    @property
    def nghost(self):
        return lib().orc_nghost(self.h)

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
