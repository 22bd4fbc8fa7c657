"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/libsoc_oracle.so (the plain-C restatement).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsoc_oracle.so")

c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int32)


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("soc_oracle.c", "soc_oracle.h", "soc_oracle_index.inc")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "libsoc_oracle.so"])
    return LIB


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "nx", "ny", "nz", "levels", "cells", "bins", "no_ps", "ps_method", "with_abu", "with_ali", "noabsorbed",
        "save_intensity", "use_emweight", "hpbg_weighted", "ffs", "step_weight", "level_threshold", "sca_exact_level")] + \
        [(n, C.c_float) for n in ("sw_a", "sw_b", "length", "factor", "adhoc", "reserved1")] + \
        [(n, C.c_int32) for n in ("with_msf", "ndust", "mirror", "map_interpolation", "hg_test", "mirror_exact", "maph_literal", "r2c",
                                  "with_roi_load", "with_roi_save", "roi_map", "roi_step", "roi_nside")] + \
        [("roi", C.c_int32 * 6), ("roi_dim", C.c_int32 * 3)]


class OrcGrid(C.Structure):
    _fields_ = [("lcells", c_ip), ("off", c_ip), ("dens", c_fp), ("par", c_ip)]


class OrcSimBufs(C.Structure):
    _fields_ = [(n, c_fp) for n in ("tabs", "xab", "intens", "intx", "inty", "intz", "emit", "emwei", "opt",
                                    "dsc", "csc")] + \
        [("abs", C.c_float), ("sca", C.c_float), ("pspos", c_fp), ("ps", c_fp), ("xps_nside", c_ip),
         ("xps_side", c_ip), ("xps_area", c_fp), ("hpbg", c_fp), ("hpbgp", c_fp), ("roi_load", c_fp), ("roi_save", c_fp),
         ("abu", c_fp), ("abs_v", c_fp),
         ("sca_v", c_fp)]


class OrcScaBufs(C.Structure):
    _fields_ = [("ndir", C.c_int32), ("npix_x", C.c_int32), ("npix_y", C.c_int32), ("reserved", C.c_int32),
                ("map_dx", C.c_float), ("centre", C.c_float * 3),
                ("odirs", c_fp), ("ora", c_fp), ("ode", c_fp), ("out", c_fp)]


class OrcCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("packets", "steps", "scatterings", "peels")]


def _fp(a):
    return None if a is None else a.ctypes.data_as(c_fp)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_ang2pix_ring.restype = C.c_int
        _lib.orc_ang2pix_ring.argtypes = [C.c_int, C.c_float, C.c_float]
    return _lib


class Oracle:
    """Holds a grid + parameter block and exposes the oracle entry points with numpy arguments.
    Keyword `opts` are the former -D macros (with_abu, noabsorbed, save_intensity, ...)."""

    def __init__(self, cloud, gl=0.01, bins=2500, **opts):
        self.L = lib()
        self.cloud = cloud
        self.opts = dict(opts)
        P = OrcParams()
        P.nx, P.ny, P.nz, P.levels, P.cells = cloud.NX, cloud.NY, cloud.NZ, cloud.LEVELS, cloud.CELLS
        P.bins = bins
        P.no_ps = max(1, opts.get("no_ps", 1))
        P.ps_method = opts.get("ps_method", 0)
        P.with_abu = opts.get("with_abu", 0)
        P.with_ali = opts.get("with_ali", 0)
        P.noabsorbed = opts.get("noabsorbed", 1)
        P.save_intensity = opts.get("save_intensity", 0)
        P.use_emweight = opts.get("use_emweight", 0)
        P.hpbg_weighted = opts.get("hpbg_weighted", 0)
        P.ffs = opts.get("ffs", 1)
        P.step_weight = opts.get("step_weight", -1)
        P.level_threshold = opts.get("level_threshold", 0)
        P.sca_exact_level = opts.get("sca_exact_level", 0)
        P.sw_a, P.sw_b = opts.get("sw_a", 0.0), opts.get("sw_b", 0.0)
        P.with_msf, P.ndust, P.mirror = opts.get("with_msf", 0), opts.get("ndust", 1), opts.get("mirror", 0)
        P.map_interpolation = opts.get("map_interpolation", 0)
        P.hg_test = opts.get("hg_test", 0)
        P.maph_literal = opts.get("maph_literal", 0)
        P.with_roi_load, P.with_roi_save = opts.get("with_roi_load", 0), opts.get("with_roi_save", 0)
        P.roi_map, P.roi_step, P.roi_nside = opts.get("roi_map", 0), opts.get("roi_step", 0), opts.get("roi_nside", 16)
        for k, v in enumerate(opts.get("roi", [0] * 6)):
            P.roi[k] = int(v)
        for k, v in enumerate(opts.get("roi_dim", [1, 1, 1])):
            P.roi_dim[k] = int(v)
        self.roi_save = None
        if P.with_roi_save:
            nx, ny, nz = [(P.roi[2 * k + 1] - P.roi[2 * k] + 1) * P.roi_step for k in range(3)]
            self.roi_save = np.zeros((nx * ny + ny * nz + nz * nx) * 12 * P.roi_nside ** 2, np.float32)
        P.mirror_exact = opts.get("mirror_exact", 0)
        P.length = float("%.5e" % (gl * 3.08567758e+18))     # -D LENGTH=%.5ef (ASOC.py:347,356)
        P.factor = 1.0e20
        P.adhoc = 1.0
        self.P = P
        self.lcells = np.ascontiguousarray(cloud.LCELLS, np.int32)
        self.off = np.ascontiguousarray(cloud.OFF, np.int32)
        self.dens = np.ascontiguousarray(cloud.DENS, np.float32)
        self.par = np.zeros(max(1, cloud.CELLS - cloud.NX * cloud.NY * cloud.NZ), np.int32)
        self.G = OrcGrid(_ip(self.lcells), _ip(self.off), _fp(self.dens), _ip(self.par))
        self.L.orc_parents(C.byref(self.P), C.byref(self.G))
        n = cloud.CELLS
        self.tabs = np.zeros(n, np.float32)
        self.xab = np.zeros(n, np.float32)
        self.int_ = np.zeros(n, np.float32)
        self.intx = np.zeros(n, np.float32)
        self.inty = np.zeros(n, np.float32)
        self.intz = np.zeros(n, np.float32)
        self.counters = OrcCounters()
        self._keep = []

    def _bufs(self, abs_=0.0, sca=0.0, dsc=None, csc=None, emit=None, emwei=None, opt=None, pspos=None, ps=None,
              xps_nside=None, xps_side=None, xps_area=None, hpbg=None, hpbgp=None, abu=None, abs_v=None, sca_v=None,
              roi_load=None):
        def f(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, np.float32)
            self._keep.append(a)
            return a

        def i(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, np.int32)
            self._keep.append(a)
            return a
        self._keep = []
        B = OrcSimBufs()
        B.tabs, B.xab, B.intens = _fp(self.tabs), _fp(self.xab), _fp(self.int_)
        B.intx, B.inty, B.intz = _fp(self.intx), _fp(self.inty), _fp(self.intz)
        B.emit, B.emwei, B.opt = _fp(f(emit)), _fp(f(emwei)), _fp(f(opt))
        B.dsc, B.csc = _fp(f(dsc)), _fp(f(csc))
        B.abs, B.sca = float(abs_), float(sca)
        B.pspos, B.ps = _fp(f(pspos)), _fp(f(ps))
        B.xps_nside, B.xps_side, B.xps_area = _ip(i(xps_nside)), _ip(i(xps_side)), _fp(f(xps_area))
        B.hpbg, B.hpbgp = _fp(f(hpbg)), _fp(f(hpbgp))
        B.abu, B.abs_v, B.sca_v = _fp(f(abu)), _fp(f(abs_v)), _fp(f(sca_v))
        B.roi_load, B.roi_save = _fp(f(roi_load)), _fp(self.roi_save)
        return B

    def clear_roi_save(self):
        if self.roi_save is not None:
            self.roi_save[:] = 0

    def zero(self, tag):
        if tag == 0:
            self.tabs[:] = 0
            self.xab[:] = 0
        else:
            self.int_[:] = 0
            self.intx[:] = 0
            self.inty[:] = 0
            self.intz[:] = 0

    def sim_pb(self, global_, source, packets, batch, seed, bg, tw, **bufs):
        B = self._bufs(**bufs)
        self.L.orc_sim_pb(C.byref(self.P), C.byref(self.G), C.byref(B), C.c_int(global_), C.c_int(source),
                          C.c_int(packets), C.c_int(batch), C.c_float(seed), C.c_float(bg), C.c_float(tw),
                          C.byref(self.counters))

    def sim_hp(self, global_, packets, batch, seed, tw, **bufs):
        B = self._bufs(**bufs)
        self.L.orc_sim_hp(C.byref(self.P), C.byref(self.G), C.byref(B), C.c_int(global_), C.c_int(packets),
                          C.c_int(batch), C.c_float(seed), C.c_float(tw), C.byref(self.counters))

    def sim_cl(self, global_, packets, batch, seed, tw, **bufs):
        B = self._bufs(**bufs)
        self.L.orc_sim_cl(C.byref(self.P), C.byref(self.G), C.byref(B), C.c_int(global_), C.c_int(packets),
                          C.c_int(batch), C.c_float(seed), C.c_float(tw), C.byref(self.counters))

    def eq_temperature(self, level, adhoc, kE, Emin, NE, ttt, emit, tnew):
        ttt = np.ascontiguousarray(ttt, np.float32)
        emit = np.ascontiguousarray(emit, np.float32)
        self.L.orc_eq_temperature(C.byref(self.P), C.byref(self.G), C.c_int(level), C.c_float(adhoc), C.c_float(kE),
                                  C.c_float(Emin), C.c_int(NE), _fp(ttt), _fp(emit), _fp(tnew))

    def emission(self, freq, fabs_, t):
        t = np.ascontiguousarray(t, np.float32)
        out = np.zeros(self.cloud.CELLS, np.float32)
        self.L.orc_emission(C.byref(self.P), C.byref(self.G), C.c_float(freq), C.c_float(fabs_), _fp(t), _fp(out))
        return out

    def emission2(self, c0, c1, freq, fabs_, t):
        t = np.ascontiguousarray(t, np.float32)
        freq = np.ascontiguousarray(freq, np.float32)
        fabs_ = np.ascontiguousarray(fabs_, np.float32)
        out = np.zeros((c1 - c0, len(freq)), np.float32)
        self.L.orc_emission2(C.byref(self.P), C.c_int(c0), C.c_int(c1), C.c_int(len(freq)), _fp(freq), _fp(fabs_), _fp(t), _fp(out))
        return out

    def mapping(self, map_dx, npx, npy, emit, dir_, ra, de, abs_, sca, centre, intobs=(-1e12, 0, 0), opt=None,
                save_colden=0):
        m = np.zeros(npx * npy, np.float32)
        t = np.zeros(npx * npy, np.float32)
        v = [np.ascontiguousarray(x, np.float32) for x in (dir_, ra, de, centre, intobs)]
        emit = np.ascontiguousarray(emit, np.float32)
        opt = None if opt is None else np.ascontiguousarray(opt, np.float32)
        self.L.orc_mapping(C.byref(self.P), C.byref(self.G), C.c_float(map_dx), C.c_int(npx), C.c_int(npy), _fp(m),
                           _fp(emit), _fp(v[0]), _fp(v[1]), _fp(v[2]), C.c_float(abs_), C.c_float(sca), _fp(v[3]),
                           _fp(v[4]), _fp(opt), _fp(t), C.c_int(save_colden), C.byref(self.counters))
        return m.reshape(npy, npx), t.reshape(npy, npx)

    def mapping_levels(self, map_dx, npx, npy, emit, dir_, ra, de, abs_, sca, centre, intobs=(-1e12, 0, 0), opt=None,
                       colden=False):
        """kernel_ASOC_map_H.c Mapping: images [LEVELS, npy, npx] (+ column density image when colden)."""
        levels = int(self.P.levels)
        m = np.zeros(levels * npx * npy, np.float32)
        cd = np.zeros(npx * npy, np.float32) if colden else None
        v = [np.ascontiguousarray(x, np.float32) for x in (dir_, ra, de, centre, intobs)]
        emit = np.ascontiguousarray(emit, np.float32)
        opt = None if opt is None else np.ascontiguousarray(opt, np.float32)
        self.L.orc_mapping_levels(C.byref(self.P), C.byref(self.G), C.c_float(map_dx), C.c_int(npx), C.c_int(npy), _fp(m),
                                  _fp(emit), _fp(v[0]), _fp(v[1]), _fp(v[2]), C.c_float(abs_), C.c_float(sca), _fp(v[3]),
                                  _fp(v[4]), _fp(opt), _fp(cd))
        m = m.reshape(levels, npy, npx)
        return (m, cd.reshape(npy, npx)) if colden else m

    def ps_tau(self, pspos, dir_, abs_, sca, opt=None):
        pp = np.ascontiguousarray(np.asarray(pspos, np.float32).reshape(-1))
        no = len(pp) // 3
        col, tau = np.zeros(no, np.float32), np.zeros(no, np.float32)
        d = np.ascontiguousarray(dir_, np.float32)
        opt = None if opt is None else np.ascontiguousarray(opt, np.float32)
        self.L.orc_ps_tau(C.byref(self.P), C.byref(self.G), C.c_int(no), _fp(pp), _fp(d), C.c_float(abs_), C.c_float(sca),
                          _fp(opt), _fp(col), _fp(tau))
        return col, tau

    def healpix_mapping(self, nside, emit, abs_, sca, intobs, opt=None, save_colden=0):
        n = 12 * nside * nside
        m, t = np.zeros(n, np.float32), np.zeros(n, np.float32)
        emit = np.ascontiguousarray(emit, np.float32)
        io = np.ascontiguousarray(intobs, np.float32)
        opt = None if opt is None else np.ascontiguousarray(opt, np.float32)
        self.L.orc_healpix_mapping(C.byref(self.P), C.byref(self.G), C.c_int(nside), _fp(m), _fp(emit),
                                   C.c_float(abs_), C.c_float(sca), _fp(io), _fp(opt), _fp(t), C.c_int(save_colden))
        return m, t

    def _sca(self, ndir, npx, npy, map_dx, centre, odirs, ora, ode):
        S = OrcScaBufs()
        S.ndir, S.npix_x, S.npix_y, S.map_dx = ndir, npx, npy, map_dx
        S.centre[0], S.centre[1], S.centre[2] = [float(x) for x in centre]
        self._sk = [np.ascontiguousarray(np.asarray(x, np.float32)[:, :3].reshape(-1)) for x in (odirs, ora, ode)]
        self.out = np.zeros(ndir * npx * npy if ndir > 0 else 12 * ndir * ndir, np.float32)
        S.odirs, S.ora, S.ode, S.out = _fp(self._sk[0]), _fp(self._sk[1]), _fp(self._sk[2]), _fp(self.out)
        return S

    def sca_ps(self, global_, packets, batch, seed, ndir, npx, npy, map_dx, centre, odirs, ora, ode, **bufs):
        B = self._bufs(**bufs)
        S = self._sca(ndir, npx, npy, map_dx, centre, odirs, ora, ode)
        self.L.orc_sca_ps(C.byref(self.P), C.byref(self.G), C.byref(B), C.byref(S), C.c_int(global_),
                          C.c_int(packets), C.c_int(batch), C.c_float(seed), C.byref(self.counters))
        return self.out.reshape(ndir, npy, npx) if ndir > 0 else self.out

    def sca_pb(self, global_, source, packets, batch, seed, bg, ndir, npx, npy, map_dx, centre, odirs, ora, ode,
               **bufs):
        B = self._bufs(**bufs)
        S = self._sca(ndir, npx, npy, map_dx, centre, odirs, ora, ode)
        self.L.orc_sca_pb(C.byref(self.P), C.byref(self.G), C.byref(B), C.byref(S), C.c_int(global_),
                          C.c_int(source), C.c_int(packets), C.c_int(batch), C.c_float(seed), C.c_float(bg),
                          C.byref(self.counters))
        return self.out.reshape(ndir, npy, npx) if ndir > 0 else self.out

    def sca_hp(self, global_, packets, batch, seed, ndir, npx, npy, map_dx, centre, odirs, ora, ode, **bufs):
        B = self._bufs(**bufs)
        S = self._sca(ndir, npx, npy, map_dx, centre, odirs, ora, ode)
        self.L.orc_sca_hp(C.byref(self.P), C.byref(self.G), C.byref(B), C.byref(S), C.c_int(global_),
                          C.c_int(packets), C.c_int(batch), C.c_float(seed), C.byref(self.counters))
        return self.out.reshape(ndir, npy, npx) if ndir > 0 else self.out

    def sca_cl(self, global_, packets, batch, seed, ndir, npx, npy, map_dx, centre, odirs, ora, ode, **bufs):
        B = self._bufs(**bufs)
        S = self._sca(ndir, npx, npy, map_dx, centre, odirs, ora, ode)
        self.L.orc_sca_cl(C.byref(self.P), C.byref(self.G), C.byref(B), C.byref(S), C.c_int(global_),
                          C.c_int(packets), C.c_int(batch), C.c_float(seed), C.byref(self.counters))
        return self.out.reshape(ndir, npy, npx) if ndir > 0 else self.out


def set_threads(n):
    lib().orc_set_threads(C.c_int(n))


def set_chunk(n):
    lib().orc_set_chunk(C.c_int(n))


def threads():
    return lib().orc_threads()


def rng_stream(seed, id_, n):
    out = np.zeros(n, np.uint32)
    st = np.zeros(2, np.uint32)
    lib().orc_rng_stream(C.c_float(seed), C.c_int64(id_), C.c_int(n), out.ctypes.data_as(C.c_void_p),
                         st.ctypes.data_as(C.c_void_p))
    return st, out


def rng_stream_base(base, id_, n):
    out = np.zeros(n, np.uint32)
    st = np.zeros(2, np.uint32)
    lib().orc_rng_stream_base(C.c_uint64(base), C.c_int64(id_), C.c_int(n), out.ctypes.data_as(C.c_void_p),
                              st.ctypes.data_as(C.c_void_p))
    return st, out


def split_absorbed(idust, rabs, abu, absorbed):
    """kernel_A2E_MABU_aux.c split_absorbed on host arrays: rabs [nfreq, ndust] float64, abu [cells, ndust], absorbed [cells, nfreq]."""
    rabs = np.ascontiguousarray(rabs, np.float64)
    abu = np.ascontiguousarray(abu, np.float32)
    a = np.ascontiguousarray(absorbed, np.float32)
    out = np.zeros_like(a)
    lib().orc_split_absorbed(C.c_int(idust), C.c_int(a.shape[0]), C.c_int(a.shape[1]), C.c_int(abu.shape[1]),
                           rabs.ctypes.data_as(C.POINTER(C.c_double)), _fp(abu), _fp(a), _fp(out))
    return out
