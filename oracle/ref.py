"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/_ref/libsocref_<tag>.so, i.e. the reference's
own kernels compiled in place by oracle/build_ref.py.  Same call shapes as oracle/orc.py so tests can
run either and compare.
"""
import ctypes as C

import numpy as np

from . import build_ref

c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int32)


def _fp(a):
    return a.ctypes.data_as(c_fp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def available(cfg=None):
    return build_ref.reference_available() or (cfg is not None and __import__("os").path.exists(build_ref.lib_path(cfg)))


class Reference:
    def __init__(self, cloud, gl=0.01, bins=2500, map_nside=None, **opts):
        cfg = dict(NX=cloud.NX, NY=cloud.NY, NZ=cloud.NZ, LEVELS=cloud.LEVELS, CELLS=cloud.CELLS, BINS=bins, GL=gl)
        for k, v in opts.items():
            if k in ("roi", "roi_dim"):          # run-time arrays, not macros
                continue
            cfg[k.upper()] = v
        self.roi = np.ascontiguousarray(opts.get("roi", [0] * 6), np.int32)
        self.roi_dim = np.ascontiguousarray(opts.get("roi_dim", [1, 1, 1]), np.int32)
        self.roi_save = None
        if map_nside is not None:
            cfg["MAP_NSIDE"] = map_nside
        self.cfg = cfg
        path = build_ref.build(cfg)
        if path is None:
            raise RuntimeError("reference library for %s not available" % build_ref.tag_of(cfg))
        self.L = C.CDLL(path)
        self.L.ref_atomic_count.restype = C.c_ulong
        self.cloud = cloud
        self.opts = opts
        self.lcells = np.ascontiguousarray(cloud.LCELLS, np.int32)
        self.off = np.ascontiguousarray(cloud.OFF, np.int32)
        self.dens = np.ascontiguousarray(cloud.DENS, np.float32)
        self.par = np.zeros(max(1, cloud.CELLS - cloud.NX * cloud.NY * cloud.NZ), np.int32)
        self.L.ref_parents(C.c_int(1024), _fp(self.dens), _ip(self.lcells), _ip(self.off), _ip(self.par))
        n = cloud.CELLS
        self.tabs = np.zeros(n, np.float32)
        self.xab = np.zeros(n, np.float32)
        self.int_ = np.zeros(n, np.float32)
        self.intx = np.zeros(n, np.float32)
        self.inty = np.zeros(n, np.float32)
        self.intz = np.zeros(n, np.float32)
        self._d = np.zeros(4, np.float32)
        self._di = np.zeros(4, np.int32)

    def threads(self):
        return self.L.ref_threads()

    def set_threads(self, n):
        self.L.ref_set_threads(C.c_int(n))

    def set_chunk(self, n):
        self.L.ref_set_chunk(C.c_int(n))

    def set_sampling(self, stride=1, offset=0, gsize=0):
        """Loop index i of the following launches runs work item offset + i*stride of a launch of `gsize` work items
        (bench.py's bounded samples: a stride covers the whole launch instead of its first work items)."""
        self.L.ref_set_sampling(C.c_long(stride), C.c_long(offset), C.c_long(gsize))

    def atomic_count(self, reset=True):
        return int(self.L.ref_atomic_count(C.c_int(1 if reset else 0)))

    def clear_roi_save(self):
        self._roi_save()
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

    def _f(self, a, dummy=None):
        if a is None:
            return _fp(self._d)
        a = np.ascontiguousarray(a, np.float32)
        self._keep.append(a)
        return _fp(a)

    def _i(self, a):
        if a is None:
            return _ip(self._di)
        a = np.ascontiguousarray(a, np.int32)
        self._keep.append(a)
        return _ip(a)

    def _roi_save(self):
        """ROI_SAVE accumulates over launches like a device buffer; sized from ROI, ROI_STEP, ROI_NSIDE."""
        if not self.opts.get("with_roi_save", 0):
            return _fp(self._d)
        if self.roi_save is None:
            st, r = self.opts["roi_step"], self.roi
            nx, ny, nz = [(int(r[2 * k + 1]) - int(r[2 * k]) + 1) * st for k in range(3)]
            self.roi_save = np.zeros((nx * ny + ny * nz + nz * nx) * 12 * self.opts.get("roi_nside", 16) ** 2, np.float32)
        return _fp(self.roi_save)

    def _as(self, abs_, sca, abs_v, sca_v):
        """ABS / SCA kernel arguments: one float each, or [NDUST] vectors with WITH_MSF."""
        if abs_v is not None:
            return np.ascontiguousarray(abs_v, np.float32), np.ascontiguousarray(sca_v, np.float32)
        return np.array([abs_], np.float32), np.array([sca], np.float32)

    def sim_pb(self, global_, source, packets, batch, seed, bg, tw, abs_=0.0, sca=0.0, dsc=None, csc=None, emit=None,
               emwei=None, opt=None, pspos=None, ps=None, xps_nside=None, xps_side=None, xps_area=None, abu=None,
               abs_v=None, sca_v=None, roi_load=None, **_):
        self._keep = []
        a, s = self._as(abs_, sca, abs_v, sca_v)
        self.L.ref_sim_pb(C.c_int(global_), C.c_int(source), C.c_int(packets), C.c_int(batch), C.c_float(seed),
                          _fp(a), _fp(s), C.c_float(bg), self._f(pspos), self._f(ps), C.c_float(tw),
                          _ip(self.lcells), _ip(self.off), _ip(self.par), _fp(self.dens), self._f(emit),
                          _fp(self.tabs), self._f(dsc), self._f(csc), _fp(self.xab), self._f(emwei),
                          _fp(self.int_), _fp(self.intx), _fp(self.inty), _fp(self.intz), self._f(opt),
                          self._f(abu), self._i(xps_nside), self._i(xps_side), self._f(xps_area), _ip(self.roi_dim),
                          self._f(roi_load), _ip(self.roi), self._roi_save())

    def sim_hp(self, global_, packets, batch, seed, tw, abs_=0.0, sca=0.0, dsc=None, csc=None, opt=None, hpbg=None,
               hpbgp=None, abu=None, abs_v=None, sca_v=None, **_):
        self._keep = []
        a, s = self._as(abs_, sca, abs_v, sca_v)
        self.L.ref_sim_hp(C.c_int(global_), C.c_int(packets), C.c_int(batch), C.c_float(seed), _fp(a), _fp(s),
                          C.c_float(tw), _ip(self.lcells), _ip(self.off), _ip(self.par), _fp(self.dens),
                          _fp(self._d), _fp(self.tabs), self._f(dsc), self._f(csc), _fp(self.xab), _fp(self.int_),
                          _fp(self.intx), _fp(self.inty), _fp(self.intz), self._f(opt), self._f(hpbg),
                          self._f(hpbgp), self._f(abu))

    def sim_cl(self, global_, packets, batch, seed, tw, abs_=0.0, sca=0.0, dsc=None, csc=None, emit=None, emwei=None,
               opt=None, abu=None, abs_v=None, sca_v=None, **_):
        self._keep = []
        a, s = self._as(abs_, sca, abs_v, sca_v)
        self.L.ref_sim_cl(C.c_int(global_), C.c_int(2), C.c_int(packets), C.c_int(batch), C.c_float(seed), _fp(a),
                          _fp(s), C.c_float(tw), _ip(self.lcells), _ip(self.off), _ip(self.par), _fp(self.dens),
                          self._f(emit), _fp(self.tabs), self._f(dsc), self._f(csc), _fp(self.xab), self._f(emwei),
                          _fp(self.int_), _fp(self.intx), _fp(self.inty), _fp(self.intz), _ip(self._di),
                          self._f(opt), self._f(abu), _ip(self.roi), self._roi_save())

    def eq_temperature(self, level, adhoc, kE, Emin, NE, ttt, emit, tnew):
        ttt = np.ascontiguousarray(ttt, np.float32)
        emit = np.ascontiguousarray(emit, np.float32)
        self.L.ref_eq_temperature(C.c_int(4096), C.c_int(level), C.c_float(adhoc), C.c_float(kE), C.c_float(Emin),
                                  C.c_int(NE), _ip(self.off), _ip(self.lcells), _fp(ttt), _fp(self.dens), _fp(emit),
                                  _fp(tnew))

    def emission(self, freq, fabs_, t):
        t = np.ascontiguousarray(t, np.float32)
        out = np.zeros(self.cloud.CELLS, np.float32)
        self.L.ref_emission(C.c_int(4096), C.c_float(freq), C.c_float(fabs_), _fp(self.dens), _fp(t), _fp(out))
        return out

    def emission2(self, c0, c1, freq, fabs_, t):
        t = np.ascontiguousarray(t, np.float32)
        freq = np.ascontiguousarray(freq, np.float32)
        fabs_ = np.ascontiguousarray(fabs_, np.float32)
        out = np.zeros((c1 - c0, len(freq)), np.float32)
        self.L.ref_emission2(C.c_int(4096), C.c_int(c0), C.c_int(c1), C.c_int(len(freq)), _fp(freq), _fp(fabs_), _fp(self.dens),
                             _fp(t), _fp(out))
        return out

    def mapping(self, map_dx, npx, npy, emit, dir_, ra, de, abs_, sca, centre, intobs=(-1e12, 0, 0), opt=None,
                save_colden=0):
        self._keep = []
        m = np.zeros(npx * npy, np.float32)
        t = np.zeros(npx * npy, np.float32)
        v = [np.ascontiguousarray(x, np.float32) for x in (dir_, ra, de, centre, intobs)]
        emit = np.ascontiguousarray(emit, np.float32)
        glob = (1 + (npx * npy) // 8) * 8
        self.L.ref_mapping(C.c_int(glob), C.c_float(map_dx), C.c_int(npx), C.c_int(npy), _fp(m), _fp(emit),
                           _fp(v[0]), _fp(v[1]), _fp(v[2]), _ip(self.lcells), _ip(self.off), _ip(self.par),
                           _fp(self.dens), C.c_float(abs_), C.c_float(sca), _fp(v[3]), _fp(v[4]), self._f(opt),
                           _fp(t), C.c_int(save_colden), _ip(self.roi))
        return m.reshape(npy, npx), t.reshape(npy, npx)

    def mapping_levels(self, map_dx, npx, npy, emit, dir_, ra, de, abs_, sca, centre, intobs=(-1e12, 0, 0), opt=None,
                       colden=False):
        """kernel_ASOC_map_H.c Mapping through its own library (build_ref.build_levels)."""
        cfg = dict(self.cfg)
        cfg["USE_ABU"] = 1 if opt is not None else 0
        cfg["WITH_COLDEN"] = 1 if colden else 0
        path = build_ref.build_levels(cfg)
        if path is None:
            raise RuntimeError("reference per-level map library not available")
        L = C.CDLL(path)
        self._keep = []
        levels = int(self.cloud.LEVELS)
        m = np.zeros(levels * npx * npy, np.float32)
        cd = np.zeros(npx * npy, np.float32)
        v = [np.ascontiguousarray(x, np.float32) for x in (dir_, ra, de, centre, intobs)]
        emit = np.ascontiguousarray(emit, np.float32)
        glob = (1 + (npx * npy) // 8) * 8
        L.ref_mapping_levels(C.c_int(glob), C.c_float(map_dx), C.c_int(npx), C.c_int(npy), _fp(m), _fp(emit),
                             _fp(v[0]), _fp(v[1]), _fp(v[2]), _ip(self.lcells), _ip(self.off), _ip(self.par),
                             _fp(self.dens), C.c_float(abs_), C.c_float(sca), _fp(v[3]), _fp(v[4]), self._f(opt), _fp(cd))
        m = m.reshape(levels, npy, npx)
        return (m, cd.reshape(npy, npx)) if colden else m

    def healpix_mapping(self, nside, emit, abs_, sca, intobs, opt=None, save_colden=0):
        self._keep = []
        n = 12 * nside * nside
        m, t = np.zeros(n, np.float32), np.zeros(n, np.float32)
        z = np.zeros(3, np.float32)
        io = np.ascontiguousarray(intobs, np.float32)
        emit = np.ascontiguousarray(emit, np.float32)
        self.L.ref_healpix_mapping(C.c_int(n), C.c_float(1.0), C.c_int(nside), C.c_int(0), _fp(m), _fp(emit),
                                   _fp(z), _fp(z), _fp(z), _ip(self.lcells), _ip(self.off), _ip(self.par),
                                   _fp(self.dens), C.c_float(abs_), C.c_float(sca), _fp(z), _fp(io), self._f(opt),
                                   _fp(t), C.c_int(save_colden), _ip(self.roi))
        return m, t

    def ps_tau(self, pspos, dir_, abs_, sca, opt=None):
        self._keep = []
        pp = np.ascontiguousarray(np.asarray(pspos, np.float32).reshape(-1, 3))
        no = len(pp)
        col, tau = np.zeros(no, np.float32), np.zeros(no, np.float32)
        d = np.ascontiguousarray(dir_, np.float32)
        z = np.zeros(3, np.float32)
        self.L.ref_pstau(C.c_int(((no + 7) // 8) * 8), C.c_int(no), _fp(pp.reshape(-1)), _fp(d), _fp(z), _fp(z), _ip(self.lcells),
                         _ip(self.off), _ip(self.par), _fp(self.dens), C.c_float(abs_), C.c_float(sca), self._f(opt),
                         _fp(col), _fp(tau))
        return col, tau

    def _v3(self, a):
        a = np.ascontiguousarray(np.asarray(a, np.float32)[:, :3].reshape(-1))
        self._keep.append(a)
        return _fp(a)

    def _out(self, ndir, npx, npy):
        return np.zeros(ndir * npx * npy if ndir > 0 else 12 * ndir * ndir, np.float32)

    @staticmethod
    def _shape(out, ndir, npx, npy):
        return out.reshape(ndir, npy, npx) if ndir > 0 else out

    def sca_ps(self, global_, packets, batch, seed, ndir, npx, npy, map_dx, centre, odirs, ora, ode, abs_=0.0,
               sca=0.0, dsc=None, csc=None, opt=None, pspos=None, ps=None, abu=None, abs_v=None, sca_v=None, **_):
        self._keep = []
        a, s = self._as(abs_, sca, abs_v, sca_v)
        out = self._out(ndir, npx, npy)
        ce = np.ascontiguousarray(centre, np.float32)
        self.L.ref_sca_ps(C.c_int(global_), C.c_int(packets), C.c_int(batch), C.c_float(seed), _fp(a), _fp(s),
                          C.c_float(0.0), self._f(pspos), self._f(ps), _ip(self.lcells), _ip(self.off),
                          _ip(self.par), _fp(self.dens), self._f(dsc), self._f(csc), C.c_int(ndir), self._v3(odirs),
                          C.c_int(npx), C.c_int(npy), C.c_float(map_dx), _fp(ce), self._v3(ora), self._v3(ode),
                          _fp(out), self._f(abu), self._f(opt), _fp(self._d), _fp(self._d), _fp(self._d))
        return self._shape(out, ndir, npx, npy)

    def sca_pb(self, global_, source, packets, batch, seed, bg, ndir, npx, npy, map_dx, centre, odirs, ora, ode,
               abs_=0.0, sca=0.0, dsc=None, csc=None, opt=None, pspos=None, ps=None, abu=None, abs_v=None, sca_v=None,
               roi_load=None, **_):
        self._keep = []
        a, s = self._as(abs_, sca, abs_v, sca_v)
        out = self._out(ndir, npx, npy)
        ce = np.ascontiguousarray(centre, np.float32)
        self.L.ref_sca_pb(C.c_int(global_), C.c_int(source), C.c_int(packets), C.c_int(batch), C.c_float(seed),
                          _fp(a), _fp(s), C.c_float(bg), self._f(pspos), self._f(ps), _ip(self.lcells),
                          _ip(self.off), _ip(self.par), _fp(self.dens), self._f(dsc), self._f(csc), C.c_int(ndir),
                          self._v3(odirs), C.c_int(npx), C.c_int(npy), C.c_float(map_dx), _fp(ce), self._v3(ora),
                          self._v3(ode), _fp(out), self._f(abu), self._f(opt), _fp(self._d), _fp(self._d),
                          _fp(self._d), _ip(self.roi_dim), self._f(roi_load))
        return self._shape(out, ndir, npx, npy)

    def sca_hp(self, global_, packets, batch, seed, ndir, npx, npy, map_dx, centre, odirs, ora, ode, abs_=0.0,
               sca=0.0, dsc=None, csc=None, opt=None, hpbg=None, hpbgp=None, abu=None, abs_v=None, sca_v=None, **_):
        self._keep = []
        a, s = self._as(abs_, sca, abs_v, sca_v)
        out = self._out(ndir, npx, npy)
        ce = np.ascontiguousarray(centre, np.float32)
        self.L.ref_sca_hp(C.c_int(global_), C.c_int(packets), C.c_int(batch), C.c_float(seed), _fp(a), _fp(s),
                          _ip(self.lcells), _ip(self.off), _ip(self.par), _fp(self.dens), self._f(dsc), self._f(csc),
                          C.c_int(ndir), self._v3(odirs), C.c_int(npx), C.c_int(npy), C.c_float(map_dx), _fp(ce),
                          self._v3(ora), self._v3(ode), _fp(out), self._f(abu), self._f(opt), self._f(hpbg),
                          self._f(hpbgp))
        return self._shape(out, ndir, npx, npy)

    def sca_cl(self, global_, packets, batch, seed, ndir, npx, npy, map_dx, centre, odirs, ora, ode, abs_=0.0,
               sca=0.0, dsc=None, csc=None, opt=None, emit=None, emwei=None, abu=None, abs_v=None, sca_v=None, **_):
        self._keep = []
        a, s = self._as(abs_, sca, abs_v, sca_v)
        out = self._out(ndir, npx, npy)
        ce = np.ascontiguousarray(centre, np.float32)
        self.L.ref_sca_cl(C.c_int(global_), C.c_int(2), C.c_int(packets), C.c_int(batch), C.c_float(seed), _fp(a),
                          _fp(s), _ip(self.lcells), _ip(self.off), _ip(self.par), _fp(self.dens), self._f(emit),
                          self._f(dsc), self._f(csc), C.c_int(ndir), self._v3(odirs), C.c_int(npx), C.c_int(npy),
                          C.c_float(map_dx), _fp(ce), self._v3(ora), self._v3(ode), _fp(out), self._f(opt),
                          self._f(abu), self._f(emwei))
        return self._shape(out, ndir, npx, npy)


def rng_stream(lib, seed, id_, gsize, n):
    out = np.zeros(n, np.uint32)
    st = np.zeros(2, np.uint32)
    lib.ref_rng_stream(C.c_float(seed), C.c_long(id_), C.c_long(gsize), C.c_int(n),
                       out.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p))
    return st, out


def split_absorbed(idust, rabs, abu, absorbed):
    """The reference's split_absorbed kernel (kernel_A2E_MABU_aux.c) on host arrays; None where the library cannot be built."""
    rabs = np.ascontiguousarray(rabs, np.float64)
    abu = np.ascontiguousarray(abu, np.float32)
    a = np.ascontiguousarray(absorbed, np.float32)
    path = build_ref.build_a2e(a.shape[1], abu.shape[1])
    if path is None:
        return None
    out = np.zeros_like(a)
    n = a.shape[0]
    C.CDLL(path).ref_split_absorbed(C.c_int(((n + 15) // 16) * 16), C.c_int(idust), C.c_int(n), rabs.ctypes.data_as(C.POINTER(C.c_double)),
                                    _fp(abu), _fp(a), _fp(out))
    return out
