"""ctypes binding of libsoc_b200.so (include/soc_b200.h) -- the layer that replaces pyopencl in the SOC
drivers.  `Device` is a thin 1:1 wrapper of the C ABI; `Backend` adds the buffer bookkeeping the reference
scripts do by hand (ASOC.py:426-539) and has the call shape of the reference kernels
(SimRAM_PB / SimRAM_HP / SimRAM_CL / Mapping / HealpixMapping / EqTemperature / Emission, and the
scattered-light SimRAM_PS / SimRAM_PB of ASOCS.py).

There is no CPU fallback: if the library is missing or no sm_100 device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_lib", "libsoc_b200.so")

(BUF_DENS, BUF_PAR, BUF_TABS, BUF_XAB, BUF_INT, BUF_INTX, BUF_INTY, BUF_INTZ, BUF_EMIT, BUF_EMWEI, BUF_OPT, BUF_DSC,
 BUF_CSC, BUF_PSPOS, BUF_PS, BUF_XPS_NSIDE, BUF_XPS_SIDE, BUF_XPS_AREA, BUF_HPBG, BUF_HPBGP, BUF_MAP, BUF_SAVETAU,
 BUF_OUT, BUF_ODIR, BUF_ORA, BUF_ODE, BUF_TTT, BUF_TNEW, BUF_FABS, BUF_ABU, BUF_ABSV, BUF_SCAV, BUF_ROI_LOAD,
 BUF_ROI_SAVE, BUF_COUNT) = range(35)

RNG_REFERENCE, RNG_PACKET = 0, 1
DEP_RED, DEP_WARP, DEP_TILE = 0, 1, 2


class SocError(RuntimeError):
    pass


class SocParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "bins", "no_ps", "ps_method", "with_abu", "with_ali", "noabsorbed", "save_intensity", "use_emweight",
        "hpbg_weighted", "ffs", "step_weight", "level_threshold", "with_msf", "mirror", "dir_weight", "do_split",
        "roi_flags", "map_interpolation")] + \
        [(n, C.c_float) for n in ("sw_a", "sw_b", "length", "factor", "adhoc", "reserved")] + \
        [("ndust", C.c_int32), ("opt_is_half", C.c_int32), ("ref_quirks", C.c_int32), ("reserved2", C.c_int32 * 1)]


class SocCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("packets", "steps", "scatterings", "peels", "launches")] + \
        [("reserved", C.c_uint64 * 3)]


_EXPORTS = """soc_last_error soc_version soc_create soc_destroy soc_sync soc_set_params soc_set_grid soc_set_rng_mode
soc_build_opt soc_split_absorbed soc_set_shard soc_set_tuning soc_set_geometry soc_set_layout soc_set_domains soc_set_roi soc_upload soc_download soc_clear soc_device_ptr soc_zero_amc soc_sim_pb soc_sim_hp
soc_sim_cl soc_absorbed_begin soc_absorbed_add soc_absorbed_finish soc_eq_temperature soc_emission soc_emission2 soc_mapping soc_mapping_levels soc_healpix_mapping soc_ps_tau soc_sca_zero_out soc_sca_ps soc_sca_pb soc_sca_hp soc_sca_cl
soc_get_counters soc_reset_counters soc_last_launch_ms soc_last_kernel soc_stream""".split()

_lib = None


def load_library(path=None):
    """dlopen the CUDA library; raises SocError when it has not been built (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("SOC_B200_LIB") or LIB_PATH        # SOC_B200_LIB: development builds (tools/)
    if not os.path.exists(p):
        raise SocError("%s not found -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(soc_b200 has no CPU fallback)" % p)
    L = C.CDLL(p)
    fp, vp, i, f = C.POINTER(C.c_float), C.c_void_p, C.c_int, C.c_float
    L.soc_last_error.restype = C.c_char_p
    L.soc_create.argtypes = [i, C.POINTER(vp)]
    L.soc_destroy.argtypes = [vp]
    L.soc_sync.argtypes = [vp]
    L.soc_set_params.argtypes = [vp, C.POINTER(SocParams)]
    L.soc_set_grid.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, vp, vp, vp]
    L.soc_set_rng_mode.argtypes = [vp, i]
    L.soc_set_shard.argtypes = [vp, i, i]
    L.soc_set_tuning.argtypes = [vp, i, i, i]
    L.soc_set_geometry.argtypes = [vp, i]
    L.soc_set_layout.argtypes = [vp, i]
    L.soc_set_domains.argtypes = [vp, i]
    L.soc_set_roi.argtypes = [vp, vp, i, i, vp]
    L.soc_upload.argtypes = [vp, i, vp, C.c_size_t]
    L.soc_build_opt.argtypes = [vp, i, fp, fp, i, i]
    L.soc_split_absorbed.argtypes = [vp, i, i, i, C.POINTER(C.c_double), vp]
    L.soc_download.argtypes = [vp, i, vp, C.c_size_t]
    L.soc_clear.argtypes = [vp, i, C.c_size_t]
    L.soc_device_ptr.argtypes = [vp, i, C.POINTER(C.c_size_t)]
    L.soc_device_ptr.restype = vp
    L.soc_stream.argtypes = [vp]
    L.soc_stream.restype = vp
    L.soc_zero_amc.argtypes = [vp, i]
    L.soc_sim_pb.argtypes = [vp, i, i, i, f, f, f, f, f, i]
    L.soc_sim_hp.argtypes = [vp, i, i, f, f, f, f, i]
    L.soc_sim_cl.argtypes = [vp, i, i, i, f, f, f, f, i]
    L.soc_eq_temperature.argtypes = [vp, i, f, f, f, i]
    L.soc_absorbed_begin.argtypes = [vp, i]
    L.soc_absorbed_add.argtypes = [vp, i]
    L.soc_absorbed_finish.argtypes = [vp, f, f, i, vp]
    L.soc_emission.argtypes = [vp, f, f]
    L.soc_emission2.argtypes = [vp, i, i, i, vp, vp, vp]
    L.soc_mapping.argtypes = [vp, f, i, i, fp, fp, fp, f, f, fp, fp, i]
    L.soc_mapping_levels.argtypes = [vp, f, i, i, fp, fp, fp, f, f, fp, fp, i]
    L.soc_healpix_mapping.argtypes = [vp, i, f, f, fp, i]
    L.soc_ps_tau.argtypes = [vp, i, fp, f, f, vp, vp]
    L.soc_sca_zero_out.argtypes = [vp, i, i, i]
    L.soc_sca_ps.argtypes = [vp, i, i, f, f, f, i, i, i, f, fp, i]
    L.soc_sca_pb.argtypes = [vp, i, i, i, f, f, f, f, i, i, i, f, fp, i]
    L.soc_sca_hp.argtypes = [vp, i, i, f, f, f, i, i, i, f, fp, i]
    L.soc_sca_cl.argtypes = [vp, i, i, i, f, f, f, i, i, i, f, fp, i]
    L.soc_get_counters.argtypes = [vp, C.POINTER(SocCounters)]
    L.soc_reset_counters.argtypes = [vp]
    L.soc_last_launch_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.soc_last_kernel.argtypes = [vp]
    L.soc_last_kernel.restype = C.c_char_p
    if path is None:
        _lib = L
    return L


def exported_symbols():
    return list(_EXPORTS)


def _f3(v):
    a = np.ascontiguousarray(np.asarray(v, np.float32).reshape(-1)[:3])
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


class Device:
    """One soc_context: a CUDA device + in-order stream.  Methods map 1:1 to the C entry points and raise
    SocError with soc_last_error() on a non-zero status."""

    def __init__(self, ordinal=0):
        self.L = load_library()
        self.ctx = C.c_void_p()
        self._ck(self.L.soc_create(int(ordinal), C.byref(self.ctx)))
        self.ordinal = int(ordinal)

    def _ck(self, status):
        if status != 0:
            raise SocError("soc_b200 status %d: %s" % (status, self.L.soc_last_error().decode()))

    def close(self):
        if self.ctx:
            self.L.soc_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._ck(self.L.soc_sync(self.ctx))

    def set_params(self, **kw):
        p = SocParams()
        p.bins, p.no_ps, p.ps_method = kw.get("bins", 2500), max(1, kw.get("no_ps", 1)), kw.get("ps_method", 0)
        p.with_abu, p.with_ali = kw.get("with_abu", 0), kw.get("with_ali", 0)
        p.noabsorbed, p.save_intensity = kw.get("noabsorbed", 1), kw.get("save_intensity", 0)
        p.use_emweight, p.hpbg_weighted = kw.get("use_emweight", 0), kw.get("hpbg_weighted", 0)
        p.ffs, p.step_weight, p.level_threshold = kw.get("ffs", 1), kw.get("step_weight", -1), kw.get("level_threshold", 0)
        p.with_msf, p.mirror, p.dir_weight = kw.get("with_msf", 0), kw.get("mirror", 0), kw.get("dir_weight", 0)
        roi_flags = kw.get("roi_flags", 1 * int(kw.get("with_roi_load", 0) > 0) + 2 * int(kw.get("with_roi_save", 0) > 0) +
                           4 * int(kw.get("roi_map", 0) > 0))
        p.do_split, p.roi_flags, p.map_interpolation = kw.get("do_split", 0), roi_flags, kw.get("map_interpolation", 0)
        p.sw_a, p.sw_b = kw.get("sw_a", 0.0), kw.get("sw_b", 0.0)
        p.length, p.factor, p.adhoc = kw["length"], kw.get("factor", 1.0e20), kw.get("adhoc", 1.0)
        p.ndust = kw.get("ndust", 1)
        p.opt_is_half = kw.get("opt_is_half", 0)
        p.ref_quirks = kw.get("ref_quirks", 0)
        self._ck(self.L.soc_set_params(self.ctx, C.byref(p)))
        self.params = p

    def set_grid(self, cloud):
        lc = np.ascontiguousarray(cloud.LCELLS, np.int32)
        of = np.ascontiguousarray(cloud.OFF, np.int32)
        de = np.ascontiguousarray(cloud.DENS, np.float32)
        self._ck(self.L.soc_set_grid(self.ctx, cloud.NX, cloud.NY, cloud.NZ, cloud.LEVELS, cloud.CELLS,
                                     lc.ctypes.data, of.ctypes.data, de.ctypes.data))
        self.sync()            # host arrays may go away after the call

    def set_rng_mode(self, mode):
        self._ck(self.L.soc_set_rng_mode(self.ctx, int(mode)))

    def set_shard(self, rank, world):
        self._ck(self.L.soc_set_shard(self.ctx, int(rank), int(world)))

    def set_tuning(self, deposit=DEP_TILE, refill=8, aggregate_steps=24):
        self._ck(self.L.soc_set_tuning(self.ctx, int(deposit), int(refill), int(aggregate_steps)))

    def set_geometry(self, mode):
        self._ck(self.L.soc_set_geometry(self.ctx, int(mode)))

    def set_roi(self, roi, roi_step=0, roi_nside=16, roi_dim=(1, 1, 1)):
        r = np.ascontiguousarray(roi, np.int32)
        d = np.ascontiguousarray(roi_dim, np.int32)
        self._ck(self.L.soc_set_roi(self.ctx, r.ctypes.data, int(roi_step), int(roi_nside), d.ctypes.data))

    def build_opt(self, kabs, ksca, first=0, single_abu=False):
        """OPT from the uploaded abundances and the per-species cross sections of one frequency (ASOC.py:1146-1161)."""
        a = np.ascontiguousarray(kabs, np.float32)
        s = np.ascontiguousarray(ksca, np.float32)
        self._ck(self.L.soc_build_opt(self.ctx, len(a), a.ctypes.data_as(C.POINTER(C.c_float)), s.ctypes.data_as(C.POINTER(C.c_float)),
                                      int(first), 1 if single_abu else 0))

    def split_absorbed(self, idust, rabs, cells):
        """Absorptions of species `idust` from the [cells, nfreq] array in buffer FABS and the abundances in buffer ABU
        (kernel_A2E_MABU_aux.c split_absorbed); rabs = [nfreq, ndust] float64."""
        r = np.ascontiguousarray(rabs, np.float64)
        out = np.empty((cells, r.shape[0]), np.float32)
        self._ck(self.L.soc_split_absorbed(self.ctx, int(idust), r.shape[1], r.shape[0], r.ctypes.data_as(C.POINTER(C.c_double)),
                                           out.ctypes.data))
        return out

    def set_layout(self, mode):
        self._ck(self.L.soc_set_layout(self.ctx, int(mode)))

    def set_domains(self, edge):
        self._ck(self.L.soc_set_domains(self.ctx, int(edge)))

    def upload(self, buf, array, dtype=np.float32):
        a = np.ascontiguousarray(array, dtype)
        # pageable host memory is staged before cudaMemcpyAsync returns; pinned arrays must stay alive until sync()
        self._ck(self.L.soc_upload(self.ctx, buf, a.ctypes.data, a.nbytes))
        return a

    def download(self, buf, n, dtype=np.float32, out=None):
        a = np.empty(n, dtype) if out is None else out
        self._ck(self.L.soc_download(self.ctx, buf, a.ctypes.data, a.nbytes))
        return a

    def clear(self, buf, nbytes):
        self._ck(self.L.soc_clear(self.ctx, buf, int(nbytes)))

    def device_ptr(self, buf):
        n = C.c_size_t()
        p = self.L.soc_device_ptr(self.ctx, buf, C.byref(n))
        return p, n.value

    def stream(self):
        return self.L.soc_stream(self.ctx)

    def zero_amc(self, tag):
        self._ck(self.L.soc_zero_amc(self.ctx, int(tag)))

    def sim_pb(self, source, packets, batch, seed, abs_, sca, bg, tw, global_):
        self._ck(self.L.soc_sim_pb(self.ctx, source, packets, batch, seed, abs_, sca, bg, tw, global_))

    def sim_hp(self, packets, batch, seed, abs_, sca, tw, global_):
        self._ck(self.L.soc_sim_hp(self.ctx, packets, batch, seed, abs_, sca, tw, global_))

    def sim_cl(self, source, packets, batch, seed, abs_, sca, tw, global_):
        self._ck(self.L.soc_sim_cl(self.ctx, source, packets, batch, seed, abs_, sca, tw, global_))

    def absorbed_begin(self, nfreq):
        self._ck(self.L.soc_absorbed_begin(self.ctx, int(nfreq)))

    def absorbed_add(self, ifreq):
        self._ck(self.L.soc_absorbed_add(self.ctx, int(ifreq)))

    def absorbed_finish(self, coeff0, nnnlimit, scale=True, out=None):
        self._ck(self.L.soc_absorbed_finish(self.ctx, coeff0, nnnlimit, 1 if scale else 0, None if out is None else out.ctypes.data))
        return out

    def eq_temperature(self, level, adhoc, kE, Emin, NE):
        self._ck(self.L.soc_eq_temperature(self.ctx, level, adhoc, kE, Emin, NE))

    def emission(self, freq, fabs_):
        self._ck(self.L.soc_emission(self.ctx, freq, fabs_))

    def mapping(self, map_dx, npx, npy, dir_, ra, de, abs_, sca, centre, intobs, save_colden):
        k = [_f3(v) for v in (dir_, ra, de, centre, intobs)]
        self._ck(self.L.soc_mapping(self.ctx, map_dx, npx, npy, k[0][1], k[1][1], k[2][1], abs_, sca, k[3][1], k[4][1],
                                    save_colden))

    def mapping_levels(self, map_dx, npx, npy, dir_, ra, de, abs_, sca, centre, intobs, save_colden):
        k = [_f3(v) for v in (dir_, ra, de, centre, intobs)]
        self._ck(self.L.soc_mapping_levels(self.ctx, map_dx, npx, npy, k[0][1], k[1][1], k[2][1], abs_, sca, k[3][1], k[4][1],
                                           save_colden))

    def healpix_mapping(self, nside, abs_, sca, intobs, save_colden):
        k = _f3(intobs)
        self._ck(self.L.soc_healpix_mapping(self.ctx, nside, abs_, sca, k[1], save_colden))

    def ps_tau(self, no, dir_, abs_, sca):
        k = _f3(dir_)
        col, tau = np.zeros(no, np.float32), np.zeros(no, np.float32)
        self._ck(self.L.soc_ps_tau(self.ctx, no, k[1], abs_, sca, col.ctypes.data, tau.ctypes.data))
        return col, tau

    def emission2(self, c0, c1, freq, fabs_, out=None):
        fr = np.ascontiguousarray(freq, np.float32)
        fa = np.ascontiguousarray(fabs_, np.float32)
        if out is None:
            out = np.empty((c1 - c0, len(fr)), np.float32)
        assert out.flags.c_contiguous and out.dtype == np.float32 and out.size == (c1 - c0) * len(fr)
        self._ck(self.L.soc_emission2(self.ctx, c0, c1, len(fr), fr.ctypes.data, fa.ctypes.data, out.ctypes.data))
        return out

    def sca_zero_out(self, ndir, npx, npy):
        self._ck(self.L.soc_sca_zero_out(self.ctx, ndir, npx, npy))

    def sca_ps(self, packets, batch, seed, abs_, sca, ndir, npx, npy, map_dx, centre, global_):
        k = _f3(centre)
        self._ck(self.L.soc_sca_ps(self.ctx, packets, batch, seed, abs_, sca, ndir, npx, npy, map_dx, k[1], global_))

    def sca_pb(self, source, packets, batch, seed, abs_, sca, bg, ndir, npx, npy, map_dx, centre, global_):
        k = _f3(centre)
        self._ck(self.L.soc_sca_pb(self.ctx, source, packets, batch, seed, abs_, sca, bg, ndir, npx, npy, map_dx, k[1],
                                   global_))

    def sca_hp(self, packets, batch, seed, abs_, sca, ndir, npx, npy, map_dx, centre, global_):
        k = _f3(centre)
        self._ck(self.L.soc_sca_hp(self.ctx, packets, batch, seed, abs_, sca, ndir, npx, npy, map_dx, k[1], global_))

    def sca_cl(self, source, packets, batch, seed, abs_, sca, ndir, npx, npy, map_dx, centre, global_):
        k = _f3(centre)
        self._ck(self.L.soc_sca_cl(self.ctx, source, packets, batch, seed, abs_, sca, ndir, npx, npy, map_dx, k[1],
                                   global_))

    def counters(self):
        c = SocCounters()
        self._ck(self.L.soc_get_counters(self.ctx, C.byref(c)))
        return c

    def reset_counters(self):
        self._ck(self.L.soc_reset_counters(self.ctx))

    def last_kernel(self):
        return (self.L.soc_last_kernel(self.ctx) or b"").decode()

    def last_launch_ms(self):
        ms = C.c_float()
        self._ck(self.L.soc_last_launch_ms(self.ctx, C.byref(ms)))
        return ms.value


class Backend:
    """Grid + parameter block + kernels with numpy in/out, the same call shape as the reference kernels driven
    from ASOC.py / ASOCS.py (tests run the same seeded cases on this class and on the CPU checker and compare).
    Keyword `opts` are the former -D macros (with_abu, noabsorbed, save_intensity, ...)."""

    def __init__(self, cloud, gl=0.01, bins=2500, ordinal=0, rng_mode=RNG_PACKET, **opts):
        self.dev = Device(ordinal)
        self.cloud = cloud
        self.opts = dict(opts)
        length = float("%.5e" % (gl * 3.08567758e+18))           # -D LENGTH=%.5ef (ASOC.py:347,356)
        self.dev.set_params(bins=bins, length=length, **opts)
        self.dev.set_grid(cloud)
        if self.dev.params.roi_flags:
            self.dev.set_roi(opts.get("roi", [0] * 6), opts.get("roi_step", 0), opts.get("roi_nside", 16), opts.get("roi_dim", (1, 1, 1)))
        self.dev.set_rng_mode(rng_mode)
        self.bins = bins
        self.n = cloud.CELLS
        self._host = {}
        self.use_int = opts.get("noabsorbed", 1) == 0 or opts.get("save_intensity", 0) in (1, 2)
        self.save2 = opts.get("save_intensity", 0) == 2
        self.with_ali = opts.get("with_ali", 0) > 0
        self.zero(0)
        self.zero(1)

    # accumulators are read back on access, like the enqueue_copy calls of ASOC.py:1484,1533
    def _get(self, buf):
        return self.dev.download(buf, self.n)

    @property
    def tabs(self):
        return self._get(BUF_TABS)

    @property
    def xab(self):
        return self._get(BUF_XAB) if self.with_ali else np.zeros(self.n, np.float32)

    @property
    def int_(self):
        return self._get(BUF_INT) if self.use_int else np.zeros(self.n, np.float32)

    @property
    def intx(self):
        return self._get(BUF_INTX) if self.save2 else np.zeros(self.n, np.float32)

    @property
    def inty(self):
        return self._get(BUF_INTY) if self.save2 else np.zeros(self.n, np.float32)

    @property
    def intz(self):
        return self._get(BUF_INTZ) if self.save2 else np.zeros(self.n, np.float32)

    def clear_roi_save(self):
        ptr, nbytes = self.dev.device_ptr(BUF_ROI_SAVE)
        if nbytes:
            self.dev.clear(BUF_ROI_SAVE, nbytes)

    @property
    def roi_save(self):
        ptr, nbytes = self.dev.device_ptr(BUF_ROI_SAVE)
        return self.dev.download(BUF_ROI_SAVE, nbytes // 4)

    @property
    def counters(self):
        return self.dev.counters()

    def zero(self, tag):
        self.dev.zero_amc(tag)

    def _put(self, **bufs):
        table = dict(dsc=(BUF_DSC, np.float32), csc=(BUF_CSC, np.float32), emit=(BUF_EMIT, np.float32),
                     emwei=(BUF_EMWEI, np.float32), opt=(BUF_OPT, np.float32), pspos=(BUF_PSPOS, np.float32),
                     ps=(BUF_PS, np.float32), xps_nside=(BUF_XPS_NSIDE, np.int32), xps_side=(BUF_XPS_SIDE, np.int32),
                     xps_area=(BUF_XPS_AREA, np.float32), hpbg=(BUF_HPBG, np.float32), hpbgp=(BUF_HPBGP, np.float32),
                     abu=(BUF_ABU, np.float32), abs_v=(BUF_ABSV, np.float32), sca_v=(BUF_SCAV, np.float32),
                     roi_load=(BUF_ROI_LOAD, np.float32))
        for k, v in bufs.items():
            if v is None:
                continue
            b, dt = table[k]
            a = np.ascontiguousarray(np.asarray(v, dt).reshape(-1))
            if k in ("xps_side", "xps_area") and a.size % 3:
                a = np.concatenate([a, np.zeros(3 - a.size % 3, dt)])
            self.dev.upload(b, a, dt)
        self.dev.sync()

    def sim_pb(self, global_, source, packets, batch, seed, bg, tw, abs_=0.0, sca=0.0, **bufs):
        self._put(**bufs)
        self.dev.sim_pb(source, packets, batch, seed, abs_, sca, bg, tw, global_)

    def sim_hp(self, global_, packets, batch, seed, tw, abs_=0.0, sca=0.0, **bufs):
        self._put(**bufs)
        self.dev.sim_hp(packets, batch, seed, abs_, sca, tw, global_)

    def sim_cl(self, global_, packets, batch, seed, tw, abs_=0.0, sca=0.0, **bufs):
        self._put(**bufs)
        self.dev.sim_cl(2, packets, batch, seed, abs_, sca, tw, global_)

    def eq_temperature(self, level, adhoc, kE, Emin, NE, ttt, emit, tnew):
        self.dev.upload(BUF_TTT, ttt)
        self.dev.upload(BUF_EMIT, emit)
        self.dev.sync()
        self.dev.eq_temperature(level, adhoc, kE, Emin, NE)
        sl = self.cloud.level_slice(level)
        tnew[sl] = self.dev.download(BUF_TNEW, self.n)[sl]

    def emission(self, freq, fabs_, t):
        self.dev.upload(BUF_TNEW, t)
        self.dev.sync()
        self.dev.emission(freq, fabs_)
        return self.dev.download(BUF_EMIT, self.n)

    def emission2(self, c0, c1, freq, fabs_, t):
        self.dev.upload(BUF_TNEW, t)
        self.dev.sync()
        return self.dev.emission2(c0, c1, freq, fabs_)

    def mapping(self, map_dx, npx, npy, emit, dir_, ra, de, abs_, sca, centre, intobs=(-1e12, 0, 0), opt=None,
                save_colden=0):
        self._put(emit=emit, opt=opt)
        self.dev.mapping(map_dx, npx, npy, dir_, ra, de, abs_, sca, centre, intobs, save_colden)
        m = self.dev.download(BUF_MAP, npx * npy)
        t = self.dev.download(BUF_SAVETAU, npx * npy)
        return m.reshape(npy, npx), t.reshape(npy, npx)

    def mapping_levels(self, map_dx, npx, npy, emit, dir_, ra, de, abs_, sca, centre, intobs=(-1e12, 0, 0), opt=None,
                       colden=False):
        self._put(emit=emit, opt=opt)
        self.dev.mapping_levels(map_dx, npx, npy, dir_, ra, de, abs_, sca, centre, intobs, 1 if colden else 0)
        levels = int(self.cloud.LEVELS)
        m = self.dev.download(BUF_MAP, levels * npx * npy).reshape(levels, npy, npx)
        return (m, self.dev.download(BUF_SAVETAU, npx * npy).reshape(npy, npx)) if colden else m

    def healpix_mapping(self, nside, emit, abs_, sca, intobs, opt=None, save_colden=0):
        self._put(emit=emit, opt=opt)
        self.dev.healpix_mapping(nside, abs_, sca, intobs, save_colden)
        n = 12 * nside * nside
        return self.dev.download(BUF_MAP, n), self.dev.download(BUF_SAVETAU, n)

    def ps_tau(self, pspos, dir_, abs_, sca, opt=None):
        self._put(pspos=pspos, opt=opt)
        return self.dev.ps_tau(len(np.asarray(pspos).reshape(-1)) // 3, dir_, abs_, sca)

    def _observers(self, odirs, ora, ode):
        for b, v in ((BUF_ODIR, odirs), (BUF_ORA, ora), (BUF_ODE, ode)):
            self.dev.upload(b, np.ascontiguousarray(np.asarray(v, np.float32)[:, :3].reshape(-1)))
        self.dev.sync()

    def _sca_out(self, ndir, npx, npy):
        if ndir < 0:
            return self.dev.download(BUF_OUT, 12 * ndir * ndir)
        return self.dev.download(BUF_OUT, ndir * npx * npy).reshape(ndir, npy, npx)

    def sca_ps(self, global_, packets, batch, seed, ndir, npx, npy, map_dx, centre, odirs, ora, ode, abs_=0.0, sca=0.0,
               **bufs):
        self._put(**bufs)
        self._observers(odirs, ora, ode)
        self.dev.sca_zero_out(ndir, npx, npy)
        self.dev.sca_ps(packets, batch, seed, abs_, sca, ndir, npx, npy, map_dx, centre, global_)
        return self._sca_out(ndir, npx, npy)

    def sca_pb(self, global_, source, packets, batch, seed, bg, ndir, npx, npy, map_dx, centre, odirs, ora, ode,
               abs_=0.0, sca=0.0, **bufs):
        self._put(**bufs)
        self._observers(odirs, ora, ode)
        self.dev.sca_zero_out(ndir, npx, npy)
        self.dev.sca_pb(source, packets, batch, seed, abs_, sca, bg, ndir, npx, npy, map_dx, centre, global_)
        return self._sca_out(ndir, npx, npy)

    def sca_hp(self, global_, packets, batch, seed, ndir, npx, npy, map_dx, centre, odirs, ora, ode, abs_=0.0, sca=0.0,
               **bufs):
        self._put(**bufs)
        self._observers(odirs, ora, ode)
        self.dev.sca_zero_out(ndir, npx, npy)
        self.dev.sca_hp(packets, batch, seed, abs_, sca, ndir, npx, npy, map_dx, centre, global_)
        return self._sca_out(ndir, npx, npy)

    def sca_cl(self, global_, packets, batch, seed, ndir, npx, npy, map_dx, centre, odirs, ora, ode, abs_=0.0, sca=0.0,
               **bufs):
        self._put(**bufs)
        self._observers(odirs, ora, ode)
        self.dev.sca_zero_out(ndir, npx, npy)
        self.dev.sca_cl(2, packets, batch, seed, abs_, sca, ndir, npx, npy, map_dx, centre, global_)
        return self._sca_out(ndir, npx, npy)

    def close(self):
        self.dev.close()
