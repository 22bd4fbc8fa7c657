"""TEST INFRASTRUCTURE ONLY: an object with the interface of soc_b200.backend.Device whose "kernels" are the
CPU oracle (oracle/orc.py).  The drivers in soc_b200/ take a device factory, so the same driver code can be
run end to end on the CUDA library and on the oracle and the output files compared -- and the host logic
(ini parsing, weights, file formats, rank sharding) can be tested on machines without a GPU.

Sharding: with world > 1 every rank runs the complete launch with a rank-dependent seed and contributes 1/world
of it, so that the sum over ranks is a valid Monte Carlo estimate like the packet-sharded CUDA path.
"""
import ctypes as C

import numpy as np

from oracle import orc
from soc_b200 import backend as bk


class _Counters:
    def __init__(self, c):
        self.packets, self.steps, self.scatterings, self.peels = c.packets, c.steps, c.scatterings, c.peels
        self.launches, self.reserved = 0, [0, 0, 0]


class OracleDevice:
    def __init__(self, ordinal=0):
        self.O = None
        self.params = {}
        self.buf = {}
        self.rank, self.world = 0, 1
        self.cloud = None
        self.roi = {}

    # ---- configuration ------------------------------------------------------------------------------------
    def set_params(self, **kw):
        if kw.get("do_split") or kw.get("roi_flags") or kw.get("dir_weight"):
            raise bk.SocError("unsupported option")
        self.params = dict(kw)
        if self.cloud is not None:
            self._make()

    def set_grid(self, cloud):
        self.cloud = cloud
        self._make()

    def _make(self):
        p = dict(self.params)
        length = p.pop("length")
        bins = p.pop("bins", 2500)
        for k in ("factor", "adhoc", "dir_weight", "do_split", "roi_flags", "opt_is_half"):
            p.pop(k, None)
        q = p.pop("ref_quirks", 0)
        p["hg_test"], p["maph_literal"] = int(bool(q & 1)), int(bool(q & 2))
        # the drivers are compared with the library's production kernels: exact mirror / scattering position
        p.update(self.roi)
        self.O = orc.Oracle(self.cloud, gl=0.01, bins=bins, mirror_exact=1, sca_exact_level=1, **p)
        self.O.P.length = length
        self.n = self.cloud.CELLS

    def set_roi(self, roi, roi_step=0, roi_nside=16, roi_dim=(1, 1, 1)):
        self.roi = dict(roi=[int(v) for v in roi], roi_step=int(roi_step), roi_nside=int(roi_nside), roi_dim=[int(v) for v in roi_dim])
        self._make()

    def set_rng_mode(self, mode):
        pass

    def set_shard(self, rank, world):
        self.rank, self.world = rank, world

    def set_tuning(self, *a, **k):
        pass

    def set_geometry(self, mode):
        pass

    def set_layout(self, mode):
        pass

    def sync(self):
        pass

    def close(self):
        pass

    def stream(self):
        return None

    # ---- buffers --------------------------------------------------------------------------------------------
    _ACC = {bk.BUF_TABS: "tabs", bk.BUF_XAB: "xab", bk.BUF_INT: "int_", bk.BUF_INTX: "intx", bk.BUF_INTY: "inty",
            bk.BUF_INTZ: "intz"}

    def host_view(self, b, count):
        if b in self._ACC:
            return getattr(self.O, self._ACC[b])[:count]
        if b == bk.BUF_ROI_SAVE:
            return self.O.roi_save[:count]
        return self.buf[b][:count]

    def upload(self, b, array, dtype=np.float32):
        if b in self._ACC:
            getattr(self.O, self._ACC[b])[:] = np.asarray(array, np.float32)
        else:
            self.buf[b] = np.array(array, dtype).reshape(-1).astype(np.float32 if dtype == np.float16 else dtype)
        return array

    def build_opt(self, kabs, ksca, first=0, single_abu=False):
        """Host restatement of ASOC.py:1146-1161 on the uploaded abundances (the library does this on the device)."""
        kabs, ksca = np.asarray(kabs, np.float32), np.asarray(ksca, np.float32)
        abu = self.buf[bk.BUF_ABU].reshape(self.n, -1)
        opt = np.zeros((self.n, 2), np.float32)
        if single_abu:
            a = abu[:, 0]
            opt[:, 0] = a * kabs[0] + (1.0 - a) * kabs[1]
            opt[:, 1] = a * ksca[0] + (1.0 - a) * ksca[1]
        else:
            for d in range(first, len(kabs)):
                opt[:, 0] += abu[:, d] * kabs[d]
                opt[:, 1] += abu[:, d] * ksca[d]
        if self.params.get("opt_is_half"):
            opt = opt.astype(np.float16).astype(np.float32)
        self.buf[bk.BUF_OPT] = opt.reshape(-1)

    def split_absorbed(self, idust, rabs, cells):
        rabs = np.ascontiguousarray(rabs, np.float64)
        return orc.split_absorbed(idust, rabs, self.buf[bk.BUF_ABU].reshape(cells, -1), self.buf[bk.BUF_FABS].reshape(cells, -1))

    def download(self, b, n, dtype=np.float32, out=None):
        src = self.host_view(b, n)
        if out is None:
            return np.array(src[:n], dtype)
        out[:] = src[:n]
        return out

    def clear(self, b, nbytes):
        if b == bk.BUF_ROI_SAVE:
            self.O.roi_save[:] = 0.0
            return
        self.buf[b] = np.zeros(nbytes // 4, np.float32)

    def device_ptr(self, b):
        if b in self._ACC:
            return 1, 4 * self.n
        a = self.buf.get(b)
        return (1, a.nbytes) if a is not None else (None, 0)

    def zero_amc(self, tag):
        self.O.zero(tag)

    # ---- launches -------------------------------------------------------------------------------------------
    def _g(self, b):
        return self.buf.get(b)

    def _msf(self):
        return dict(abu=self._g(bk.BUF_ABU), abs_v=self._g(bk.BUF_ABSV), sca_v=self._g(bk.BUF_SCAV))

    def _sharded(self, run, seed):
        """Run a launch; with world > 1 every rank contributes 1/world of a full launch with its own seed."""
        if self.world == 1:
            run(seed)
            return
        keep = {k: getattr(self.O, k).copy() for k in self._ACC.values()}
        for k in self._ACC.values():
            getattr(self.O, k)[:] = 0.0
        run(float(np.fmod(seed + 0.37 * self.rank + 0.011, 1.0)))
        for k in self._ACC.values():
            a = getattr(self.O, k)
            a[:] = keep[k] + a / np.float32(self.world)

    def sim_pb(self, source, packets, batch, seed, abs_, sca, bg, tw, global_):
        self._sharded(lambda s: self.O.sim_pb(global_, source, packets, batch, s, bg, tw, abs_=abs_, sca=sca,
                                              dsc=self._g(bk.BUF_DSC), csc=self._g(bk.BUF_CSC), opt=self._g(bk.BUF_OPT),
                                              pspos=self._g(bk.BUF_PSPOS), ps=self._g(bk.BUF_PS),
                                              xps_nside=self._g(bk.BUF_XPS_NSIDE), xps_side=self._g(bk.BUF_XPS_SIDE),
                                              xps_area=self._g(bk.BUF_XPS_AREA), roi_load=self._g(bk.BUF_ROI_LOAD), **self._msf()), seed)

    def sim_hp(self, packets, batch, seed, abs_, sca, tw, global_):
        self._sharded(lambda s: self.O.sim_hp(global_, packets, batch, s, tw, abs_=abs_, sca=sca, dsc=self._g(bk.BUF_DSC),
                                              csc=self._g(bk.BUF_CSC), opt=self._g(bk.BUF_OPT), hpbg=self._g(bk.BUF_HPBG),
                                              hpbgp=self._g(bk.BUF_HPBGP), **self._msf()), seed)

    def sim_cl(self, source, packets, batch, seed, abs_, sca, tw, global_):
        self._sharded(lambda s: self.O.sim_cl(global_, packets, batch, s, tw, abs_=abs_, sca=sca, dsc=self._g(bk.BUF_DSC),
                                              csc=self._g(bk.BUF_CSC), opt=self._g(bk.BUF_OPT), emit=self._g(bk.BUF_EMIT),
                                              emwei=self._g(bk.BUF_EMWEI), **self._msf()), seed)

    def absorbed_begin(self, nfreq):
        self.fabs = np.zeros((self.n, nfreq), np.float32)
        self.buf[bk.BUF_FABS] = self.fabs.reshape(-1)

    def absorbed_add(self, ifreq):
        self.fabs[:, ifreq] += self.O.int_

    def absorbed_finish(self, coeff0, nnnlimit, scale=True, out=None):
        if scale:
            c = self.cloud
            for level in range(c.LEVELS):
                sl = c.level_slice(level)
                with np.errstate(all='ignore'):
                    self.fabs[sl] *= (np.float32(coeff0) * np.float32(8.0 ** level) / c.DENS[sl]).reshape(-1, 1)
                self.fabs[sl][c.DENS[sl] <= nnnlimit] = -1.0e20
        if out is not None:
            out[...] = self.fabs.reshape(out.shape)
        return out

    def eq_temperature(self, level, adhoc, kE, Emin, NE):
        if bk.BUF_TNEW not in self.buf or self.buf[bk.BUF_TNEW].size != self.n:
            self.buf[bk.BUF_TNEW] = np.zeros(self.n, np.float32)
        self.O.eq_temperature(level, adhoc, kE, Emin, NE, self.buf[bk.BUF_TTT], self.buf[bk.BUF_EMIT], self.buf[bk.BUF_TNEW])

    def emission(self, freq, fabs_):
        self.buf[bk.BUF_EMIT] = self.O.emission(freq, fabs_, self.buf[bk.BUF_TNEW])

    def emission2(self, c0, c1, freq, fabs_, out=None):
        e = self.O.emission2(c0, c1, freq, fabs_, self.buf[bk.BUF_TNEW])
        if out is None:
            return e
        out[...] = e
        return out

    def mapping(self, map_dx, npx, npy, dir_, ra, de, abs_, sca, centre, intobs, save_colden):
        m, t = self.O.mapping(map_dx, npx, npy, self.buf[bk.BUF_EMIT], dir_, ra, de, abs_, sca, centre, intobs=intobs,
                              opt=self._g(bk.BUF_OPT), save_colden=save_colden)
        self.buf[bk.BUF_MAP], self.buf[bk.BUF_SAVETAU] = m.reshape(-1), t.reshape(-1)

    def mapping_levels(self, map_dx, npx, npy, dir_, ra, de, abs_, sca, centre, intobs, save_colden):
        r = self.O.mapping_levels(map_dx, npx, npy, self.buf[bk.BUF_EMIT], dir_, ra, de, abs_, sca, centre, intobs=intobs,
                                  opt=self._g(bk.BUF_OPT), colden=bool(save_colden))
        if save_colden:
            self.buf[bk.BUF_MAP], self.buf[bk.BUF_SAVETAU] = r[0].reshape(-1), r[1].reshape(-1)
        else:
            self.buf[bk.BUF_MAP] = r.reshape(-1)

    def healpix_mapping(self, nside, abs_, sca, intobs, save_colden):
        m, t = self.O.healpix_mapping(nside, self.buf[bk.BUF_EMIT], abs_, sca, intobs, opt=self._g(bk.BUF_OPT),
                                      save_colden=save_colden)
        self.buf[bk.BUF_MAP], self.buf[bk.BUF_SAVETAU] = m, t

    def ps_tau(self, no, dir_, abs_, sca):
        return self.O.ps_tau(self.buf[bk.BUF_PSPOS][:3 * no], dir_, abs_, sca, opt=self._g(bk.BUF_OPT))

    def sca_zero_out(self, ndir, npx, npy):
        self.buf[bk.BUF_OUT] = np.zeros(ndir * npx * npy if ndir > 0 else 12 * ndir * ndir, np.float32)

    def _obs(self, ndir):
        if ndir < 0:            # Healpix observer: ODIR holds the position, ORA / ODE are not used
            z = np.zeros((1, 3), np.float32)
            return self.buf[bk.BUF_ODIR][:3].reshape(1, 3), z, z
        return [self.buf[b].reshape(ndir, 3) for b in (bk.BUF_ODIR, bk.BUF_ORA, bk.BUF_ODE)]

    def _sca(self, fun, lead, seed, ndir, npx, npy, map_dx, centre, **bufs):
        od, ra, de = self._obs(ndir)
        if self.world > 1:
            seed = float(np.fmod(seed + 0.37 * self.rank + 0.011, 1.0))
        out = fun(*lead, seed, ndir, npx, npy, map_dx, centre, od, ra, de, dsc=self._g(bk.BUF_DSC), csc=self._g(bk.BUF_CSC),
                  opt=self._g(bk.BUF_OPT), **self._msf(), **bufs)
        self.buf[bk.BUF_OUT] = self.buf[bk.BUF_OUT] + out.reshape(-1) / np.float32(self.world)

    def sca_ps(self, packets, batch, seed, abs_, sca, ndir, npx, npy, map_dx, centre, global_):
        self._sca(self.O.sca_ps, (global_, packets, batch), seed, ndir, npx, npy, map_dx, centre, abs_=abs_, sca=sca,
                  pspos=self._g(bk.BUF_PSPOS), ps=self._g(bk.BUF_PS))

    def sca_pb(self, source, packets, batch, seed, abs_, sca, bg, ndir, npx, npy, map_dx, centre, global_):
        od, ra, de = self._obs(ndir)
        if self.world > 1:
            seed = float(np.fmod(seed + 0.37 * self.rank + 0.011, 1.0))
        out = self.O.sca_pb(global_, source, packets, batch, seed, bg, ndir, npx, npy, map_dx, centre, od, ra, de, abs_=abs_,
                            sca=sca, dsc=self._g(bk.BUF_DSC), csc=self._g(bk.BUF_CSC), opt=self._g(bk.BUF_OPT),
                            pspos=self._g(bk.BUF_PSPOS), ps=self._g(bk.BUF_PS), roi_load=self._g(bk.BUF_ROI_LOAD), **self._msf())
        self.buf[bk.BUF_OUT] = self.buf[bk.BUF_OUT] + out.reshape(-1) / np.float32(self.world)

    def sca_hp(self, packets, batch, seed, abs_, sca, ndir, npx, npy, map_dx, centre, global_):
        self._sca(self.O.sca_hp, (global_, packets, batch), seed, ndir, npx, npy, map_dx, centre, abs_=abs_, sca=sca,
                  hpbg=self._g(bk.BUF_HPBG), hpbgp=self._g(bk.BUF_HPBGP))

    def sca_cl(self, source, packets, batch, seed, abs_, sca, ndir, npx, npy, map_dx, centre, global_):
        self._sca(self.O.sca_cl, (global_, packets, batch), seed, ndir, npx, npy, map_dx, centre, abs_=abs_, sca=sca,
                  emit=self._g(bk.BUF_EMIT), emwei=self._g(bk.BUF_EMWEI))

    def counters(self):
        return _Counters(self.O.counters)

    def reset_counters(self):
        C.memset(C.byref(self.O.counters), 0, C.sizeof(self.O.counters))

    def last_launch_ms(self):
        return 0.0
