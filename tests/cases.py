"""Small seeded cases of the hot path, written against the common call shape shared by
oracle.orc.Oracle, oracle.ref.Reference and soc_b200.backend.Backend, so that the same case can be run
on the reference kernels (CPU shim), the plain-C oracle and the CUDA library and compared.

Each case is   name -> (make_cloud, opts, run)   where run(X) returns {output name: ndarray}.
"""
import numpy as np

from soc_b200 import synth
from soc_b200.formats import Cloud
from soc_b200.hostmath import observer_directions

BINS = 2500
DSC6, CSC6 = synth.hg_tables(0.6, BINS)
DSC0, CSC0 = synth.hg_tables(0.0, BINS)


def _reg(n):
    return lambda: synth.regular_cloud(n)


def _oct(nroot, levels, frac=0.2, seed=3):
    return lambda: synth.octree_cloud(nroot, levels, refine_fraction=frac, seed=seed)


def _tau_scale(cloud, tau):
    """ABS/SCA per unit density so that the mean optical depth across the root grid is ~tau."""
    m = cloud.DENS[:cloud.NX * cloud.NY * cloud.NZ]
    mean = float(np.mean(np.where(m > 0, m, 1.0)))
    return tau / (cloud.NX * mean)


def run_bg(batch=2, tau_a=1.5, tau_s=2.5, seed=0.4, g=True):
    def run(X):
        c = X.cloud
        glob = 8 * c.AREA
        k = _tau_scale(c, 1.0)
        X.zero(0)
        X.zero(1)
        X.sim_pb(glob, 1, glob * batch, batch, seed, 1.0, 0.7, abs_=tau_a * k, sca=tau_s * k,
                 dsc=DSC6 if g else DSC0, csc=CSC6 if g else CSC0)
        return dict(tabs=X.tabs.copy(), int=X.int_.copy())
    return run


def run_ps(pspos, batch=6, glob=2048, tau_a=2.0, tau_s=2.0, seed=0.31, **extra):
    def run(X):
        c = X.cloud
        k = _tau_scale(c, 1.0)
        ps = np.linspace(1.0, 2.0, len(pspos)).astype(np.float32)
        X.zero(0)
        X.zero(1)
        X.sim_pb(glob, 0, glob * batch, batch * len(pspos), seed, 0.0, 1.3, abs_=tau_a * k, sca=tau_s * k,
                 dsc=DSC6, csc=CSC6, pspos=np.asarray(pspos, np.float32).reshape(-1), ps=ps, **extra)
        return dict(tabs=X.tabs.copy(), int=X.int_.copy())
    return run


def run_abu(batch=2, seed=0.77):
    def run(X):
        c = X.cloud
        glob = 8 * c.AREA
        k = _tau_scale(c, 1.0)
        rng = np.random.default_rng(5)
        opt = np.empty((c.CELLS, 2), np.float32)
        opt[:, 0] = k * (1.0 + rng.random(c.CELLS))
        opt[:, 1] = k * (1.5 + 2 * rng.random(c.CELLS))
        X.zero(0)
        X.zero(1)
        X.sim_pb(glob, 1, glob * batch, batch, seed, 2.0, 1.0, dsc=DSC6, csc=CSC6, opt=opt.reshape(-1))
        return dict(tabs=X.tabs.copy(), int=X.int_.copy())
    return run


def run_hp(weighted, batch=3, seed=0.52):
    def run(X):
        c = X.cloud
        k = _tau_scale(c, 1.0)
        rng = np.random.default_rng(11)
        sky = (0.2 + rng.random(49152)).astype(np.float32)
        sky[20000:20400] *= 30.0
        hpbgp = None
        if weighted:
            p = sky.astype(np.float64) / sky.mean()
            p = np.clip(p, 1e-3, 1e4)
            p /= p.sum()
            w = (1.0 / 49152.0) / p
            hpbgp = np.cumsum(p)
            hpbgp[-1] = 1.00001
            sky = (sky * w).astype(np.float32)
            hpbgp = hpbgp.astype(np.float32)
        glob = 1024
        X.zero(0)
        X.zero(1)
        X.sim_hp(glob, glob * batch, batch, seed, 0.9, abs_=1.2 * k, sca=2.0 * k, dsc=DSC6, csc=CSC6, hpbg=sky,
                 hpbgp=hpbgp)
        return dict(tabs=X.tabs.copy())
    return run


def run_cl(emweight, batch=2, glob=512, seed=0.13):
    def run(X):
        c = X.cloud
        k = _tau_scale(c, 1.0)
        rng = np.random.default_rng(7)
        emit = np.where(c.DENS > 0, c.DENS * (0.5 + rng.random(c.CELLS)), 0.0).astype(np.float32)
        emwei = None
        if emweight:
            emwei = (3.0 * rng.random(c.CELLS)).astype(np.float32)
            emwei[::7] = 0.0
        X.zero(0)
        X.zero(1)
        X.sim_cl(glob, c.CELLS * batch, batch, seed, 1.1, abs_=2.0 * k, sca=1.5 * k, dsc=DSC6, csc=CSC6, emit=emit,
                 emwei=emwei)
        return dict(tabs=X.tabs.copy(), xab=X.xab.copy(), int=X.int_.copy())
    return run


def run_map(npix, dirs, map_dx=1.0, intobs=None, colden=0, abu=False, centre_off=0.0):
    def run(X):
        c = X.cloud
        k = _tau_scale(c, 1.0)
        rng = np.random.default_rng(9)
        emit = np.where(c.DENS > 0, 1.0 + rng.random(c.CELLS), 0.0).astype(np.float32)
        opt = None
        if abu:
            opt = np.empty((c.CELLS, 2), np.float32)
            opt[:, 0] = k * (0.5 + rng.random(c.CELLS))
            opt[:, 1] = k * (0.5 + rng.random(c.CELLS))
            opt = opt.reshape(-1)
        out = {}
        _, od, ra, de = observer_directions([d[0] for d in dirs], [d[1] for d in dirs])
        centre = np.array([0.5 * c.NX + centre_off, 0.5 * c.NY, 0.5 * c.NZ - centre_off], np.float32)
        for i in range(len(dirs)):
            io = (-1e12, 0.0, 0.0) if intobs is None else intobs
            m, t = X.mapping(map_dx, npix[0], npix[1], emit, od[i], ra[i], de[i], 1.2 * k, 0.8 * k, centre,
                             intobs=io, opt=opt, save_colden=colden)
            out["map%d" % i] = m.copy()
            out["tau%d" % i] = t.copy()
        return out
    return run


def run_maplev(npix, dirs, map_dx=1.0, intobs=None, colden=False, abu=False, centre_off=0.0):
    """Per-level images (kernel_ASOC_map_H.c Mapping)."""
    def run(X):
        c = X.cloud
        k = _tau_scale(c, 1.0)
        rng = np.random.default_rng(9)
        emit = np.where(c.DENS > 0, 1.0 + rng.random(c.CELLS), 0.0).astype(np.float32)
        opt = None
        if abu:
            opt = np.empty((c.CELLS, 2), np.float32)
            opt[:, 0] = k * (0.5 + rng.random(c.CELLS))
            opt[:, 1] = k * (0.5 + rng.random(c.CELLS))
            opt = opt.reshape(-1)
        out = {}
        _, od, ra, de = observer_directions([d[0] for d in dirs], [d[1] for d in dirs])
        centre = np.array([0.5 * c.NX + centre_off, 0.5 * c.NY, 0.5 * c.NZ - centre_off], np.float32)
        for i in range(len(dirs)):
            io = (-1e12, 0.0, 0.0) if intobs is None else intobs
            r = X.mapping_levels(map_dx, npix[0], npix[1], emit, od[i], ra[i], de[i], 1.2 * k, 0.8 * k, centre,
                                 intobs=io, opt=opt, colden=colden)
            if colden:
                out["map%d" % i], out["colden%d" % i] = r[0].copy(), r[1].copy()
            else:
                out["map%d" % i] = r.copy()
        return out
    return run


def run_hpmap(nside, intobs):
    def run(X):
        c = X.cloud
        k = _tau_scale(c, 1.0)
        rng = np.random.default_rng(9)
        emit = np.where(c.DENS > 0, 1.0 + rng.random(c.CELLS), 0.0).astype(np.float32)
        m, t = X.healpix_mapping(nside, emit, 1.0 * k, 1.0 * k, np.asarray(intobs, np.float32))
        return dict(map=m.copy(), tau=t.copy())
    return run


def run_pstau(pspos, dirs, abu=False):
    def run(X):
        c = X.cloud
        k = _tau_scale(c, 1.0)
        rng = np.random.default_rng(9)
        opt = None
        if abu:
            opt = np.empty((c.CELLS, 2), np.float32)
            opt[:, 0] = k * (0.5 + rng.random(c.CELLS))
            opt[:, 1] = k * (0.5 + rng.random(c.CELLS))
            opt = opt.reshape(-1)
        _, od, ra, de = observer_directions([d[0] for d in dirs], [d[1] for d in dirs])
        out = {}
        for i in range(len(dirs)):
            col, tau = X.ps_tau(np.asarray(pspos, np.float32).reshape(-1), od[i], 1.2 * k, 0.8 * k, opt=opt)
            out["colden%d" % i], out["tau%d" % i] = col.copy(), tau.copy()
        return out
    return run


def _msf_inputs(c, k, ndust=2, seed=21):
    """Two dust species with different scattering functions: ABU[cells, ndust], ABS/SCA[ndust], OPT = sums."""
    rng = np.random.default_rng(seed)
    abu = (0.2 + rng.random((c.CELLS, ndust))).astype(np.float32)
    abs_v = (k * np.linspace(0.8, 1.6, ndust)).astype(np.float32)
    sca_v = (k * np.linspace(2.5, 1.0, ndust)).astype(np.float32)
    opt = np.empty((c.CELLS, 2), np.float32)
    opt[:, 0] = (abu * abs_v).sum(axis=1)
    opt[:, 1] = (abu * sca_v).sum(axis=1)
    gs = np.linspace(0.7, -0.2, ndust)
    tabs = [synth.hg_tables(g, BINS) for g in gs]
    dsc = np.concatenate([t[0] for t in tabs]).astype(np.float32)
    csc = np.concatenate([t[1] for t in tabs]).astype(np.float32)
    return dict(abu=abu.reshape(-1), abs_v=abs_v, sca_v=sca_v, opt=opt.reshape(-1), dsc=dsc, csc=csc)


def run_bg_msf(batch=2, seed=0.44):
    def run(X):
        c = X.cloud
        glob = 8 * c.AREA
        k = _tau_scale(c, 1.0)
        X.zero(0)
        X.zero(1)
        X.sim_pb(glob, 1, glob * batch, batch, seed, 1.0, 0.7, **_msf_inputs(c, k))
        return dict(tabs=X.tabs.copy(), int=X.int_.copy())
    return run


def with_roi_save(run):
    """Wrap a Monte Carlo case: also return the photons recorded on entering the region of interest."""
    def wrapped(X):
        X.clear_roi_save()
        out = run(X)
        out["roi_save"] = np.array(X.roi_save, np.float32).copy()
        return out
    return wrapped


def run_roi_load(roi_dim, nside, rounds=2, seed=0.27, tau_s=2.0):
    """SOURCE == 3: packets from a stored external field, one Healpix map of directions per surface element."""
    def run(X):
        c = X.cloud
        k = _tau_scale(c, 1.0)
        nelem = roi_dim[0] * roi_dim[1] + roi_dim[1] * roi_dim[2] + roi_dim[2] * roi_dim[0]
        npix = 12 * nside * nside
        rng = np.random.default_rng(13)
        field = rng.random(nelem * npix).astype(np.float32)
        field[::5] = 0.0                                   # empty directions are skipped without a packet
        X.zero(0)
        X.zero(1)
        X.sim_pb(100 * nelem, 3, nelem, rounds * npix, seed, 0.0, 0.9, abs_=1.5 * k, sca=tau_s * k, dsc=DSC6, csc=CSC6,
                 roi_load=field)
        return dict(tabs=X.tabs.copy())
    return run


def run_sca(kind, npix=(24, 20), dirs=((0.0, 0.0), (70.0, 30.0)), batch=3, glob=1024, seed=0.61, pspos=None,
            hp_observer=None, nside=4, msf=False, emweight=False):
    """kind: ps / bg / hp / cl.  hp_observer = (x,y,z): one Healpix image (NSIDE `nside`) seen from that position
    instead of orthographic maps."""
    def run(X):
        c = X.cloud
        k = _tau_scale(c, 1.0)
        _, od, ra, de = observer_directions([d[0] for d in dirs], [d[1] for d in dirs])
        centre = np.array([0.5 * c.NX, 0.5 * c.NY, 0.5 * c.NZ], np.float32)
        ndir = len(dirs)
        if hp_observer is not None:
            ndir = -nside
            od = np.asarray([hp_observer], np.float32)
            ra, de = ra[:1], de[:1]
        args = (ndir, npix[0], npix[1], 0.9 * c.NX / npix[0], centre, od, ra, de)
        opac = _msf_inputs(c, k) if msf else dict(abs_=1.0 * k, sca=2.5 * k, dsc=DSC6, csc=CSC6)
        if kind == "ps":
            pp = np.asarray(pspos, np.float32)
            ps = np.linspace(1.0, 2.0, len(pp)).astype(np.float32)
            out = X.sca_ps(glob, glob * batch, batch * len(pp), seed, *args, pspos=pp.reshape(-1), ps=ps, **opac)
        elif kind == "bg":
            g = 8 * c.AREA
            out = X.sca_pb(g, 1, g * batch, batch, seed, 1.5, *args, **opac)
        elif kind == "roi":                                  # SOURCE == 3: the stored external field
            rd_, ns_ = X.opts["roi_dim"], X.opts["roi_nside"]
            nelem = rd_[0] * rd_[1] + rd_[1] * rd_[2] + rd_[2] * rd_[0]
            rng = np.random.default_rng(13)
            field = rng.random(nelem * 12 * ns_ * ns_).astype(np.float32)
            field[::5] = 0.0
            out = X.sca_pb(100 * nelem, 3, nelem, batch * 12 * ns_ * ns_, seed, 0.0, *args, roi_load=field, **opac)
        elif kind == "hp":
            rng = np.random.default_rng(11)
            sky = (0.2 + rng.random(49152)).astype(np.float32)
            sky[20000:20400] *= 30.0
            hpbgp = None
            if X.opts.get("hpbg_weighted", 0):
                p = np.clip(sky.astype(np.float64) / sky.mean(), 1e-3, 1e4)
                p /= p.sum()
                hpbgp = np.cumsum(p)
                hpbgp[-1] = 1.00001
                sky = (sky * (1.0 / 49152.0) / p).astype(np.float32)
                hpbgp = hpbgp.astype(np.float32)
            out = X.sca_hp(glob, glob * batch, batch, seed, *args, hpbg=sky, hpbgp=hpbgp, **opac)
        else:
            rng = np.random.default_rng(7)
            emit = np.where(c.DENS > 0, c.DENS * (0.5 + rng.random(c.CELLS)), 0.0).astype(np.float32)
            emwei = None
            if emweight:
                emwei = (3.0 * rng.random(c.CELLS)).astype(np.float32)
                emwei[::7] = 0.0
            out = X.sca_cl(glob, c.CELLS * batch, batch, seed, *args, emit=emit, emwei=emwei, **opac)
        return dict(out=out.copy())
    return run


# name -> (cloud factory, option dict (former -D macros), run)
CASES = {
    "bg_reg16":        (_reg(16), {}, run_bg()),
    "bg_reg16_iso":    (_reg(16), {}, run_bg(g=False, tau_s=6.0)),
    "bg_reg12_int":    (_reg(12), dict(noabsorbed=0), run_bg(batch=3, seed=0.9)),
    "bg_reg12_int2":   (_reg(12), dict(save_intensity=2), run_bg(batch=2, seed=0.23)),
    "bg_reg12_abu":    (_reg(12), dict(with_abu=1), run_abu()),
    "bg_oct8_3":       (_oct(8, 3), {}, run_bg(batch=2)),
    "bg_oct6_4":       (_oct(6, 4, 0.25, 8), dict(noabsorbed=0), run_bg(batch=3, seed=0.66)),
    "ps_reg16_in":     (_reg(16), dict(no_ps=2), run_ps([(8.3, 8.3, 8.3), (3.7, 11.2, 5.1)])),
    "ps_oct8_in":      (_oct(8, 3), dict(no_ps=1), run_ps([(4.3, 4.2, 3.9)], batch=12)),
    "ps_reg12_ext0":   (_reg(12), dict(no_ps=1, ps_method=0), run_ps([(6.0, 6.0, 20.0)], batch=40)),
    "ps_reg12_ext1":   (_reg(12), dict(no_ps=1, ps_method=1), run_ps([(-9.0, 5.0, 7.0)], batch=20)),
    "ps_reg12_ext2":   (_reg(12), dict(no_ps=1, ps_method=2),
                        run_ps([(6.0, 17.0, 16.0)], batch=10, xps_nside=[2], xps_side=[2, 4, 0],
                               xps_area=[0.5, 0.5, 0.0])),
    "ps_reg12_ext5":   (_reg(12), dict(no_ps=1, ps_method=5),
                        run_ps([(6.0, 6.0, 25.0)], batch=10, xps_nside=[1], xps_side=[4, 0, 0],
                               xps_area=[0.8, 0.0, 0.0])),
    # PS_METHOD 4 (kernel_ASOC.c:377-397): source above the cloud in z, packets sent into the cone that holds the cloud
    "ps_reg12_ext4":   (_reg(12), dict(no_ps=1, ps_method=4), run_ps([(6.0, 6.0, 30.0)], batch=20)),
    # NX > 100 with LEVELS >= 3: positions in double precision (DIMLIM, kernel_ASOC_aux.c:25-37, 207-211)
    "bg_oct101_dbl":   (lambda: synth.box_cloud(101, 6, 6, levels=3, refine_fraction=0.15), dict(noabsorbed=0), run_bg(batch=2, seed=0.45)),
    "hp_reg12":        (_reg(12), {}, run_hp(False)),
    "hp_reg12_w":      (_reg(12), dict(hpbg_weighted=1), run_hp(True)),
    "cl_reg10":        (_reg(10), {}, run_cl(False)),
    "cl_oct6_ew_ali":  (_oct(6, 3), dict(use_emweight=1, with_ali=1), run_cl(True)),
    # ALI with per-frequency absorptions: INT takes the absorptions of the emitting cell too (kernel_ASOC.c:1486-1499)
    "cl_reg10_ali_int": (_reg(10), dict(with_ali=1, noabsorbed=0), run_cl(False, seed=0.17)),
    "map_reg16":       (_reg(16), {}, run_map((20, 16), [(0.0, 0.0), (90.0, 0.0), (60.0, 30.0)])),
    "map_reg16_colden": (_reg(16), {}, run_map((16, 16), [(35.0, 110.0)], map_dx=0.7, colden=1, centre_off=0.8)),
    "map_reg120_dbl":  (lambda: Cloud(120, 8, 8, [120 * 64], synth.plummer_density(120)[56:64, 56:64, :].ravel()),
                        {}, run_map((30, 8), [(90.0, 90.0), (50.0, 20.0)], map_dx=3.9)),
    # NX >= 200: the other ray set-up branch of Mapping (kernel_ASOC_map.c:571-626), the one 256^3 / 512^3 runs take
    "map_reg200_far":  (lambda: Cloud(200, 8, 8, [200 * 64], synth.plummer_density(200)[96:104, 96:104, :].ravel()),
                        {}, run_map((40, 8), [(90.0, 90.0), (50.0, 20.0), (0.0, 0.0)], map_dx=4.9)),
    "map_oct8_3":      (_oct(8, 3), dict(with_abu=1), run_map((24, 24), [(0.0, 0.0), (60.0, 30.0)], map_dx=0.4, abu=True)),
    "map_oct6_4_thr":  (_oct(6, 4, 0.25, 8), dict(level_threshold=1), run_map((20, 20), [(120.0, 200.0)], map_dx=0.35)),
    "map_reg16_persp": (_reg(16), {}, run_map((32, 16), [(0.0, 0.0)], intobs=(7.3, 8.4, 9.1))),
    "hpmap_oct8_3":    (_oct(8, 3), {}, run_hpmap(8, (4.2, 3.3, 5.1))),
    "map_lev_oct8_3":  (_oct(8, 3), {}, run_maplev((24, 20), [(0.0, 0.0), (60.0, 30.0)], map_dx=0.4)),
    "map_lev_oct6_4_abu": (_oct(6, 4, 0.25, 8), dict(with_abu=1), run_maplev((20, 20), [(120.0, 200.0)], map_dx=0.35, abu=True, colden=True, centre_off=0.3)),
    "map_lev_reg16_persp": (_reg(16), {}, run_maplev((32, 16), [(0.0, 0.0)], intobs=(7.3, 8.4, 9.1))),
    "map_reg16_int1":  (_reg(16), dict(map_interpolation=1), run_map((20, 16), [(0.0, 0.0), (60.0, 30.0)])),
    "map_oct8_3_int1": (_oct(8, 3), dict(map_interpolation=1), run_map((24, 24), [(35.0, 110.0)], map_dx=0.4)),
    "map_reg16_int2":  (_reg(16), dict(map_interpolation=2), run_map((20, 16), [(90.0, 0.0), (60.0, 30.0)])),
    "map_oct6_4_int2": (_oct(6, 4, 0.25, 8), dict(map_interpolation=2, with_abu=1), run_map((20, 20), [(120.0, 200.0)], map_dx=0.35, abu=True)),
    "map_pstau_reg16":  (_reg(16), dict(no_ps=3), run_pstau([(8.3, 8.3, 8.3), (2.2, 13.1, 5.5), (15.6, 0.7, 9.9)], [(0.0, 0.0), (60.0, 30.0)])),
    "map_pstau_oct8":   (_oct(8, 3), dict(no_ps=2, with_abu=1), run_pstau([(4.3, 4.2, 3.9), (1.1, 6.8, 2.4)], [(35.0, 110.0)], abu=True)),
    "sca_ps_reg16":    (_reg(16), dict(no_ps=2), run_sca("ps", pspos=[(8.3, 8.3, 8.3), (4.1, 10.7, 12.2)])),
    "sca_ps_oct8":     (_oct(8, 3), dict(no_ps=1, ffs=0), run_sca("ps", pspos=[(4.3, 4.2, 3.9)], batch=8)),
    "sca_bg_reg12":    (_reg(12), {}, run_sca("bg", batch=1)),
    "sca_bg_oct6":     (_oct(6, 3), {}, run_sca("bg", batch=1, dirs=((45.0, 45.0),))),
    "sca_hp_reg12":    (_reg(12), {}, run_sca("hp", batch=4, glob=512)),
    "sca_hp_oct6_w":   (_oct(6, 3), dict(hpbg_weighted=1, ffs=0), run_sca("hp", batch=4, glob=512, dirs=((120.0, 200.0),))),
    "sca_cl_reg10":    (_reg(10), {}, run_sca("cl", batch=1, glob=256)),
    "sca_cl_oct6_ew":  (_oct(6, 3), dict(use_emweight=1), run_sca("cl", batch=1, glob=256, emweight=True, dirs=((35.0, 110.0),))),
    "sca_ps_reg12_hpobs": (_reg(12), dict(no_ps=1), run_sca("ps", pspos=[(6.3, 6.2, 5.9)], batch=6, glob=512,
                                                           hp_observer=(5.1, 7.3, 6.6))),
    "sca_bg_oct6_hpobs":  (_oct(6, 3), {}, run_sca("bg", batch=1, hp_observer=(20.0, -5.0, 3.0), nside=8)),
    "sca_hp_reg12_hpobs": (_reg(12), {}, run_sca("hp", batch=4, glob=512, hp_observer=(5.1, 7.3, 6.6))),
    "sca_cl_reg10_hpobs": (_reg(10), {}, run_sca("cl", batch=1, glob=256, hp_observer=(5.1, 3.3, 6.6))),
    "bg_reg12_msf":    (_reg(12), dict(with_abu=1, with_msf=1, ndust=2, noabsorbed=0), run_bg_msf()),
    "bg_oct6_msf":     (_oct(6, 3), dict(with_abu=1, with_msf=1, ndust=2), run_bg_msf(seed=0.91)),
    "sca_bg_reg12_msf": (_reg(12), dict(with_abu=1, with_msf=1, ndust=2), run_sca("bg", batch=1, msf=True)),
    "sca_ps_oct6_msf": (_oct(6, 3), dict(no_ps=1, with_abu=1, with_msf=1, ndust=2),
                        run_sca("ps", pspos=[(3.3, 3.2, 2.9)], batch=8, glob=512, msf=True)),
    "bg_reg12_roisave": (_reg(12), dict(with_roi_save=1, roi=[3, 8, 2, 7, 4, 9], roi_step=2, roi_nside=2),
                         with_roi_save(run_bg(batch=2, seed=0.35))),
    "ps_oct6_roisave": (_oct(6, 3), dict(no_ps=1, with_roi_save=1, roi=[3, 4, 2, 4, 1, 3], roi_step=1, roi_nside=4),
                        with_roi_save(run_ps([(1.3, 1.2, 4.9)], batch=12))),
    "cl_reg10_roisave": (_reg(10), dict(with_roi_save=1, roi=[2, 6, 3, 7, 2, 5], roi_step=1, roi_nside=2),
                         with_roi_save(run_cl(False))),
    "roi_reg12_load":  (_reg(12), dict(with_roi_load=1, roi_dim=[4, 4, 4], roi_nside=2), run_roi_load([4, 4, 4], 2)),
    "roi_oct6_load":   (_oct(6, 3), dict(with_roi_load=1, roi_dim=[3, 3, 3], roi_nside=2, noabsorbed=0), run_roi_load([3, 3, 3], 2, rounds=1, tau_s=0.5)),   # few scatterings: few paths flip on libm rounding
    "sca_roi_reg12_load": (_reg(12), dict(with_roi_load=1, roi_dim=[4, 4, 4], roi_nside=2), run_sca("roi", batch=1)),
    "map_reg16_roi":   (_reg(16), dict(roi_map=1, roi=[4, 11, 3, 9, 5, 12]), run_map((20, 16), [(0.0, 0.0), (60.0, 30.0)])),
    "map_oct8_3_roi":  (_oct(8, 3), dict(roi_map=1, roi=[2, 5, 1, 6, 3, 4]), run_map((24, 24), [(35.0, 110.0)], map_dx=0.4)),
    "bg_reg12_mirror": (_reg(12), dict(mirror=1 + 4), run_bg(batch=2, seed=0.57)),
    "bg_oct6_mirror":  (_oct(6, 3), dict(mirror=32), run_bg(batch=2, seed=0.58)),
    "cl_reg10_mirror": (_reg(10), dict(mirror=16 + 2), run_cl(False)),
    "sca_bg_reg12_mirror": (_reg(12), dict(mirror=1 + 8), run_sca("bg", batch=1)),
}

MAP_NSIDE = {"hpmap_oct8_3": 8}
