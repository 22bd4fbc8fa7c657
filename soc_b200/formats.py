"""On-disk formats of SOC: cloud / temperature hierarchies, dust,
scattering functions, source spectra, absorbed / emitted files and maps.

Behaviour follows the reference readers (ASOC_aux.py:557-803, 1092-1123, 1420-1442) but is written
against numpy only (no pyopencl / matplotlib imports).
"""
import numpy as np

from .constants import PARSEC


# ------------------------------------------------------------------------------------------------
# link encoding of the octree hierarchy (ASOC_aux.py:14-20, kernel_ASOC_aux.c:156-158)
# ------------------------------------------------------------------------------------------------
def links_to_float(first_child_index):
    """int32 index within the next level -> float32 link value (-bitcast(index)); index 0 -> -0.0."""
    idx = np.asarray(first_child_index, np.int32).view(np.uint32)
    return (idx | np.uint32(0x80000000)).view(np.float32)


def float_to_links(values):
    """float32 link values -> int32 index within the next level."""
    return (np.asarray(values, np.float32).view(np.uint32) & np.uint32(0x7FFFFFFF)).view(np.int32)


def is_link(values):
    """True where a hierarchy value is a link (value <= 0, including -0.0), cf. kernel_ASOC_aux.c:149."""
    return ~(np.asarray(values, np.float32) > 0.0)


class Cloud:
    """Density hierarchy: NX,NY,NZ root grid, LEVELS levels, LCELLS/OFF per level, DENS[CELLS]."""

    def __init__(self, nx, ny, nz, lcells, dens):
        self.NX, self.NY, self.NZ = int(nx), int(ny), int(nz)
        self.LCELLS = np.asarray(lcells, np.int32)
        self.LEVELS = len(self.LCELLS)
        self.OFF = np.zeros(self.LEVELS, np.int32)
        self.OFF[1:] = np.cumsum(self.LCELLS)[:-1]
        self.DENS = np.ascontiguousarray(dens, np.float32)
        self.CELLS = int(self.LCELLS.sum())
        assert self.DENS.size == self.CELLS
        self.AREA = 2 * (self.NX * self.NY + self.NY * self.NZ + self.NZ * self.NX)

    def level_slice(self, level):
        a = int(self.OFF[level])
        return slice(a, a + int(self.LCELLS[level]))

    def level_of_cells(self):
        lev = np.zeros(self.CELLS, np.int8)
        for l in range(self.LEVELS):
            lev[self.level_slice(l)] = l
        return lev

    def leaf_mask(self):
        return self.DENS > 0.0


def write_cloud(filename, cloud, values=None):
    """Write a hierarchy file (cloud, temperature, ...): ASOC_aux.py:734-744, ASOC.py:2142-2149."""
    vals = cloud.DENS if values is None else np.asarray(values, np.float32)
    with open(filename, "wb") as fp:
        np.asarray([cloud.NX, cloud.NY, cloud.NZ, cloud.LEVELS, cloud.CELLS], np.int32).tofile(fp)
        for l in range(cloud.LEVELS):
            np.asarray([cloud.LCELLS[l]], np.int32).tofile(fp)
            vals[cloud.level_slice(l)].tofile(fp)


def read_cloud(filename, kdensity=1.0):
    """Read a cloud file; leaf densities are scaled by `kdensity` and clipped to [1e-6, 1e20]
    (ASOC_aux.py:745-803); links are left untouched."""
    with open(filename, "rb") as fp:
        nx, ny, nz, levels, cells = np.fromfile(fp, np.int32, 5)
        lcells, parts = [], []
        for _ in range(levels):
            n = int(np.fromfile(fp, np.int32, 1)[0])
            if n < 0:
                break
            tmp = np.fromfile(fp, np.float32, n)
            if kdensity != 1.0:
                m = tmp > 0.0
                tmp[m] = np.clip(kdensity * tmp[m], 1.0e-6, 1e20)
            lcells.append(n)
            parts.append(tmp)
    return Cloud(nx, ny, nz, lcells, np.concatenate(parts))


def cut_levels(infile, outfile, maxlevel):
    """Write a copy of a hierarchy file without the levels > maxlevel (0, 1, ...): parent cells of the removed
    levels become leaves holding the plain average of their eight children (OT_cut_levels, ASOC_aux.py:651-713 with
    kernel AverageParent, kernel_OT_tools.c:5-22; float32 sum in child order, then / 8)."""
    with open(infile, "rb") as fp:
        nx, ny, nz, levels, cells = np.fromfile(fp, np.int32, 5)
        H = []
        for _ in range(levels):
            n = int(np.fromfile(fp, np.int32, 1)[0])
            H.append(np.fromfile(fp, np.float32, n))
    maxlevel = min(levels - 1, maxlevel)
    for i in range(levels - 2, maxlevel - 1, -1):
        par = np.nonzero(H[i] <= 1.0e-9)[0]
        first = (-H[i][par]).view(np.int32)
        acc = np.zeros(len(par), np.float32)
        for k in range(8):
            acc = acc + H[i + 1][first + k]
        H[i][par] = acc / np.float32(8.0)
    with open(outfile, "wb") as fp:
        np.asarray([nx, ny, nz, maxlevel + 1, sum(len(h) for h in H[:maxlevel + 1])], np.int32).tofile(fp)
        for h in H[:maxlevel + 1]:
            np.asarray([len(h)], np.int32).tofile(fp)
            np.asarray(h, np.float32).tofile(fp)


def read_otfile(filename):
    """Hierarchy file -> flat value vector (ASOC_aux.py:1420-1442)."""
    with open(filename, "rb") as fp:
        nx, ny, nz, levels, cells = np.fromfile(fp, np.int32, 5)
        val = np.zeros(cells, np.float32)
        a = 0
        for _ in range(levels):
            n = int(np.fromfile(fp, np.int32, 1)[0])
            val[a:a + n] = np.fromfile(fp, np.float32, n)
            a += n
    return val


# ------------------------------------------------------------------------------------------------
# dust and scattering functions
# ------------------------------------------------------------------------------------------------
def read_dust(filename, gl):
    """'eqdust' text file -> FREQ, G, ABS, SCA with ABS/SCA = optical depth per unit density per root
    cell length GL [pc] (ASOC_aux.py:579-587)."""
    lines = open(filename).readlines()
    grain_density = float(lines[1].split()[0])
    grain_size = float(lines[2].split()[0])
    coeff = grain_density * np.pi * grain_size ** 2.0 * gl * PARSEC
    d = np.loadtxt(filename, skiprows=4, ndmin=2)
    return (np.asarray(d[:, 0], np.float32), np.asarray(d[:, 1], np.float32),
            np.asarray(d[:, 2] * coeff, np.float32), np.asarray(d[:, 3] * coeff, np.float32))


def write_dust(filename, freq, g, qabs, qsca, grain_density=1.0e-7, grain_size=1.0e-4):
    """Writer matching DustLib.write_simple_dust's text layout (DustLib.py:1701-1708)."""
    with open(filename, "w") as fp:
        fp.write("eqdust\n%12.5e\n%12.5e\n%d\n" % (grain_density, grain_size, len(freq)))
        for i in range(len(freq)):
            fp.write("%12.5e  %8.5f  %12.5e %12.5e\n" % (freq[i], g[i], qabs[i], qsca[i]))


def read_dsc(filename, nfreq, bins):
    """.dsc file -> DSC[nfreq,bins] (phase function per steradian on cos(theta)=linspace(-1,1)) and
    CSC[nfreq,bins] (cos(theta) at cumulative probability linspace(0,1)) (ASOC_aux.py:643-646)."""
    with open(filename, "rb") as fp:
        dsc = np.fromfile(fp, np.float32, nfreq * bins).reshape(nfreq, bins)
        csc = np.fromfile(fp, np.float32).reshape(nfreq, bins)
    return dsc, csc


def write_dsc(filename, dsc, csc):
    with open(filename, "wb") as fp:
        np.asarray(dsc, np.float32).tofile(fp)
        np.asarray(csc, np.float32).tofile(fp)


# ------------------------------------------------------------------------------------------------
# absorbed / emitted / maps
# ------------------------------------------------------------------------------------------------
def write_cells_freq_file(filename, data):
    """absorbed/emitted layout: int32 CELLS,NFREQ + float32 [CELLS,NFREQ] (ASOC.py:2866-2875, 3971-3975)."""
    data = np.asarray(data, np.float32)
    with open(filename, "wb") as fp:
        np.asarray(data.shape, np.int32).tofile(fp)
        data.tofile(fp)


def read_cells_freq_file(filename):
    with open(filename, "rb") as fp:
        cells, nfreq = np.fromfile(fp, np.int32, 2)
        return np.fromfile(fp, np.float32).reshape(cells, nfreq)


def read_map_file(filename):
    """map_dir_XX.bin: int32 NPIX.x,NPIX.y + float32 [nfreq, NPIX.y, NPIX.x] (ASOC.py:2994-2998,3151)."""
    with open(filename, "rb") as fp:
        nx, ny = np.fromfile(fp, np.int32, 2)
        return np.fromfile(fp, np.float32).reshape(-1, ny, nx)


def read_outcoming(filename):
    """outcoming.socs of ASOCS.py (orthographic form, ASOCS.py:391-397)."""
    with open(filename, "rb") as fp:
        ny, nx, nfreq = np.fromfile(fp, np.int32, 3)
        freq = np.fromfile(fp, np.float32, nfreq)
        data = np.fromfile(fp, np.float32)
    ndir = data.size // (nfreq * ny * nx)
    return freq, data.reshape(nfreq, ndir, ny, nx)
