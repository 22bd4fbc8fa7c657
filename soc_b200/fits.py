"""Minimal FITS writer for the image products of ASOC.py / ASOCS.py.

The reference builds its FITS files with astropy (`MakeFits`, ASOC_aux.py:1723-1768): a primary HDU holding a
float32 image [ny, nx] or cube [nfreq, ny, nx] with a TAN projection centred on (lon, lat).  The same cards
are written here directly (80-character cards, 2880-byte blocks, big-endian data) so that no extra package
is needed; the files open in astropy / ds9 like the reference's.
"""
import numpy as np


def _card(key, value=None, comment=None):
    if key == "COMMENT":
        s = "COMMENT " + str(value)
    elif key == "END":
        s = "END"
    else:
        if isinstance(value, bool):
            v = "%20s" % ("T" if value else "F")
        elif isinstance(value, (int, np.integer)):
            v = "%20d" % int(value)
        elif isinstance(value, (float, np.floating)):
            v = "%20s" % ("%.13G" % float(value) if "." in ("%.13G" % float(value)) or "E" in ("%.13G" % float(value))
                          else "%.13G." % float(value))
        else:
            v = "'%-8s'" % str(value)
        s = "%-8s= %s" % (key, v)
        if comment:
            s += " / " + comment
    return ("%-80s" % s)[:80]


def make_header(lon, lat, pix, nx, ny, freq=(), galactic=False):
    """Cards of MakeFits(lon, lat, pix, m=nx, n=ny, freq): pix is the pixel size in radians."""
    nchn = max(1, len(freq))
    cube = nchn > 1
    cards = [_card("SIMPLE", True, "conforms to FITS standard"), _card("BITPIX", -32, "array data type"),
             _card("NAXIS", 3 if cube else 2, "number of array dimensions"), _card("NAXIS1", int(nx)), _card("NAXIS2", int(ny))]
    if cube:
        cards.append(_card("NAXIS3", nchn))
    cards.append(_card("EXTEND", True))
    cards += [_card("CRVAL1", float(lon)), _card("CRVAL2", float(lat)),
              _card("CDELT1", -pix * 180.0 / np.pi), _card("CDELT2", pix * 180.0 / np.pi),
              _card("CRPIX1", 0.5 * (nx + 1) + 0.5), _card("CRPIX2", 0.5 * (ny + 1) + 0.5)]
    if galactic:
        cards += [_card("CTYPE1", "GLON-TAN"), _card("CTYPE2", "GLAT-TAN"), _card("COORDSYS", "GALACTIC")]
    else:
        cards += [_card("CTYPE1", "RA---TAN"), _card("CTYPE2", "DEC--TAN"), _card("COORDSYS", "EQUATORIAL"),
                  _card("EQUINOX", 2000.0)]
    if cube:
        cards += [_card("CRPIX3", 1), _card("CRVAL3", 0.0), _card("CDELT3", 1), _card("CTYPE3", "channel")]
        for i in range(nchn):
            cards.append(_card("COMMENT", "F[ %3d ] = %.4e" % (i, freq[i])))
    cards.append(_card("END"))
    return cards


def write_fits(filename, data, lon, lat, pix, freq=(), galactic=False):
    """data: [ny, nx] or [nfreq, ny, nx] (float32 on disk, as in the reference)."""
    a = np.asarray(data, np.float32)
    ny, nx = a.shape[-2], a.shape[-1]
    if a.ndim == 3 and a.shape[0] != max(1, len(freq)):
        raise ValueError("cube has %d planes for %d frequencies" % (a.shape[0], len(freq)))
    hdr = "".join(make_header(lon, lat, pix, nx, ny, freq if a.ndim == 3 else (), galactic)).encode("ascii")
    hdr += b" " * (-len(hdr) % 2880)
    raw = a.astype(">f4").tobytes()
    raw += b"\0" * (-len(raw) % 2880)
    with open(filename, "wb") as fp:
        fp.write(hdr)
        fp.write(raw)


def read_fits(filename):
    """Reads back a primary-HDU image written by write_fits (tests); returns (header dict, array)."""
    with open(filename, "rb") as fp:
        blob = fp.read()
    hdr, pos, done = {}, 0, False
    while not done:
        block = blob[pos:pos + 2880].decode("ascii")
        pos += 2880
        for i in range(0, 2880, 80):
            card = block[i:i + 80]
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] == "= ":
                val = card[10:].split(" / ")[0].strip()
                if val.startswith("'"):
                    hdr[key] = val.strip("'").strip()
                elif val in ("T", "F"):
                    hdr[key] = val == "T"
                else:
                    hdr[key] = float(val) if any(c in val for c in ".E") else int(val)
    shape = [hdr["NAXIS%d" % (i + 1)] for i in range(hdr["NAXIS"])][::-1]
    n = int(np.prod(shape))
    return hdr, np.frombuffer(blob[pos:pos + 4 * n], ">f4").reshape(shape).astype(np.float32)
