"""TEST INFRASTRUCTURE: materialises the reference's runnable example (soc_example.zip, BASELINE.json configs[0]) from
the fixtures in tests/golden/soc_example/ (generator: tests/golden/make_soc_example.py)."""
import lzma
import os
import shutil

import numpy as np

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "soc_example")


def decode_dsc():
    nfreq, bins = np.fromfile(os.path.join(FIX, "tmp.dsc.shape"), np.int32)
    planes = np.frombuffer(lzma.decompress(open(os.path.join(FIX, "tmp.dsc.delta.xz"), "rb").read()), np.uint8)
    d = np.ascontiguousarray(planes.reshape(4, -1).T).view(np.int32).reshape(2 * nfreq, bins)
    raw = (np.cumsum(d.astype(np.int64), axis=1) & 0xffffffff).astype(np.uint32).view(np.int32)     # sums wrap like int32
    return raw.view(np.float32).reshape(2, nfreq, bins)


def write_example(path, n=64, edits=None, drop=()):
    """Writes the example model into `path`.  The cloud follows the archive's make_cloud.py: an n^3 cube of unit
    density (n = 64 there).  `edits` = {keyword: new line or None (remove)} applied to my.ini; `drop` keywords are removed."""
    os.makedirs(path, exist_ok=True)
    for name in ("freq.dat", "tmp.dust", "bg_intensity.bin"):
        shutil.copy(os.path.join(FIX, name), os.path.join(path, name))
    decode_dsc().tofile(os.path.join(path, "tmp.dsc"))
    with open(os.path.join(path, "tmp.cloud"), "wb") as fp:
        np.asarray([n, n, n, 1, n * n * n], np.int32).tofile(fp)
        np.asarray([n * n * n], np.int32).tofile(fp)
        np.ones((n, n, n), np.float32).tofile(fp)
    edits = dict(edits or {})
    lines = []
    for line in open(os.path.join(FIX, "my.ini")):
        key = line.split()[0] if line.split() else ""
        if key in drop:
            continue
        if key in edits:
            new = edits.pop(key)
            if new is not None:
                lines.append(new.rstrip("\n") + "\n")
            continue
        lines.append(line)
    lines += [v.rstrip("\n") + "\n" for v in edits.values() if v is not None]
    ini = os.path.join(path, "my.ini")
    open(ini, "w").writelines(lines)
    return ini
