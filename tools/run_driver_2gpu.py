"""Development tool: the ASOC driver on 1 GPU and on 2 GPUs (torchrun, NCCL) for the same model; the absorbed file and
the temperatures must agree to the order of float additions (same Philox streams).  The constant sources are sharded by
frequency by default (rank r runs frequencies r, r+2, ... whole: no per-frequency collective), by packet index with the
ini key PACKETSHARD; both are run."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soc_b200.formats import read_cells_freq_file, read_otfile, read_map_file, read_outcoming  # noqa: E402
from tests.model import write_model  # noqa: E402

base = os.environ.get("SOC_2GPU_DIR", "/tmp/soc2gpu")


def pair(name, script="ASOC.py", **kw):
    for tag in ("one", "two"):
        write_model(os.path.join(base, name, tag), **kw)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "bin", script), "model.ini"], cwd=os.path.join(base, name, "one"))
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", "29533", os.path.join(ROOT, "bin", script), "model.ini"], cwd=os.path.join(base, name, "two"))
    return os.path.join(base, name, "one"), os.path.join(base, name, "two")


# absorbed file: the [CELLS, NFREQ] array stays on each rank's device and is all-reduced once at the end
for name, extra in (("abs_freqshard", ""), ("abs_packetshard", "PACKETSHARD\n")):
    d1, d2 = pair(name, n=16, bgpac=400000, pspac=330000, noabsorbed=False, absorbed=True, maps=False, extra=extra)
    a1 = read_cells_freq_file(os.path.join(d1, "abs.data")).astype(np.float64)
    a2 = read_cells_freq_file(os.path.join(d2, "abs.data")).astype(np.float64)
    print("absorbed (%s): max rel diff %.3e (sum %.6e vs %.6e)" % (name, np.abs(a1 - a2).max() / np.abs(a1).max(), a1.sum(), a2.sum()))
    assert np.abs(a1 - a2).max() <= 1e-4 * np.abs(a1).max()
# integrated absorptions -> temperatures -> maps, octree
d1, d2 = pair("temp", n=8, octree=True, bgpac=200000, pspac=330000, extra="CLT\nCLE\n")
t1, t2 = read_otfile(os.path.join(d1, "model.T")), read_otfile(os.path.join(d2, "model.T"))
m1, m2 = read_map_file(os.path.join(d1, "map_dir_00.bin")), read_map_file(os.path.join(d2, "map_dir_00.bin"))
print("temperature: max abs diff %.3e K" % np.abs(t1 - t2).max())
print("map: max rel diff %.3e" % (np.abs(m1 - m2).max() / np.abs(m1).max()))
assert np.abs(t1 - t2).max() < 1e-2 and np.abs(m1 - m2).max() <= 1e-3 * np.abs(m1).max()
# scattered light (ASOCS.py): the images are all-reduced
d1, d2 = pair("sca", script="ASOCS.py", n=8, octree=True, bgpac=200000, pspac=330000)
o1, o2 = read_outcoming(os.path.join(d1, "outcoming.socs")), read_outcoming(os.path.join(d2, "outcoming.socs"))
i1, i2 = np.asarray(o1[-1] if isinstance(o1, tuple) else o1, np.float64), np.asarray(o2[-1] if isinstance(o2, tuple) else o2, np.float64)
print("scattered light: max rel diff %.3e (sum %.6e vs %.6e)" % (np.abs(i1 - i2).max() / np.abs(i1).max(), i1.sum(), i2.sum()))
assert np.abs(i1 - i2).max() <= 1e-3 * np.abs(i1).max()
print("2-GPU driver runs agree with the 1-GPU runs")
