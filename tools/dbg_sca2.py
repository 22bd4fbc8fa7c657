import os, sys, tempfile, numpy as np
sys.path.insert(0, '/root/repo')
from tests.model import write_model
from tests.oracle_device import OracleDevice
from soc_b200 import asocs
from soc_b200.formats import read_outcoming
base = tempfile.mkdtemp()
for label, kw in (("ps", dict(bgpac=0, pspac=655360)), ("bg", dict(bgpac=200000, pspac=0))):
    for name, fac, extra in (("gpu", None, ""), ("gpuref", None, "REFSTREAMS\n"), ("cpu", OracleDevice, "")):
        d = os.path.join(base, label + name)
        write_model(d, n=10, extra=extra, **kw)
        os.chdir(d)
        asocs.main(["ASOCS.py", "model.ini"], device_factory=fac)
        out = read_outcoming(os.path.join(d, "outcoming.socs"))[1].astype(np.float64)
        print(label, name, " ".join("%.4e" % out[f].sum() for f in range(8)))
