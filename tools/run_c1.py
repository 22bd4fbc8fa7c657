"""Development tool: the ASOC driver end to end on a config-1-like model (64^3, 44 frequencies, 1e6 background packets per
frequency, equilibrium temperatures, two maps) with a cProfile summary of the host side."""
import os, sys, time
sys.path.insert(0, os.getcwd())
from tests.model import write_model
from soc_b200 import asoc
import cProfile, pstats
path = os.environ.get("SOC_C1_DIR", "/tmp/c1run")
ini, cloud = write_model(path, n=64, nfreq=44, bgpac=1000000, extra="CLT\nCLE\n")
os.chdir(path)
t0 = time.time()
pr = cProfile.Profile(); pr.enable()
asoc.main(["ASOC.py", "model.ini"])
pr.disable()
print("wall %.2f s" % (time.time() - t0))
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
