"""Guard against a miscompilation seen with nvcc 12.9: inside the lean / look-ahead packet kernels the expression
`px ? 7 : (py ? brick_by : brick_bz)` was folded to `px ? 7 : brick_bz` (the y stride was never loaded) once the kernel
also defined pz = !px && !py -- packets stepping along y then jumped to the wrong brick.  The kernels are written in a
form that survives; this test compiles sim.cu to PTX and checks that every bricked kernel still loads BOTH strides and
every linear-layout kernel the x-y slab size."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")
def test_brick_strides_survive_in_ptx(tmp_path):
    csrc = os.path.join(ROOT, "soc_b200", "csrc")
    inc = os.path.join(ROOT, "include")
    off = tmp_path / "off.cu"
    off.write_text('#include <cstdio>\n#include <cstddef>\n#include "sim.cuh"\n'
                   'int main(){ printf("%zu %zu %zu\\n", offsetof(SimArgs,slab_xy), offsetof(SimArgs,brick_by), offsetof(SimArgs,brick_bz)); }\n')
    exe = str(tmp_path / "off")
    subprocess.check_call([NVCC, "-std=c++17", "-I", csrc, "-I", inc, "-o", exe, str(off)], stderr=subprocess.DEVNULL)
    slab, by, bz = (int(x) for x in subprocess.check_output([exe]).split())
    ptx = str(tmp_path / "sim.ptx")
    subprocess.check_call([NVCC, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I", inc, "-I", csrc,
                           "--ftz=false", "--prec-div=true", "--prec-sqrt=true", "-DSOC_BUILDING", "-ptx",
                           os.path.join(csrc, "sim.cu"), "-o", ptx])
    entries = re.split(r"\n\.entry ", open(ptx).read())[1:]
    seen = 0
    for e in entries:
        name = e.split("(")[0]
        m = re.search(r"sim_(lean|ahead)_kernelILi\dELb([01])", name)
        if not m:
            continue
        seen += 1
        n = {k: len(re.findall(r"ld\.param\.u32\s+%%r\d+, \[%%rd\d+\+%d\]" % o, e)) for k, o in (("slab", slab), ("by", by), ("bz", bz))}
        if m.group(2) == "1":
            assert n["by"] >= 1 and n["bz"] >= 1, (name, n)
        else:
            assert n["slab"] >= 1, (name, n)
    assert seen >= 12
