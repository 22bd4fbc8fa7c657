"""TEST INFRASTRUCTURE: the reference-kernel libraries (oracle/_ref) the `-m gpu` tests and bench.py load on the GPU
box, where /root/reference does not exist.  __graft_entry__.build() calls prebuild() in the container that has the
reference sources; the .so files travel with the working tree."""
from oracle import build_ref


def configs():
    from tests.test_config_parity import c3_cloud
    c3 = c3_cloud()
    base = dict(BINS=2500, GL=0.01)
    return [
        dict(NX=256, NY=256, NZ=256, LEVELS=1, CELLS=256 ** 3, NO_PS=1, NOABSORBED=0, **base),          # C2 (and bench.py)
        dict(NX=512, NY=512, NZ=512, LEVELS=1, CELLS=512 ** 3, NO_PS=1, NOABSORBED=0, **base),          # C4/C5 (bench.py extra workloads)
        dict(NX=256, NY=256, NZ=256, LEVELS=1, CELLS=256 ** 3, NO_PS=1, NOABSORBED=0, WITH_ABU=1, **base),
        dict(NX=c3.NX, NY=c3.NY, NZ=c3.NZ, LEVELS=c3.LEVELS, CELLS=c3.CELLS, NO_PS=1, NOABSORBED=0, **base),   # C3 absorptions
        dict(NX=c3.NX, NY=c3.NY, NZ=c3.NZ, LEVELS=c3.LEVELS, CELLS=c3.CELLS, NO_PS=1, FFS=1, **base),          # C3 scattered light
    ]


def prebuild(verbose=False):
    if not build_ref.reference_available():
        return []
    out = []
    for cfg in configs():
        out.append(build_ref.build(cfg))
        if verbose:
            print(out[-1])
    return out
