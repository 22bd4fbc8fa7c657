// TEST INFRASTRUCTURE ONLY.  Runs the reference's kernel_A2E_MABU_aux.c (split_absorbed: the hand-off of the absorbed file
// to the dust solver of one species, A2E_MABU.py:700-705) on the host through cl_shim.h.  NFREQ and NDUST are macros of
// that file, so one library per (NFREQ, NDUST) (oracle/build_ref.py: build_a2e).
#include "ref_common.h"

thread_local size_t clshim_gid = 0, clshim_gsize = 1;
long ref_stride = 1, ref_offset = 0, ref_gsize = 0;
thread_local unsigned long clshim_atomic_ok = 0;

namespace refa {
#include "kernel_A2E_MABU_aux.c"
}

extern "C" {

void ref_split_absorbed(int global, int IDUST, int N, double *RABS, float *ABU, float *IN, float *OUT) {
    REF_PARALLEL_FOR(global, refa::split_absorbed(IDUST, N, RABS, ABU, IN, OUT));
}

}  // extern "C"
