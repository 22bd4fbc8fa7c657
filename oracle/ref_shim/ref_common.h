// TEST INFRASTRUCTURE ONLY.  Shared bits of the three harness translation units
// that execute the reference's own OpenCL kernels on the host (see cl_shim.h).
#pragma once
#include "cl_shim.h"
#include <omp.h>

// Count of successful float atomic adds (the reference's CAS loop goes through
// atomic_cmpxchg, which the shim owns): one per absorption update, i.e. the
// "cell-step" unit of SURVEY.md section 8(d).
extern thread_local unsigned long clshim_atomic_ok;

// Sampling of the work items (bench.py's bounded CPU samples): loop index i runs work item ref_offset + i * ref_stride
// of a launch of ref_gsize work items (0 = the loop count).  Defaults 1 / 0 / 0 = the plain launch.
extern long ref_stride, ref_offset, ref_gsize;

#define REF_PARALLEL_FOR(GLOBAL, BODY)                                          \
    do {                                                                        \
        const long _g = (long)(GLOBAL);                                         \
        const long _gs = ref_gsize > 0 ? ref_gsize : _g;                        \
        _Pragma("omp parallel for schedule(runtime)")                       \
        for (long _i = 0; _i < _g; ++_i) {                                      \
            clshim_gid = (size_t)(ref_offset + _i * ref_stride); clshim_gsize = (size_t)_gs; \
            BODY;                                                               \
        }                                                                       \
    } while (0)
