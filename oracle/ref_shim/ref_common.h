// TEST INFRASTRUCTURE ONLY.  Shared bits of the three harness translation units
// that execute the reference's own OpenCL kernels on the host (see cl_shim.h).
#pragma once
#include "cl_shim.h"
#include <omp.h>

// Count of successful float atomic adds (the reference's CAS loop goes through
// atomic_cmpxchg, which the shim owns): one per absorption update, i.e. the
// "cell-step" unit of SURVEY.md section 8(d).
extern thread_local unsigned long clshim_atomic_ok;

#define REF_PARALLEL_FOR(GLOBAL, BODY)                                          \
    do {                                                                        \
        const long _g = (long)(GLOBAL);                                         \
        _Pragma("omp parallel for schedule(runtime)")                       \
        for (long _id = 0; _id < _g; ++_id) {                                   \
            clshim_gid = (size_t)_id; clshim_gsize = (size_t)_g;                \
            BODY;                                                               \
        }                                                                       \
    } while (0)
