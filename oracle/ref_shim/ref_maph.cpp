// TEST INFRASTRUCTURE ONLY.  Runs the reference's kernel_ASOC_map_H.c (Mapping: one image per hierarchy level) on the
// host through cl_shim.h.  Built as a library of its own (oracle/build_ref.py: build_levels) from the part of the file
// up to and including Mapping.
#include "ref_common.h"

thread_local size_t clshim_gid = 0, clshim_gsize = 1;
long ref_stride = 1, ref_offset = 0, ref_gsize = 0;
thread_local unsigned long clshim_atomic_ok = 0;

namespace refh {
#include "kernel_ASOC_map_H.c"
}

extern "C" {

void ref_mapping_levels(int global, float DX, int npx, int npy, float *MAP, float *EMIT, const float *DIR,
                        const float *RA, const float *DE, const int *LCELLS, const int *OFF, int *PAR, float *DENS,
                        float ABS, float SCA, const float *CENTRE, const float *INTOBS, float *OPT, float *COLDEN) {
    int2 NPIX(npx, npy);
    float3 d(DIR[0], DIR[1], DIR[2]), ra(RA[0], RA[1], RA[2]), de(DE[0], DE[1], DE[2]);
    float3 c(CENTRE[0], CENTRE[1], CENTRE[2]), io(INTOBS[0], INTOBS[1], INTOBS[2]);
    REF_PARALLEL_FOR(global,
        refh::Mapping(DX, NPIX, MAP, EMIT, d, ra, de, LCELLS, OFF, PAR, DENS, ABS, SCA, c, io, OPT, COLDEN));
}

}  // extern "C"
