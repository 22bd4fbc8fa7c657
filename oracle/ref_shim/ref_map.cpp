// TEST INFRASTRUCTURE ONLY.  Runs the reference's kernel_ASOC_map.c (Mapping,
// HealpixMapping) on the host through cl_shim.h.  Separate translation unit:
// the map kernels carry their own IndexG/Index/GetStep with different EPS/PEPS.
#include "ref_common.h"
#include <vector>

namespace refm {
#include "kernel_ASOC_map.c"
}

extern "C" {

void ref_mapping(int global, float MAP_DX, int npx, int npy, float *MAP, float *EMIT, const float *DIR,
                 const float *RA, const float *DE, const int *LCELLS, const int *OFF, int *PAR, float *DENS,
                 float ABS, float SCA, const float *CENTRE, const float *INTOBS, float *OPT, float *SAVETAU,
                 int SAVE_COLDEN, const int *ROI) {
    int2 NPIX(npx, npy);
    float3 d(DIR[0], DIR[1], DIR[2]), ra(RA[0], RA[1], RA[2]), de(DE[0], DE[1], DE[2]);
    float3 c(CENTRE[0], CENTRE[1], CENTRE[2]), io(INTOBS[0], INTOBS[1], INTOBS[2]);
    REF_PARALLEL_FOR(global,
        refm::Mapping(MAP_DX, NPIX, MAP, EMIT, d, ra, de, LCELLS, OFF, PAR, DENS, ABS, SCA, c, io, OPT, SAVETAU,
                      SAVE_COLDEN
#if (ROI_MAP>0)
                      , ROI
#endif
                      ));
}

void ref_healpix_mapping(int global, float MAP_DX, int npx, int npy, float *MAP, float *EMIT, const float *DIR,
                         const float *RA, const float *DE, const int *LCELLS, const int *OFF, int *PAR, float *DENS,
                         float ABS, float SCA, const float *CENTRE, const float *INTOBS, float *OPT,
                         float *SAVETAU, int SAVE_COLDEN, const int *ROI) {
    int2 NPIX(npx, npy);
    float3 d(DIR[0], DIR[1], DIR[2]), ra(RA[0], RA[1], RA[2]), de(DE[0], DE[1], DE[2]);
    float3 c(CENTRE[0], CENTRE[1], CENTRE[2]), io(INTOBS[0], INTOBS[1], INTOBS[2]);
    REF_PARALLEL_FOR(global,
        refm::HealpixMapping(MAP_DX, NPIX, MAP, EMIT, d, ra, de, LCELLS, OFF, PAR, DENS, ABS, SCA, c, io, OPT,
                             SAVETAU, SAVE_COLDEN
#if (ROI_MAP>0)
                             , ROI
#endif
                             ));
}

void ref_pstau(int global, int no, const float *PSPOS_xyz, const float *DIR, const float *RA, const float *DE,
               const int *LCELLS, const int *OFF, int *PAR, float *DENS, float ABS, float SCA, float *OPT,
               float *pscolden, float *pstau) {
    float3 d(DIR[0], DIR[1], DIR[2]), ra(RA[0], RA[1], RA[2]), de(DE[0], DE[1], DE[2]);
    std::vector<float3> pos(no);
    for (int i = 0; i < no; i++) pos[i] = float3(PSPOS_xyz[3 * i], PSPOS_xyz[3 * i + 1], PSPOS_xyz[3 * i + 2]);
    REF_PARALLEL_FOR(global, refm::PSTau(no, pos.data(), d, ra, de, LCELLS, OFF, PAR, DENS, ABS, SCA, OPT, pscolden, pstau));
}

}  // extern "C"
