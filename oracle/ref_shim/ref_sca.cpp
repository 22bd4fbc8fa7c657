// TEST INFRASTRUCTURE ONLY.  Runs the reference's kernel_ASOC_sca.c (scattered
// light, peel-off) on the host through cl_shim.h.
#include "ref_common.h"

namespace refs {
#include "kernel_ASOC_sca.c"
}

extern "C" {

void ref_sca_zero_out(int global, int NDIR, int npx, int npy, float *OUT) {
    int2 NPIX(npx, npy);
    REF_PARALLEL_FOR(global, refs::zero_out(NDIR, NPIX, OUT));
}

void ref_sca_ps(int global, int PACKETS, int BATCH, float SEED, float *ABS, float *SCA, float BG, float *PSPOS_xyz,
                float *PS, const int *LCELLS, const int *OFF, int *PAR, float *DENS, const float *DSC,
                const float *CSC, int NDIR, float *ODIRS_xyz, int npx, int npy, float MAP_DX, const float *CENTRE,
                float *ORA_xyz, float *ODE_xyz, float *OUT, float *ABU, float *OPT, float *XPS_NSIDE,
                float *XPS_SIDE, float *XPS_AREA) {
    int2 NPIX(npx, npy);
    float3 c(CENTRE[0], CENTRE[1], CENTRE[2]);
    REF_PARALLEL_FOR(global,
        refs::SimRAM_PS(PACKETS, BATCH, SEED, ABS, SCA, BG, (float3 *)PSPOS_xyz, PS, LCELLS, OFF, PAR, DENS, DSC, CSC,
                        NDIR, (float3 *)ODIRS_xyz, NPIX, MAP_DX, c, (float3 *)ORA_xyz, (float3 *)ODE_xyz, OUT, ABU,
                        OPT, XPS_NSIDE, XPS_SIDE, XPS_AREA));
}

void ref_sca_pb(int global, int SOURCE, int PACKETS, int BATCH, float SEED, float *ABS, float *SCA, float BG,
                float *PSPOS_xyz, float *PS, const int *LCELLS, const int *OFF, int *PAR, float *DENS,
                const float *DSC, const float *CSC, int NDIR, float *ODIRS_xyz, int npx, int npy, float MAP_DX,
                const float *CENTRE, float *ORA_xyz, float *ODE_xyz, float *OUT, float *ABU, float *OPT,
                float *XPS_NSIDE, float *XPS_SIDE, float *XPS_AREA, const int *roi_dim, float *roi_load) {
    int2 NPIX(npx, npy);
    float3 c(CENTRE[0], CENTRE[1], CENTRE[2]);
    REF_PARALLEL_FOR(global,
        refs::SimRAM_PB(SOURCE, PACKETS, BATCH, SEED, ABS, SCA, BG, (float3 *)PSPOS_xyz, PS, LCELLS, OFF, PAR, DENS,
                        DSC, CSC, NDIR, (float3 *)ODIRS_xyz, NPIX, MAP_DX, c, (float3 *)ORA_xyz, (float3 *)ODE_xyz,
                        OUT, ABU, OPT, XPS_NSIDE, XPS_SIDE, XPS_AREA, roi_dim, roi_load));
}

void ref_sca_hp(int global, int PACKETS, int BATCH, float SEED, float *ABS, float *SCA, const int *LCELLS,
                const int *OFF, int *PAR, float *DENS, const float *DSC, const float *CSC, int NDIR,
                float *ODIRS_xyz, int npx, int npy, float MAP_DX, const float *CENTRE, float *ORA_xyz,
                float *ODE_xyz, float *OUT, float *ABU, float *OPT, float *BG, float *HPBGP) {
    int2 NPIX(npx, npy);
    float3 c(CENTRE[0], CENTRE[1], CENTRE[2]);
    REF_PARALLEL_FOR(global,
        refs::SimRAM_HP(PACKETS, BATCH, SEED, ABS, SCA, LCELLS, OFF, PAR, DENS, DSC, CSC, NDIR, (float3 *)ODIRS_xyz,
                        NPIX, MAP_DX, c, (float3 *)ORA_xyz, (float3 *)ODE_xyz, OUT, ABU, OPT, BG, HPBGP));
}

void ref_sca_cl(int global, int SOURCE, int PACKETS, int BATCH, float SEED, float *ABS, float *SCA,
                const int *LCELLS, const int *OFF, int *PAR, float *DENS, float *EMIT, const float *DSC,
                const float *CSC, int NDIR, float *ODIRS_xyz, int npx, int npy, float MAP_DX, const float *CENTRE,
                float *ORA_xyz, float *ODE_xyz, float *OUT, float *OPT, float *ABU, float *EMWEI) {
    int2 NPIX(npx, npy);
    float3 c(CENTRE[0], CENTRE[1], CENTRE[2]);
    REF_PARALLEL_FOR(global,
        refs::SimRAM_CL(SOURCE, PACKETS, BATCH, SEED, ABS, SCA, LCELLS, OFF, PAR, DENS, EMIT, DSC, CSC, NDIR,
                        (float3 *)ODIRS_xyz, NPIX, MAP_DX, c, (float3 *)ORA_xyz, (float3 *)ODE_xyz, OUT, OPT, ABU,
                        EMWEI));
}

}  // extern "C"
