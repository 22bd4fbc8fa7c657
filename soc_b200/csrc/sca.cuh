// soc_b200 -- arguments of the scattered-light (peel-off) kernels.
#pragma once
#include "common.cuh"

struct ScaArgs {
    GridDesc G;
    float *out;                                            // [ndir*npy*npx]
    const float *__restrict__ opt, *__restrict__ dsc, *__restrict__ csc;
    const float *__restrict__ pspos, *__restrict__ ps;
    const float *__restrict__ xps_area;                    // unused (PS_METHOD 0/1 only), keeps emit.cuh generic
    const int *__restrict__ xps_nside, *__restrict__ xps_side;
    const float *__restrict__ odir, *__restrict__ ora, *__restrict__ ode;   // [ndir*3]; Healpix observer: odir[0..2] = position
    const float *__restrict__ hpbg, *__restrict__ hpbgp;   // SimRAM_HP: sky map [49152] and its cumulative probability
    const float *__restrict__ emit, *__restrict__ emwei;   // SimRAM_CL: emission and emission weights [cells]
    const float *__restrict__ abu, *__restrict__ scav;     // WITH_MSF: ABU[cells*ndust], SCA[ndust]; dsc/csc are [ndust*bins]
    float kabs, ksca, bg, map_dx;
    vec3 centre;
    int kind;                          // source: SIM_PS / SIM_BG / SIM_HP / SIM_CL (sim.cuh numbering)
    int flavour;                       // arithmetic flavour of the reference kernel: 0 SimRAM_PS, 1 SimRAM_PB, 2 SimRAM_HP, 3 SimRAM_CL
    int batch, global, ndir, npx, npy;
    int nside;                         // > 0: one Healpix image of this NSIDE seen from odir[0..2] (reference: NDIR = -NSIDE)
    int bins, no_ps, ps_method, with_abu, ffs;
    int hpbg_weighted, use_emweight, with_ali, with_msf, ndust, mirror;
    int hg_test;                       // SimRAM_HP / SimRAM_CL as shipped: `#ifdef HG_TEST` peel-off weight (soc_params.ref_quirks & 1)
    const int *__restrict__ nbr;       // octrees: neighbour table [6*cells] (linkwalk.cuh) or nullptr
    RoiDesc roi;                       // WITH_ROI_LOAD source (kind 4)
    int roi_nelem;
    long long nunits;
    int rank, world, max_steps, ref_geometry;
    int nav_hops;                      // octree navigation rounds (climb / cross / descend) per loop iteration
    int ev_batch;                      // production kernel: handle ray ends when this many lanes of the warp wait
    unsigned long long *counters;      // packets, steps, scatterings, stuck, peels
    unsigned long long *work;
    MwcLaunch mwc;
    PhiloxLaunch phx;
};

void launch_sca(const ScaArgs &S, int rng_mode, int blocks, int threads, cudaStream_t stream);
int  sca_blocks_per_sm(bool octree, bool dbl, int threads);
