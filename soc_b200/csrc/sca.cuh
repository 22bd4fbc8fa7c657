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
    const float *__restrict__ odir, *__restrict__ ora, *__restrict__ ode;   // [ndir*3]
    float kabs, ksca, bg, map_dx;
    vec3 centre;
    int kind, flavour, batch, global, ndir, npx, npy;
    int bins, no_ps, ps_method, with_abu, ffs;
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
