// soc_b200 -- arguments of the map ray-tracer kernels.
#pragma once
#include "common.cuh"

struct MapArgs {
    GridDesc G;
    float *map, *savetau;                                  // [npy*npx] or [12*nside^2]
    const float *__restrict__ emit, *__restrict__ opt;     // [cells], [2*cells]
    vec3 dir, ra, de, centre, intobs;
    float map_dx, kabs, ksca, length;
    int npx, npy, nside, with_abu, level_threshold, save_colden;
    RoiDesc roi;                                           // ROI_MAP (flags & 4): only cells inside ROI emit
    int maph_literal;                                      // per-level maps: the file's own Index() (soc_params.ref_quirks & 2)
    int map_interpolation;                                 // MAP_INTERPOLATION 0/1/2 (orthographic and perspective maps)
    unsigned long long *counters;
};

void launch_mapping(const MapArgs &M, bool healpix, cudaStream_t stream);
void launch_mapping_levels(const MapArgs &M, cudaStream_t stream);    // map = [levels*npy*npx], savetau = column density or nullptr
void launch_pstau(const MapArgs &M, int no, const float *pspos, float *colden, float *tau, cudaStream_t stream);
