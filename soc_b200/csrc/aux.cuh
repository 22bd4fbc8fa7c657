// soc_b200 -- small grid-wide kernels: parent table, equilibrium temperature, emission.
#pragma once
#include "common.cuh"

void launch_parents(const GridDesc &G, int *par, cudaStream_t stream);
void launch_eq_temperature(const GridDesc &G, int level, float adhoc, float kE, float Emin, int NE, float factor,
                           float length, const float *ttt, const float *emit, float *tnew, cudaStream_t stream);
void launch_emission(int cells, float freq, float fabs_, float factor, float length, const float *t, float *emit,
                     cudaStream_t stream);
void launch_absorbed_add(float *fabs, const float *inten, int cells, int nfreq, int ifreq, cudaStream_t stream);
void launch_absorbed_scale(const GridDesc &G, float *fabs, int nfreq, float coeff0, float nnnlimit, cudaStream_t stream);
void launch_emission2(int c0, int c1, int nfreq, float factor, float length, const float *freq, const float *fabs_, const float *t,
                      float *emit, cudaStream_t stream);
#define SOC_MAX_DUSTS 32
void launch_build_opt(const float *abu, float *opt, long long cells, int ndust, int first, int single_abu, int half,
                      const float *kabs, const float *ksca, cudaStream_t stream);   // OPT from ABU on the device
void launch_split_absorbed(int idust, long long cells, int nfreq, int ndust, const double *rabs, const float *abu, const float *in, float *out,
                           cudaStream_t stream);   // kernel_A2E_MABU_aux.c split_absorbed
void launch_half_to_float(const void *src, float *dst, long long n, cudaStream_t stream);
void launch_neighbours(const GridDesc &G, int *nbr, cudaStream_t stream);      // neighbour table of linkwalk.cuh, [6*cells]
