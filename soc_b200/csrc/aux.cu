// soc_b200 -- small grid-wide kernels around the packet kernels.
//   Parents        kernel_ASOC_aux.c:688-718   PAR[child] = parent, from the links stored in DENS
//   EqTemperature  kernel_ASOC_aux.c:745-787   absorbed energy -> equilibrium dust temperature via the E->T table
//   Emission       kernel_ASOC_aux.c:793-807   modified black body emission of every cell at one frequency
// All three stream through the cell arrays once (grid-stride, coalesced).
#include "aux.cuh"
#include <cuda_fp16.h>
#include "linkwalk.cuh"

namespace {

__global__ void parents_kernel(GridDesc G, int *par, int level) {
    const int n = G.lcells[level];
    for (int ipar = blockIdx.x * blockDim.x + threadIdx.x; ipar < n; ipar += gridDim.x * blockDim.x) {
        float link = G.dens[G.off[level] + ipar];
        if (link < 1.0e-10f) {
            int first = link_index(link);
            int *dst = par + (G.off[level + 1] - G.nxyz + first);
            #pragma unroll
            for (int i = 0; i < 8; i++) dst[i] = ipar;
        }
    }
}

__global__ void eq_temperature_kernel(GridDesc G, int level, float adhoc, float kE, float Emin, int NE, float factor,
                                      float length, const float *__restrict__ ttt, const float *__restrict__ emit,
                                      float *__restrict__ tnew) {
    const float scale = (6.62607e-27f * factor) / length;
    const float oplgkE = 1.0f / log10f(kE);
    const float l8 = powf(8.0f, (float)level);
    const int n = G.lcells[level];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int ind = G.off[level] + i;
        float rho = G.dens[ind];
        float Ein = (scale / adhoc) * emit[ind] * l8 / rho;
        int iE = clampi((int)floorf(oplgkE * log10f(Ein / Emin)), 0, NE - 2);
        float wi = (Emin * powf(kE, (float)(iE + 1)) - Ein) / (Emin * powf(kE, (float)iE) * (kE - 1.0f));
        tnew[ind] = (rho > 1.0e-7f) ? clampf(wi * ttt[iE] + (1.0f - wi) * ttt[iE + 1], 3.0f, 1600.0f) : 10.0f;
    }
}

__global__ void emission_kernel(int cells, float freq, float fabs_, float factor, float length,
                                const float *__restrict__ t, float *__restrict__ emit) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += gridDim.x * blockDim.x)
        emit[i] = (2.79639459e-20f * factor) * fabs_ * (freq * freq / (expf(4.7995074e-11f * freq / t[i]) - 1.0f)) / length;
}

// Emission2 (kernel_ASOC_aux.c:862-888): cells [c0,c1[ x nfreq frequencies in one pass, EMIT[(icell-c0)*nfreq + ifreq] -- the
// layout of the emitted file, written coalesced (one thread per element)
__global__ void emission2_kernel(int c0, long long total, int nfreq, float factor, float length, const float *__restrict__ freq,
                                 const float *__restrict__ fabs_, const float *__restrict__ t, float *__restrict__ emit) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int icell = c0 + (int)(e / nfreq), ifreq = (int)(e % nfreq);
        const float f = freq[ifreq];
        emit[e] = (2.79639459e-20f * factor) * fabs_[ifreq] * (f * f / (expf(4.7995074e-11f * f / t[icell]) - 1.0f)) / length;
    }
}

// FABS[cell, ifreq] += INT[cell]: coalesced read, one 4-byte write per cell row of the [cells, nfreq] array
__global__ void absorbed_add_kernel(float *__restrict__ fabs, const float *__restrict__ inten, int cells, int nfreq, int ifreq) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += gridDim.x * blockDim.x) {
        float v = inten[i];
        if (v != 0.0f) fabs[(size_t)i * nfreq + ifreq] += v;
    }
}

// absorbed-file scaling of ASOC.py:2793-2809: photons -> FACTOR * photons per H; parents / thin cells -> -1e20
__global__ void absorbed_scale_kernel(GridDesc G, float *__restrict__ fabs, int nfreq, float coeff0, float nnnlimit) {
    const size_t total = (size_t)G.cells * nfreq;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int cell = (int)(e / nfreq);
        int level = 0;
        while (level + 1 < G.levels && cell >= G.off[level + 1]) level++;
        const float rho = G.dens[cell];
        const float k = coeff0 * __int_as_float((127 + 3 * level) << 23) / rho;         // 8^level = 2^(3 level)
        fabs[e] = (rho <= nnnlimit) ? -1.0e20f : fabs[e] * k;
    }
}

inline int grid_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    const long long cap = 148LL * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

void launch_parents(const GridDesc &G, int *par, cudaStream_t stream) {
    for (int level = 0; level < G.levels - 1; level++)
        parents_kernel<<<grid_for(G.lcells[level], 256), 256, 0, stream>>>(G, par, level);
}
void launch_eq_temperature(const GridDesc &G, int level, float adhoc, float kE, float Emin, int NE, float factor,
                           float length, const float *ttt, const float *emit, float *tnew, cudaStream_t stream) {
    eq_temperature_kernel<<<grid_for(G.lcells[level], 256), 256, 0, stream>>>(G, level, adhoc, kE, Emin, NE, factor,
                                                                             length, ttt, emit, tnew);
}
void launch_emission(int cells, float freq, float fabs_, float factor, float length, const float *t, float *emit,
                     cudaStream_t stream) {
    emission_kernel<<<grid_for(cells, 256), 256, 0, stream>>>(cells, freq, fabs_, factor, length, t, emit);
}
void launch_absorbed_add(float *fabs, const float *inten, int cells, int nfreq, int ifreq, cudaStream_t stream) {
    absorbed_add_kernel<<<grid_for(cells, 256), 256, 0, stream>>>(fabs, inten, cells, nfreq, ifreq);
}
void launch_absorbed_scale(const GridDesc &G, float *fabs, int nfreq, float coeff0, float nnnlimit, cudaStream_t stream) {
    absorbed_scale_kernel<<<grid_for((long long)G.cells * nfreq, 256), 256, 0, stream>>>(G, fabs, nfreq, coeff0, nnnlimit);
}

void launch_emission2(int c0, int c1, int nfreq, float factor, float length, const float *freq, const float *fabs_, const float *t,
                      float *emit, cudaStream_t stream) {
    const long long total = (long long)(c1 - c0) * nfreq;
    long long b = (total + 255) / 256;
    const long long cap = 148LL * 16;
    emission2_kernel<<<(int)(b < 1 ? 1 : (b > cap ? cap : b)), 256, 0, stream>>>(c0, total, nfreq, factor, length, freq, fabs_, t, emit);
}

// OPT_IS_HALF: widen the uploaded half-precision opacities (vload_half in the reference, kernel_ASOC_aux.c:12-16)
__global__ void half_to_float_kernel(const __half *__restrict__ src, float *__restrict__ dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = __half2float(src[i]);
}
void launch_half_to_float(const void *src, float *dst, long long n, cudaStream_t stream) {
    long long b = (n + 255) / 256;
    const long long cap = 148LL * 16;
    half_to_float_kernel<<<(int)(b < 1 ? 1 : (b > cap ? cap : b)), 256, 0, stream>>>(reinterpret_cast<const __half *>(src), dst, n);
}

// OPT[cells,2] = per-cell (KABS, KSCA) from the abundances: sum over the dust species of ABU[cell, d] * K[d], in the
// reference's order and precision (ASOC.py:1146-1161: float32 products added one species after the other; SINGLE_ABU:
// a*K0 + (1-a)*K1; OPT_IS_HALF: the result rounded to half precision).  Nothing of CELLS size crosses PCIe per frequency.
struct OptCoeffs { float kabs[SOC_MAX_DUSTS], ksca[SOC_MAX_DUSTS]; };
__global__ void __launch_bounds__(256) build_opt_kernel(const float *__restrict__ abu, float2 *__restrict__ opt, long long cells, int ndust,
                                                        int first, int single_abu, int half, const OptCoeffs K) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
        float a = 0.0f, s = 0.0f;
        if (single_abu) {
            const float x = abu[i * ndust], y = __fsub_rn(1.0f, x);
            a = __fadd_rn(__fmul_rn(x, K.kabs[0]), __fmul_rn(y, K.kabs[1]));
            s = __fadd_rn(__fmul_rn(x, K.ksca[0]), __fmul_rn(y, K.ksca[1]));
        } else {
            for (int d = first; d < ndust; d++) {
                const float x = abu[i * ndust + d];
                a = __fadd_rn(a, __fmul_rn(x, K.kabs[d]));
                s = __fadd_rn(s, __fmul_rn(x, K.ksca[d]));
            }
        }
        if (half) { a = __half2float(__float2half_rn(a)); s = __half2float(__float2half_rn(s)); }
        opt[i] = make_float2(a, s);
    }
}
void launch_build_opt(const float *abu, float *opt, long long cells, int ndust, int first, int single_abu, int half,
                      const float *kabs, const float *ksca, cudaStream_t stream) {
    OptCoeffs K;
    for (int d = 0; d < SOC_MAX_DUSTS; d++) { K.kabs[d] = d < ndust ? kabs[d] : 0.0f; K.ksca[d] = d < ndust ? ksca[d] : 0.0f; }
    long long b = (cells + 255) / 256;
    const long long cap = 148LL * 16;
    build_opt_kernel<<<(int)(b < 1 ? 1 : (b > cap ? cap : b)), 256, 0, stream>>>(abu, reinterpret_cast<float2 *>(opt), cells, ndust, first, single_abu,
                                                                                  half, K);
}

// split_absorbed (kernel_A2E_MABU_aux.c:3-24, A2E_MABU.py:700-705): the absorbed array handed to the dust solver of species
// IDUST: OUT[cell, f] = IN[cell, f] * RABS[f, IDUST] / sum_d ABU[cell, d] * RABS[f, d].  Same precision as the reference: RABS
// double, `den` a float that takes the double products one by one, quotient in double, result float.  One thread per
// (cell, frequency); both arrays are streamed once.
__global__ void __launch_bounds__(256) split_absorbed_kernel(int idust, long long cells, int nfreq, int ndust, const double *__restrict__ rabs,
                                                             const float *__restrict__ abu, const float *__restrict__ in, float *__restrict__ out) {
    const long long n = cells * nfreq, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long icell = i / nfreq;
        const int ifreq = (int)(i - icell * nfreq);
        float den = 0.0f;
        for (int d = 0; d < ndust; d++) den = (float)__dadd_rn((double)den, __dmul_rn((double)abu[icell * ndust + d], rabs[ifreq * ndust + d]));
        out[i] = (float)__ddiv_rn(__dmul_rn((double)in[i], rabs[ifreq * ndust + idust]), (double)den);
    }
}
void launch_split_absorbed(int idust, long long cells, int nfreq, int ndust, const double *rabs, const float *abu, const float *in, float *out,
                           cudaStream_t stream) {
    long long b = (cells * nfreq + 255) / 256;
    const long long cap = 148LL * 16;
    split_absorbed_kernel<<<(int)(b < 1 ? 1 : (b > cap ? cap : b)), 256, 0, stream>>>(idust, cells, nfreq, ndust, rabs, abu, in, out);
}

// Neighbour table of linkwalk.cuh: NBR[6*cell + face], face = 2*axis + (towards +axis).  For every cell the cell of
// the same level behind the face if the hierarchy has it (possibly refined further), else the coarser leaf covering it.
__global__ void neighbours_kernel(GridDesc G, int *__restrict__ nbr) {
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < G.cells; g += gridDim.x * blockDim.x) {
        int level = 0;
        while (level + 1 < G.levels && g >= G.off[level + 1]) level++;
        // integer coordinates of the cell at its own level
        int cx = 0, cy = 0, cz = 0, i = g - G.off[level];
        for (int l = level, sh = 0; l > 0; l--, sh++) {
            cx |= (i & 1) << sh; cy |= ((i >> 1) & 1) << sh; cz |= ((i >> 2) & 1) << sh;
            i = G.par[G.off[l] + i - G.nxyz];
        }
        cx |= (i % G.nx) << level; cy |= ((i / G.nx) % G.ny) << level; cz |= (i / (G.nx * G.ny)) << level;
        for (int f = 0; f < 6; f++) {
            const int ax = f >> 1, sgn = (f & 1) ? 1 : -1;
            const int tx = cx + (ax == 0 ? sgn : 0), ty = cy + (ax == 1 ? sgn : 0), tz = cz + (ax == 2 ? sgn : 0);
            int e = -1;
            float erho = 0.0f;
            if (tx >= 0 && ty >= 0 && tz >= 0 && tx < (G.nx << level) && ty < (G.ny << level) && tz < (G.nz << level)) {
                int lev = 0;
                int c = ((tz >> level) * G.ny + (ty >> level)) * G.nx + (tx >> level);
                float rho = G.dens[c];
                while (!is_leaf(rho) && lev < level) {
                    const int base = link_index(rho);
                    lev++;
                    const int sh = level - lev;
                    c = G.off[lev] + base + (((tz >> sh) & 1) << 2 | ((ty >> sh) & 1) << 1 | ((tx >> sh) & 1));
                    rho = G.dens[c];
                }
                e = (lev << SOC_NBR_LEVEL_SHIFT) | c;
                erho = rho;
            }
            reinterpret_cast<int2 *>(nbr)[6 * (size_t)g + f] = make_int2(e, __float_as_int(erho));
        }
    }
}
void launch_neighbours(const GridDesc &G, int *nbr, cudaStream_t stream) {
    long long b = ((long long)G.cells + 255) / 256;
    const long long cap = 148LL * 16;
    neighbours_kernel<<<(int)(b < 1 ? 1 : (b > cap ? cap : b)), 256, 0, stream>>>(G, nbr);
}
