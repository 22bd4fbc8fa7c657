// soc_b200 -- device-side building blocks shared by the packet, map and scattered-light kernels.
//
// Geometry is evaluated with explicitly rounded single operations (__fadd_rn/__fmul_rn/__fdiv_rn are
// never contracted into FMAs) so that cell-boundary decisions are reproducible against the CPU oracle;
// everything else is left to the compiler.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "soc_b200.h"

#define SOC_MAX_LEVELS 16

// kernel_ASOC_aux.c:5-9, 99-114 (simulation) and kernel_ASOC_map.c:10-18 (map) constants
#define SOC_TWOPI    6.28318531f
#define SOC_TAULIM   5.0e-4f
#define SOC_PIHALF   1.5707963268f
#define SOC_TWOTHIRD 0.6666666667f
#define SOC_PI       3.1415926535897f
#define SOC_PEPS     1.0e-4f
#define SOC_DEPS     5.0e-5f
#define SOC_MAP_EPS  2.5e-4f
#define SOC_MAP_PEPS 5.0e-4f
#define SOC_MAP_PI    3.1415926536f
#define SOC_MAP_TWOPI 6.2831853072f

struct GridDesc {
    int nx, ny, nz, levels, cells, nxyz, area;
    int dbl_sim;                 // Index() works in double: NX > DIMLIM (kernel_ASOC_aux.c:25-37)
    int dbl_map;                 // map Index() works in double: NX > 100 (kernel_ASOC_map.c:302)
    int off[SOC_MAX_LEVELS];
    int lcells[SOC_MAX_LEVELS];
    const float *__restrict__ dens;
    const int *__restrict__ par;
};

struct vec3 { float x, y, z; };

// ---- exactly rounded helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
// fmod(x, 1) for finite x (keeps the sign of x like C fmod); exact in floating point
__device__ __forceinline__ float  fmod1(float x)  { return __fsub_rn(x, truncf(x)); }
__device__ __forceinline__ double fmod1(double x) { return __dsub_rn(x, trunc(x)); }
__device__ __forceinline__ float  floor_r(float x)  { return floorf(x); }
__device__ __forceinline__ double floor_r(double x) { return floor(x); }
__device__ __forceinline__ bool is_leaf(float v) { return v > 0.0f; }
__device__ __forceinline__ int link_index(float v) { return (int)(__float_as_uint(v) & 0x7FFFFFFFu); }
__device__ __forceinline__ vec3 normalize3(vec3 a) {
    float l = sqrtf(xadd(xadd(xmul(a.x, a.x), xmul(a.y, a.y)), xmul(a.z, a.z)));
    vec3 r = { xdiv(a.x, l), xdiv(a.y, l), xdiv(a.z, l) };
    return r;
}
// a.x*b.x + a.y*b.y + a.z*b.z evaluated left to right without contraction
__device__ __forceinline__ float dot3(const vec3 &a, const vec3 &b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
// 2^x through the SFU (ex2.approx.ftz: max relative error 2^-22)
__device__ __forceinline__ float exp2f_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 1/x through the SFU (rcp.approx.ftz: max relative error 2^-23), for normal x
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ---- random numbers --------------------------------------------------------------------------------------
// MWC64X (mwc64x_rng.cl, skip_mwc.cl): modulus M = A*2^32 - 1.
#define MWC_A 4294883355ull
#define MWC_M 18446383549859758079ull
#define MWC_R 360523849793537ull          // 2^64 mod M
#define MWC_BASEID 4077358422479273989ull

__host__ __device__ __forceinline__ void mul64wide(uint64_t a, uint64_t b, uint64_t &hi, uint64_t &lo) {
#ifdef __CUDA_ARCH__
    lo = a * b; hi = __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128)a * b; lo = (uint64_t)p; hi = (uint64_t)(p >> 64);
#endif
}
// (a*b) mod M for a,b < M: fold the high word with 2^64 = R (mod M) until it vanishes
__host__ __device__ inline uint64_t mwc_mulmod(uint64_t a, uint64_t b) {
    uint64_t hi, lo;
    mul64wide(a, b, hi, lo);
    while (hi != 0) {
        uint64_t h2, l2;
        mul64wide(hi, MWC_R, h2, l2);
        uint64_t s = l2 + lo;
        hi = h2 + (s < lo ? 1 : 0);
        lo = s;
    }
    while (lo >= MWC_M) lo -= MWC_M;
    return lo;
}
__host__ __device__ inline uint64_t mwc_powmod(uint64_t a, uint64_t e) {
    uint64_t sqr = a, acc = 1;
    while (e) { if (e & 1) acc = mwc_mulmod(acc, sqr); sqr = mwc_mulmod(sqr, sqr); e >>= 1; }
    return acc;
}

// Per-launch constants of the reference stream layout: stream id*2^38 after a base offset derived from
// the float seed (kernel_ASOC.c:74-77).  pow2k[k] = A^(2^(38+k)) mod M, base_state = BASEID*A^base mod M.
struct MwcLaunch {
    uint64_t base_state;
    uint64_t base_offset;
    uint64_t pow2k[26];
};

struct RngMwc {
    uint32_t x, c;
    __device__ __forceinline__ void seed(const MwcLaunch &L, uint64_t id) {
        uint64_t s;
        if (id < (1ull << 26)) {
            s = L.base_state;
            #pragma unroll 1
            for (int k = 0; id != 0; k++, id >>= 1) if (id & 1) s = mwc_mulmod(s, L.pow2k[k]);
        } else {   // the exponent wraps modulo 2^64 exactly like the reference's ulong arithmetic
            s = mwc_mulmod(MWC_BASEID, mwc_powmod(MWC_A, L.base_offset + id * 274877906944ull));
        }
        x = (uint32_t)(s / MWC_A); c = (uint32_t)(s % MWC_A);
    }
    __device__ __forceinline__ uint32_t next() {
        uint32_t r = x ^ c;
        uint64_t t = MWC_A * (uint64_t)x + c;
        x = (uint32_t)t; c = (uint32_t)(t >> 32);
        return r;
    }
    // Rand(): uint/4294967295.0f (kernel_ASOC_aux.c:127); the divisor rounds to 2^32 in float
    __device__ __forceinline__ float uniform() { return __uint2float_rn(next()) * 2.3283064365386963e-10f; }
};

// Philox4x32-10 (Salmon et al. 2011): counter = (packet lo, packet hi, draw block, stream tag), key from the seed.
struct PhiloxLaunch { uint32_t k0, k1, tag, pad; };
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t a, uint32_t b,
                                              uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3) {
    #pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ a; c1 = lo1; c2 = hi0 ^ c3 ^ b; c3 = lo0;
        a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}
// uint32 -> [0,1] like Rand() of the reference (kernel_ASOC_aux.c:127)
__device__ __forceinline__ float u32_to_unit(uint32_t v) { return __uint2float_rn(v) * 2.3283064365386963e-10f; }
struct RngPhilox {
    uint32_t p0, p1, blk, tag, k0, k1;
    uint32_t buf0, buf1, buf2, buf3;
    int have;
    __device__ __forceinline__ void seed(const PhiloxLaunch &L, uint64_t id) {
        p0 = (uint32_t)id; p1 = (uint32_t)(id >> 32); blk = 0; tag = L.tag; k0 = L.k0; k1 = L.k1; have = 0;
    }
    __device__ __forceinline__ void refill() {
        philox4x32_10(p0, p1, blk, tag, k0, k1, buf0, buf1, buf2, buf3);
        blk++; have = 4;
    }
    __device__ __forceinline__ uint32_t next() {
        if (have == 0) refill();
        have--;
        uint32_t r = buf0; buf0 = buf1; buf1 = buf2; buf2 = buf3;
        return r;
    }
    __device__ __forceinline__ float uniform() { return __uint2float_rn(next()) * 2.3283064365386963e-10f; }
};

struct RngBlock {                    // four uniforms from one Philox block
    uint32_t a, b, c, d; int k;
    __device__ __forceinline__ RngBlock(const PhiloxLaunch &L, unsigned long long id, unsigned blk) {
        philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), blk, L.tag, L.k0, L.k1, a, b, c, d);
        k = 0;
    }
    __device__ __forceinline__ float uniform() {
        uint32_t v = (k == 0) ? a : ((k == 1) ? b : ((k == 2) ? c : d));
        k++;
        return u32_to_unit(v);
    }
};

// ---- grid navigation ------------------------------------------------------------------------------------
// IndexG (kernel_ASOC_aux.c:131-165 / kernel_ASOC_map.c:187-220): global position -> leaf (level, ind);
// converts p to level-local coordinates.  Returns the leaf density through `rho`.
template <bool OCT, bool MAPK>
__device__ __forceinline__ void index_global(const GridDesc &G, vec3 &p, int &level, int &ind, float &rho) {
    ind = -1;
    if (!(p.x > 0.0f && p.y > 0.0f && p.z > 0.0f && p.x < G.nx && p.y < G.ny && p.z < G.nz)) return;
    level = 0;
    ind = (int)floorf(p.z) * G.nx * G.ny + (int)floorf(p.y) * G.nx + (int)floorf(p.x);
    rho = G.dens[ind];
    if (!OCT) return;
    if (is_leaf(rho)) return;
    p.x = xmul(2.0f, fmod1(p.x)); p.y = xmul(2.0f, fmod1(p.y)); p.z = xmul(2.0f, fmod1(p.z));
    for (;;) {
        ind = link_index(rho);
        level++;
        ind += 4 * (int)floorf(p.z) + 2 * (int)floorf(p.y) + (int)floorf(p.x);
        rho = G.dens[G.off[level] + ind];
        if (is_leaf(rho)) return;
        if (MAPK) {     // same values, the map kernel just writes the update differently
            p.x = xmul(xsub(p.x, floorf(p.x)), 2.0f); p.y = xmul(xsub(p.y, floorf(p.y)), 2.0f); p.z = xmul(xsub(p.z, floorf(p.z)), 2.0f);
        } else {
            p.x = xmul(2.0f, fmod1(p.x)); p.y = xmul(2.0f, fmod1(p.y)); p.z = xmul(2.0f, fmod1(p.z));
        }
    }
}

// Index(): the leaf containing a position that has just left cell (level, ind)
// (kernel_ASOC_aux.c:198-278; map flavour kernel_ASOC_map.c:294-379 with its z<=0 containment test).
// MAPH_LIT: the copy of Index() in kernel_ASOC_map_H.c:216-291, which returns without storing the root coordinates when the
// ray climbs into a root-grid leaf (:250) -- only on request (soc_params.ref_quirks & 2).
template <typename REAL, bool MAPK, bool MAPH_LIT = false>
__device__ __forceinline__ void index_octree(const GridDesc &G, vec3 &pos, int &level, int &ind, float &rho) {
    REAL px = pos.x, py = pos.y, pz = pos.z;
    const int NX = G.nx, NY = G.ny, NZ = G.nz;
    if (level == 0) {
        if (!(pos.x > 0.0f && pos.x < NX && pos.y > 0.0f && pos.y < NY && pos.z > 0.0f && pos.z < NZ)) { ind = -1; return; }
        ind = (int)floorf(pos.z) * NX * NY + (int)floorf(pos.y) * NX + (int)floorf(pos.x);
        rho = G.dens[ind];
        if (is_leaf(rho)) return;
    } else {
        bool need_rho = true;
        while (level > 0) {
            ind = G.par[G.off[level] + ind - G.nxyz]; level--;
            px = xmul(px, (REAL)0.5); py = xmul(py, (REAL)0.5); pz = xmul(pz, (REAL)0.5);
            if (level == 0) {
                px = xadd(px, (REAL)(ind % NX)); py = xadd(py, (REAL)((ind / NX) % NY)); pz = xadd(pz, (REAL)(ind / (NX * NY)));
                if (!(px > 0 && px < NX && py > 0 && py < NY && pz > 0 && pz < NZ)) {
                    ind = -1; pos.x = (float)px; pos.y = (float)py; pos.z = (float)pz; return;
                }
                ind = (int)floor_r(pz) * NX * NY + (int)floor_r(py) * NX + (int)floor_r(px);
                rho = G.dens[ind]; need_rho = false;
                if (is_leaf(rho)) { if (!MAPH_LIT) { pos.x = (float)px; pos.y = (float)py; pos.z = (float)pz; } return; }
                break;
            } else {
                int sid = ind & 7;
                px = xadd(px, (REAL)(sid & 1)); py = xadd(py, (REAL)((sid >> 1) & 1)); pz = xadd(pz, (REAL)(sid >> 2));
                if (!MAPK) {
                    if (px >= 0 && px <= 2 && py >= 0 && py <= 2 && pz >= 0 && pz <= 2) {
                        ind += -sid + 4 * (int)floor_r(pz) + 2 * (int)floor_r(py) + (int)floor_r(px);
                        break;
                    }
                } else {
                    if (px >= 0 && px <= 2 && py >= 0 && py <= 2 && pz >= 0 && pz <= 0) break;
                }
            }
        }
        if (need_rho) rho = G.dens[G.off[level] + ind];
    }
    while (!is_leaf(rho)) {
        px = xmul((REAL)2, fmod1(px)); py = xmul((REAL)2, fmod1(py)); pz = xmul((REAL)2, fmod1(pz));
        ind = link_index(rho);
        level++;
        ind += 4 * (int)floor_r(pz) + 2 * (int)floor_r(py) + (int)floor_r(px);
        rho = G.dens[G.off[level] + ind];
    }
    pos.x = (float)px; pos.y = (float)py; pos.z = (float)pz;
}

// GetStep (kernel_ASOC_aux.c:282-315, kernel_ASOC_map.c:387-429): distance to the next cell face with PEPS
// overshoot, advance the local position, look up the neighbour.  Returns the step in root-grid units and the
// density of the cell entered (`rho`, undefined when ind<0).
template <bool OCT, bool DBL, bool MAPK, bool MAPH_LIT = false>
__device__ __forceinline__ float get_step(const GridDesc &G, vec3 &p, const vec3 &d, int &level, int &ind, float &rho) {
    const float peps = MAPK ? SOC_MAP_PEPS : SOC_PEPS;
    float dx = (d.x > 0.0f) ? xdiv(xsub(xadd(1.0f, peps), fmod1(p.x)), d.x) : xdiv(xsub(-peps, fmod1(p.x)), d.x);
    float dy = (d.y > 0.0f) ? xdiv(xsub(xadd(1.0f, peps), fmod1(p.y)), d.y) : xdiv(xsub(-peps, fmod1(p.y)), d.y);
    float dz = (d.z > 0.0f) ? xdiv(xsub(xadd(1.0f, peps), fmod1(p.z)), d.z) : xdiv(xsub(-peps, fmod1(p.z)), d.z);
    dx = fminf(dx, fminf(dy, dz));
    p.x = xadd(p.x, xmul(dx, d.x)); p.y = xadd(p.y, xmul(dx, d.y)); p.z = xadd(p.z, xmul(dx, d.z));
    if (OCT) {
        dx = ldexpf(dx, -level);
        if (DBL) index_octree<double, MAPK, MAPH_LIT>(G, p, level, ind, rho);
        else     index_octree<float, MAPK, MAPH_LIT>(G, p, level, ind, rho);
    } else {
        if (!(p.x > 0.0f && p.x < G.nx && p.y > 0.0f && p.y < G.ny && p.z > 0.0f && p.z < G.nz)) { ind = -1; }
        else {
            ind = (int)floorf(p.z) * G.nx * G.ny + (int)floorf(p.y) * G.nx + (int)floorf(p.x);
            rho = G.dens[ind];
        }
    }
    return dx;
}

// RootPos (kernel_ASOC_aux.c:169-190)
__device__ __forceinline__ void root_position(const GridDesc &G, vec3 &p, int level, int ind) {
    while (level > 0) {
        ind = G.par[G.off[level] + ind - G.nxyz]; level--;
        p.x = xmul(p.x, 0.5f); p.y = xmul(p.y, 0.5f); p.z = xmul(p.z, 0.5f);
        if (level == 0) {
            p.x = xadd(p.x, (float)(ind % G.nx)); p.y = xadd(p.y, (float)((ind / G.nx) % G.ny)); p.z = xadd(p.z, (float)(ind / (G.nx * G.ny)));
        } else {
            int sid = ind & 7;
            p.x = xadd(p.x, (float)(sid & 1)); p.y = xadd(p.y, (float)((sid >> 1) & 1)); p.z = xadd(p.z, (float)(sid >> 2));
        }
    }
}

// Surface (kernel_ASOC_aux.c:912-940): step an outside point onto the cloud surface
__device__ __forceinline__ void to_surface(const GridDesc &G, vec3 &p, const vec3 &d) {
    float dx, dy, dz;
    if (d.x > 0.0f) dx = (p.x < 0.0f) ? xdiv(xsub(SOC_PEPS, p.x), d.x) : -1.0e10f;
    else            dx = (p.x > G.nx) ? xdiv(xsub(xsub((float)G.nx, SOC_PEPS), p.x), d.x) : -1.0e10f;
    if (d.y > 0.0f) dy = (p.y < 0.0f) ? xdiv(xsub(SOC_PEPS, p.y), d.y) : -1.0e10f;
    else            dy = (p.y > G.ny) ? xdiv(xsub(xsub((float)G.ny, SOC_PEPS), p.y), d.y) : -1.0e10f;
    if (d.z > 0.0f) dz = (p.z < 0.0f) ? xdiv(xsub(SOC_PEPS, p.z), d.z) : -1.0e10f;
    else            dz = (p.z > G.nz) ? xdiv(xsub(xsub((float)G.nz, SOC_PEPS), p.z), d.z) : -1.0e10f;
    dx = fmaxf(dx, fmaxf(dy, dz));
    p.x = xadd(p.x, xmul(dx, d.x)); p.y = xadd(p.y, xmul(dx, d.y)); p.z = xadd(p.z, xmul(dx, d.z));
}

// ---- region of interest ------------------------------------------------------------------------------------------
struct RoiDesc {
    int flags;                 // 1 WITH_ROI_LOAD, 2 WITH_ROI_SAVE, 4 ROI_MAP
    int lim[6];                // [x0,x1,y0,y1,z0,z1], inclusive root-cell limits
    int step, nside;           // ROI_STEP, ROI_NSIDE
    int dim[3];                // dimensions of the loaded ROI file
    const float *__restrict__ load;
    float *save;
};
// InRoi (kernel_ASOC_aux.c:1031-1048 / kernel_ASOC_map.c:37-56): is the root ancestor of cell (level, ind) inside the box?
template <bool OCT>
__device__ __forceinline__ bool in_roi(const GridDesc &G, const RoiDesc &R, int level, int ind) {
    int i = ind;
    if (OCT) for (int k = level; k > 0; k--) i = G.par[G.off[k] + i - G.nxyz];
    const int k = i / (G.nx * G.ny), j = (i / G.nx) % G.ny;
    i = i % G.nx;
    return i >= R.lim[0] && i <= R.lim[1] && j >= R.lim[2] && j <= R.lim[3] && k >= R.lim[4] && k <= R.lim[5];
}
__device__ __forceinline__ bool in_roi_xyz(const RoiDesc &R, int i, int j, int k) {
    return i >= R.lim[0] && i <= R.lim[1] && j >= R.lim[2] && j <= R.lim[3] && k >= R.lim[4] && k <= R.lim[5];
}
__device__ inline int ang2pix_ring(int nside, float phi, float theta);
// A packet has stepped into ROI at root-grid position `rp` moving along `d`: ROI_SAVE[surface element, direction pixel] += photons
// (kernel_ASOC.c:617-642, 1510-1535)
__device__ __forceinline__ void roi_save_add(const RoiDesc &R, const vec3 &rp, const vec3 &d, float photons) {
    const int RNX = (R.lim[1] - R.lim[0] + 1) * R.step, RNY = (R.lim[3] - R.lim[2] + 1) * R.step, RNZ = (R.lim[5] - R.lim[4] + 1) * R.step;
    const float st = (float)R.step;
    int ii = 0, jj;
    if (rp.x < xadd((float)R.lim[0], 1.0e-3f) || rp.x > xadd((float)R.lim[1], 0.999f)) {
        ii = clampi((int)floorf(xmul(xsub(rp.y, (float)R.lim[2]), st)), 0, RNY - 1); jj = clampi((int)floorf(xmul(xsub(rp.z, (float)R.lim[4]), st)), 0, RNZ - 1);
        ii = ii + RNY * jj;
    }
    if (rp.y < xadd((float)R.lim[2], 1.0e-3f) || rp.y > xadd((float)R.lim[3], 0.999f)) {
        ii = clampi((int)floorf(xmul(xsub(rp.x, (float)R.lim[0]), st)), 0, RNX - 1); jj = clampi((int)floorf(xmul(xsub(rp.z, (float)R.lim[4]), st)), 0, RNZ - 1);
        ii = RNY * RNZ + ii + RNX * jj;
    }
    if (rp.z < xadd((float)R.lim[4], 1.0e-3f) || rp.z > xadd((float)R.lim[5], 0.999f)) {
        ii = clampi((int)floorf(xmul(xsub(rp.x, (float)R.lim[0]), st)), 0, RNX - 1); jj = clampi((int)floorf(xmul(xsub(rp.y, (float)R.lim[2]), st)), 0, RNY - 1);
        ii = RNY * RNZ + RNX * RNZ + ii + RNX * jj;
    }
    const float theta = acosf(d.z), phi = atan2f(d.y, d.x);
    jj = ang2pix_ring(R.nside, phi, theta);
    ii = clampi(ii, 0, RNX * RNY + RNY * RNZ + RNZ * RNX - 1);
    jj = clampi(jj, 0, 12 * R.nside * R.nside - 1);
    atomicAdd(&R.save[(size_t)ii * 12 * R.nside * R.nside + jj], photons);
}

// Mirror (kernel_ASOC_aux.c:1054-1083), literally: every enabled border negates its direction component whether or
// not it was crossed (`if (c) a ; b ;`).  Used by the parity / reference-geometry kernels; the production kernels
// reflect only the crossed border.
template <bool OCT>
__device__ __forceinline__ void mirror_literal(const GridDesc &G, int mask, vec3 &p, vec3 &d, int &level, int &ind, float &rho) {
    const float EPS = 5.0e-4f;
    if (mask & 1)  { if (p.x < 0.0f)         p.x = EPS;                       d.x = -d.x; index_global<OCT, false>(G, p, level, ind, rho); }
    if (mask & 2)  { if (p.x > (float)G.nx)  p.x = xsub((float)G.nx, EPS);    d.x = -d.x; index_global<OCT, false>(G, p, level, ind, rho); }
    if (mask & 4)  { if (p.y < 0.0f)         p.y = EPS;                       d.y = -d.y; index_global<OCT, false>(G, p, level, ind, rho); }
    if (mask & 8)  { if (p.y > (float)G.ny)  p.y = xsub((float)G.ny, EPS);    d.y = -d.y; index_global<OCT, false>(G, p, level, ind, rho); }
    if (mask & 16) { if (p.z < 0.0f)         p.z = EPS;                       d.z = -d.z; index_global<OCT, false>(G, p, level, ind, rho); }
    if (mask & 32) { if (p.z > (float)G.nz)  p.z = xsub((float)G.nz, EPS);    d.z = -d.z; index_global<OCT, false>(G, p, level, ind, rho); }
}

// ---- scattering ---------------------------------------------------------------------------------------------
// WITH_MSF: the dust species that scatters in cell `oind`, drawn with probability ABU*SCA / sum(ABU*SCA)
// (kernel_ASOC.c:777-794; the sum is OPT[2*oind+1]).  `u` is a uniform random number.
__device__ __forceinline__ int msf_pick(const float *__restrict__ abu, const float *__restrict__ scav, int ndust, float sum, int oind, float u) {
    float ds = xmul(0.99999f, u);
    int i;
    for (i = 0; i < ndust; i++) {
        ds = xsub(ds, xdiv(xmul(abu[i + (size_t)oind * ndust], scav[i]), sum));
        if (ds <= 0.0f) break;
    }
    return min(i, ndust - 1);
}

__device__ __forceinline__ void fix_direction(vec3 &d) {             // kernel_ASOC.c:508-511
    if (fabsf(d.x) < SOC_DEPS) d.x = SOC_DEPS;
    if (fabsf(d.y) < SOC_DEPS) d.y = SOC_DEPS;
    if (fabsf(d.z) < SOC_DEPS) d.z = SOC_DEPS;
    d = normalize3(d);
}
// Deflect (kernel_ASOC_aux.c:499-533): rotate d by polar angle acos(cos_theta) and azimuth phi
__device__ __forceinline__ void deflect(vec3 &d, float cos_theta, float phi) {
    float cx = d.x, cy = d.y, cz = d.z;
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    float sin_phi, cos_phi;
    sincosf(phi, &sin_phi, &cos_phi);
    float ox = sin_theta * cos_phi, oy = sin_theta * sin_phi, oz = cos_theta;
    float theta0 = acosf(cz / sqrtf(cx * cx + cy * cy + cz * cz + SOC_DEPS));
    float phi0 = acosf(cx / sqrtf(cx * cx + cy * cy + SOC_DEPS));
    if (d.y < 0.0f) phi0 = SOC_TWOPI - phi0;
    float st, ct, sp, cp;
    sincosf(-theta0, &st, &ct);
    sincosf(-phi0, &sp, &cp);
    d.x = +ox * ct * cp + oy * sp - oz * st * cp;
    d.y = -ox * ct * sp + oy * cp + oz * st * sp;
    d.z = +ox * st + oz * ct;
}
// Scatter (kernel_ASOC_aux.c:540-561)
template <class RNG>
__device__ __forceinline__ void scatter_direction(vec3 &d, const float *__restrict__ csc, int bins, RNG &rng) {
    float cos_theta = csc[clampi((int)floorf(rng.uniform() * bins), 0, bins - 1)];
    deflect(d, cos_theta, SOC_TWOPI * rng.uniform());
    fix_direction(d);
}

// New direction after a scattering: rotate d by the polar angle acos(ct) and a uniform azimuth phi -- the same
// distribution as Deflect() (kernel_ASOC_aux.c:499-533; the azimuth is uniform either way) without its acos /
// sincos chain -- then the reference's clamp |d_i| >= DEPS and renormalisation (kernel_ASOC.c:508-511).
// Approximate SFU functions (2 ulp) are enough here: the result is a random direction.
__device__ __forceinline__ void scatter_rotate(vec3 &d, float ct, float phi) {
    const float s2 = fmaxf(0.0f, 1.0f - ct * ct);
    const float st = s2 * rsqrtf(fmaxf(s2, 1.0e-30f));
    float sp, cp;
    __sincosf(phi, &sp, &cp);
    const float w2 = 1.0f - d.z * d.z;
    vec3 n;
    if (w2 > 1.0e-6f) {
        const float iw = rsqrtf(w2);
        n.x = st * (d.x * d.z * cp - d.y * sp) * iw + d.x * ct;
        n.y = st * (d.y * d.z * cp + d.x * sp) * iw + d.y * ct;
        n.z = -st * cp * w2 * iw + d.z * ct;
    } else {
        n.x = st * cp; n.y = st * sp; n.z = (d.z > 0.0f) ? ct : -ct;
    }
    if (fabsf(n.x) < SOC_DEPS) n.x = SOC_DEPS;
    if (fabsf(n.y) < SOC_DEPS) n.y = SOC_DEPS;
    if (fabsf(n.z) < SOC_DEPS) n.z = SOC_DEPS;
    const float il = rsqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
    d.x = n.x * il; d.y = n.y * il; d.z = n.z * il;
}
// ---- Healpix, RING scheme (kernel_ASOC_aux.c:945-1026; map flavour kernel_ASOC_map.c:59-140) ---------------
__device__ inline int ang2pix_ring(int nside, float phi, float theta) {
    int nl2, nl4, ncap, npix, jp, jm, ipix1, ir, ip, kshift;
    float z, za, tt, tp, tmp;
    if (theta < 0.0f || theta > SOC_PI) return -1;
    z = cosf(theta); za = fabsf(z);
    if (phi >= SOC_TWOPI) phi -= SOC_TWOPI;
    if (phi < 0.0f) phi += SOC_TWOPI;
    tt = phi / SOC_PIHALF;
    nl2 = 2 * nside; nl4 = 4 * nside; ncap = nl2 * (nside - 1); npix = 12 * nside * nside;
    if (za <= SOC_TWOTHIRD) {
        jp = (int)(nside * (0.5f + tt - z * 0.75f));
        jm = (int)(nside * (0.5f + tt + z * 0.75f));
        ir = nside + 1 + jp - jm;
        kshift = (ir % 2 == 0) ? 1 : 0;
        ip = (int)((jp + jm - nside + kshift + 1) / 2) + 1;
        if (ip > nl4) ip -= nl4;
        ipix1 = ncap + nl4 * (ir - 1) + ip;
    } else {
        tp = tt - (int)(tt);
        tmp = sqrtf(3.0f * (1.0f - za));
        jp = (int)(nside * tp * tmp);
        jm = (int)(nside * (1.0f - tp) * tmp);
        ir = jp + jm + 1;
        ip = (int)(tt * ir) + 1;
        if (ip > 4 * ir) ip -= 4 * ir;
        ipix1 = 2 * ir * (ir - 1) + ip;
        if (z <= 0.0f) ipix1 = npix - 2 * ir * (ir + 1) + ip;
    }
    return ipix1 - 1;
}
__device__ inline void pix2ang_ring(int nside, int ipix, float &phi, float &theta, const float pi_) {
    int nl2, nl4, npix, ncap, iring, iphi, ip, ipix1;
    float fact1, fact2, fodd, hip, fihip;
    npix = 12 * nside * nside; ipix1 = ipix + 1; nl2 = 2 * nside; nl4 = 4 * nside;
    ncap = 2 * nside * (nside - 1); fact1 = 1.5f * nside; fact2 = 3.0f * nside * nside;
    if (ipix1 <= ncap) {
        hip = ipix1 / 2.0f; fihip = (int)(hip);
        iring = (int)(sqrtf(hip - sqrtf(fihip))) + 1;
        iphi = ipix1 - 2 * iring * (iring - 1);
        theta = acosf(1.0f - iring * iring / fact2);
        phi = (iphi - 0.5f) * pi_ / (2.0f * iring);
    } else if (ipix1 <= nl2 * (5 * nside + 1)) {
        ip = ipix1 - ncap - 1;
        iring = (int)(ip / nl4) + nside;
        iphi = (ip % nl4) + 1;
        fodd = 0.5f * (1 + (iring + nside) % 2);
        theta = acosf((nl2 - iring) / fact1);
        phi = (iphi - fodd) * pi_ / (2.0f * nside);
    } else {
        ip = npix - ipix1 + 1;
        hip = ip / 2.0f; fihip = (int)(hip);
        iring = (int)(sqrtf(hip - sqrtf(fihip))) + 1;
        iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
        theta = acosf(-1.0f + iring * iring / fact2);
        phi = (iphi - 0.5f) * pi_ / (2.0f * iring);
    }
}

// ---- counters ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_add_counter(unsigned long long *dst, unsigned long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, v);
}
