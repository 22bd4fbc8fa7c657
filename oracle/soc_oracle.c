/* TEST INFRASTRUCTURE ONLY -- see soc_oracle.h for the contract.
 *
 * Sequential, plain-C restatement of the reference algorithm.  One "work item"
 * of the reference = one call of a static function here; the exported drivers
 * loop over work items with OpenMP.  Arithmetic follows the reference's order
 * of operations and single-precision intermediate types so that, run with the
 * same MWC64X streams and one thread, results agree with oracle/_ref to float
 * rounding (tests/test_oracle_vs_ref.py).
 */
#include "soc_oracle.h"
#include <math.h>
#include <string.h>
#include <omp.h>

/* ---- constants: kernel_ASOC_aux.c:5-9, 99-114 ------------------------------------------------ */
#define TWOPI    6.28318531f
#define TAULIM   5.0e-4f
#define PIHALF   1.5707963268f
#define TWOTHIRD 0.6666666667f
#define PI_F     3.1415926535897f
#define S_EPS    5.0e-4f
#define S_PEPS   1.0e-4f
#define S_DEPS   5.0e-5f
/* map kernel's own constants: kernel_ASOC_map.c:10-18 */
#define M_EPS    2.5e-4f
#define M_PEPS   5.0e-4f
#define M_PI_F   3.1415926536f
#define M_TWOPI  6.2831853072f

typedef struct { float x, y, z; } v3;

static inline float  fminf_(float a, float b) { return a < b ? a : b; }
static inline float  fmaxf_(float a, float b) { return a > b ? a : b; }
static inline float  clampf(float v, float lo, float hi) { return fminf_(fmaxf_(v, lo), hi); }
static inline int    clampi(int v, int lo, int hi) { int t = v > lo ? v : lo; return t < hi ? t : hi; }
static inline int32_t f2i(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
static inline v3     v3_norm(v3 a) { float l = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); v3 r = { a.x / l, a.y / l, a.z / l }; return r; }

/* =================================================================================================
 * MWC64X (mwc64x_rng.cl:12-49, skip_mwc.cl:11-76).  The reference computes the modular products
 * with shift-and-add loops; 128-bit integer arithmetic gives the same residues.
 * ================================================================================================= */
#define MWC_A      4294883355ULL
#define MWC_M      18446383549859758079ULL
#define MWC_BASEID 4077358422479273989ULL
typedef struct { uint32_t x, c; } rng_t;

static inline uint64_t mulmod(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) % MWC_M); }
static uint64_t powmod(uint64_t a, uint64_t e) {
    uint64_t sqr = a, acc = 1;
    while (e) { if (e & 1) acc = mulmod(acc, sqr); sqr = mulmod(sqr, sqr); e >>= 1; }
    return acc;
}
/* skip_mwc.cl:62-76 with vecSize=1, vecOffset=0, streamGap=2^38 (kernel_ASOC.c:74-77) */
static void rng_seed_base(rng_t *s, uint64_t base, uint64_t id) {
    uint64_t dist = base + id * 274877906944ULL;       /* wraps mod 2^64 exactly like the ulong arithmetic */
    uint64_t x = mulmod(MWC_BASEID, powmod(MWC_A, dist));
    s->x = (uint32_t)(x / MWC_A); s->c = (uint32_t)(x % MWC_A);
}
/* kernel_ASOC.c:77 : base offset from the float seed */
static uint64_t seed_to_base(float seed) { return (uint64_t)(fmodf(seed * 7.0f * PI_F, 1.0f) * 4294967296L); }
static void rng_seed(rng_t *s, float seed, uint64_t id) { rng_seed_base(s, seed_to_base(seed), id); }
/* mwc64x_rng.cl:17-28,44-49 */
static inline uint32_t rng_next(rng_t *s) {
    uint32_t res = s->x ^ s->c;
    uint64_t t = (uint64_t)MWC_A * s->x + s->c;
    s->x = (uint32_t)t; s->c = (uint32_t)(t >> 32);
    return res;
}
/* kernel_ASOC_aux.c:127 -- inclusive [0,1] */
static inline float rnd(rng_t *s) { return (float)rng_next(s) / 4294967295.0f; }

void orc_rng_stream(float seed, int64_t id, int n, uint32_t *out, uint32_t *state_xc) {
    rng_t r; rng_seed(&r, seed, (uint64_t)id); state_xc[0] = r.x; state_xc[1] = r.c;
    for (int i = 0; i < n; i++) out[i] = rng_next(&r);
}
void orc_rng_stream_base(uint64_t base, int64_t id, int n, uint32_t *out, uint32_t *state_xc) {
    rng_t r; rng_seed_base(&r, base, (uint64_t)id); state_xc[0] = r.x; state_xc[1] = r.c;
    for (int i = 0; i < n; i++) out[i] = rng_next(&r);
}
int orc_threads(void) { return omp_get_max_threads(); }
void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
/* work items handed to a host thread at a time (OpenMP dynamic schedule) */
void orc_set_chunk(int n) { omp_set_schedule(omp_sched_dynamic, n > 0 ? n : 64); }
__attribute__((constructor)) static void orc_init_schedule(void) { omp_set_schedule(omp_sched_dynamic, 64); }

/* =================================================================================================
 * Grid navigation.  `KIND` selects the simulation kernels' variant (kernel_ASOC_aux.c) or the map
 * kernel's private copies (kernel_ASOC_map.c), which differ in PEPS, in when positions switch to
 * double, and in the octet-containment test while climbing.
 * ================================================================================================= */
typedef struct {
    int nx, ny, nz, levels, nxyz;
    const int32_t *off; const float *dens; const int32_t *par;
    int dbl_sim;     /* kernel_ASOC_aux.c:25-37,46,207 : NX > DIMLIM (399 if LEVELS<3 else 100) */
    int dbl_map;     /* kernel_ASOC_map.c:302 : NX > 100 */
} nav_t;

static nav_t nav_make(const OrcParams *P, const OrcGrid *G) {
    nav_t n; n.nx = P->nx; n.ny = P->ny; n.nz = P->nz; n.levels = P->levels; n.nxyz = P->nx * P->ny * P->nz;
    n.off = G->off; n.dens = G->dens; n.par = G->par;
    n.dbl_sim = P->nx > (P->levels < 3 ? 399 : 100);
    n.dbl_map = P->nx > 100;
    return n;
}

/* kernel_ASOC_aux.c:131-165 (sim) */
static void index_g(const nav_t *N, v3 *p, int *level, int *ind) {
    *ind = -1;
    if (p->x <= 0.0f || p->y <= 0.0f || p->z <= 0.0f) return;
    if (p->x >= N->nx || p->y >= N->ny || p->z >= N->nz) return;
    *level = 0;
    *ind = (int)floorf(p->z) * N->nx * N->ny + (int)floorf(p->y) * N->nx + (int)floorf(p->x);
    if (N->dens[*ind] > 0.0f) return;
    for (;;) {
        p->x = 2.0f * fmodf(p->x, 1.0f); p->y = 2.0f * fmodf(p->y, 1.0f); p->z = 2.0f * fmodf(p->z, 1.0f);
        float link = -N->dens[N->off[*level] + *ind];
        *ind = f2i(link);
        (*level)++;
        *ind += 4 * (int)floorf(p->z) + 2 * (int)floorf(p->y) + (int)floorf(p->x);
        if (N->dens[N->off[*level] + *ind] > 0.0f) return;
    }
}

/* kernel_ASOC_map.c:187-220 (map): same walk, coordinate update written differently */
static void index_g_map(const nav_t *N, v3 *p, int *level, int *ind) {
    *ind = -1;
    if (p->x <= 0.0f || p->y <= 0.0f || p->z <= 0.0f) return;
    if (p->x >= N->nx || p->y >= N->ny || p->z >= N->nz) return;
    *level = 0;
    *ind = (int)floorf(p->z) * N->nx * N->ny + (int)floorf(p->y) * N->nx + (int)floorf(p->x);
    if (N->dens[*ind] > 0.0f) return;
    p->x = 2.0f * fmodf(p->x, 1.0f); p->y = 2.0f * fmodf(p->y, 1.0f); p->z = 2.0f * fmodf(p->z, 1.0f);
    for (;;) {
        float link = -N->dens[N->off[*level] + *ind];
        *ind = f2i(link);
        (*level)++;
        *ind += 4 * (int)floorf(p->z) + 2 * (int)floorf(p->y) + (int)floorf(p->x);
        if (N->dens[N->off[*level] + *ind] > 0.0f) return;
        p->x -= floorf(p->x); p->y -= floorf(p->y); p->z -= floorf(p->z);
        p->x *= 2.0f; p->y *= 2.0f; p->z *= 2.0f;
    }
}

/* Index(): kernel_ASOC_aux.c:198-278 (sim) / kernel_ASOC_map.c:294-379 (map).  Body instantiated for
 * float and double working precision. */
#define REAL float
#define FLOOR floorf
#define FMOD fmodf
#define INDEX_NAME index_walk_f
#include "soc_oracle_index.inc"
#undef REAL
#undef FLOOR
#undef FMOD
#undef INDEX_NAME
#define REAL double
#define FLOOR floor
#define FMOD fmod
#define INDEX_NAME index_walk_d
#include "soc_oracle_index.inc"
#undef REAL
#undef FLOOR
#undef FMOD
#undef INDEX_NAME

static inline void index_sim(const nav_t *N, v3 *p, int *level, int *ind) {
    if (N->dbl_sim) index_walk_d(N, p, level, ind, 0); else index_walk_f(N, p, level, ind, 0);
}
static inline void index_map(const nav_t *N, v3 *p, int *level, int *ind) {
    if (N->dbl_map) index_walk_d(N, p, level, ind, 1); else index_walk_f(N, p, level, ind, 1);
}
static inline void index_maph(const nav_t *N, v3 *p, int *level, int *ind) {     /* kernel_ASOC_map_H.c:216-291 as shipped */
    if (N->dbl_map) index_walk_d(N, p, level, ind, 2); else index_walk_f(N, p, level, ind, 2);
}

/* GetStep(): kernel_ASOC_aux.c:282-315 (float branch; the NX>9999 double branch is dead for int32 grids
 * we accept) and kernel_ASOC_map.c:387-429.  Only PEPS and the Index flavour differ. */
static inline float get_step_any(const nav_t *N, v3 *p, const v3 *d, int *level, int *ind, float peps, int map) {
    float dx = (d->x > 0.0f) ? ((1.0f + peps - fmodf(p->x, 1.0f)) / d->x) : ((-peps - fmodf(p->x, 1.0f)) / d->x);
    float dy = (d->y > 0.0f) ? ((1.0f + peps - fmodf(p->y, 1.0f)) / d->y) : ((-peps - fmodf(p->y, 1.0f)) / d->y);
    float dz = (d->z > 0.0f) ? ((1.0f + peps - fmodf(p->z, 1.0f)) / d->z) : ((-peps - fmodf(p->z, 1.0f)) / d->z);
    dx = fminf_(dx, fminf_(dy, dz));
    p->x += dx * d->x; p->y += dx * d->y; p->z += dx * d->z;
    dx = ldexpf(dx, -(*level));
    if (map == 2) index_maph(N, p, level, ind); else if (map) index_map(N, p, level, ind); else index_sim(N, p, level, ind);
    return dx;
}
static inline float get_step(const nav_t *N, v3 *p, const v3 *d, int *level, int *ind) {
    return get_step_any(N, p, d, level, ind, S_PEPS, 0);
}
static inline float get_step_map(const nav_t *N, v3 *p, const v3 *d, int *level, int *ind) {
    return get_step_any(N, p, d, level, ind, M_PEPS, 1);
}

/* RootPos(): kernel_ASOC_aux.c:169-190 */
static void root_pos(const nav_t *N, v3 *p, int level, int ind) {
    if (level == 0) return;
    while (level > 0) {
        ind = N->par[N->off[level] + ind - N->nxyz]; level--;
        p->x *= 0.5f; p->y *= 0.5f; p->z *= 0.5f;
        if (level == 0) {
            p->x += ind % N->nx; p->y += (ind / N->nx) % N->ny; p->z += ind / (N->nx * N->ny);
            return;
        } else {
            int sid = ind % 8;
            p->x += sid % 2; p->y += (sid / 2) % 2; p->z += sid / 4;
        }
    }
}

static inline void addf(float *p, float v);
/* InRoi(): kernel_ASOC_aux.c:1031-1048 -- root cell index if the cell lies inside ROI, else -1 */
static int in_roi(const nav_t *N, const int32_t *ROI, int level, int ind) {
    int i = ind, k = level, j;
    while (k > 0) { i = N->par[N->off[k] + i - N->nxyz]; k--; }
    k = i / (N->nx * N->ny);
    j = (i / N->nx) % N->ny;
    if ((i % N->nx) >= ROI[0] && (i % N->nx) <= ROI[1] && j >= ROI[2] && j <= ROI[3] && k >= ROI[4] && k <= ROI[5]) return i;
    return -1;
}
/* a packet has stepped into ROI: add it to ROI_SAVE[surface element, direction pixel] (kernel_ASOC.c:617-642, 1510-1535) */
static void roi_save_add(const OrcParams *P, const nav_t *N, float *roi_save, v3 pos, v3 dir, int level, int ind, float photons) {
    const int32_t *ROI = P->roi;
    const int RNX = (ROI[1] - ROI[0] + 1) * P->roi_step, RNY = (ROI[3] - ROI[2] + 1) * P->roi_step, RNZ = (ROI[5] - ROI[4] + 1) * P->roi_step;
    const float st = (float)P->roi_step;
    int ii = 0, jj;
    v3 R = pos;
    root_pos(N, &R, level, ind);
    if (R.x < (ROI[0] + 1.0e-3f) || R.x > (ROI[1] + 0.999f)) {
        ii = clampi((int)floorf((R.y - ROI[2]) * st), 0, RNY - 1); jj = clampi((int)floorf((R.z - ROI[4]) * st), 0, RNZ - 1);
        ii = ii + RNY * jj;
    }
    if (R.y < (ROI[2] + 1.0e-3f) || R.y > (ROI[3] + 0.999f)) {
        ii = clampi((int)floorf((R.x - ROI[0]) * st), 0, RNX - 1); jj = clampi((int)floorf((R.z - ROI[4]) * st), 0, RNZ - 1);
        ii = RNY * RNZ + ii + RNX * jj;
    }
    if (R.z < (ROI[4] + 1.0e-3f) || R.z > (ROI[5] + 0.999f)) {
        ii = clampi((int)floorf((R.x - ROI[0]) * st), 0, RNX - 1); jj = clampi((int)floorf((R.y - ROI[2]) * st), 0, RNY - 1);
        ii = RNY * RNZ + RNX * RNZ + ii + RNX * jj;
    }
    float theta = acosf(dir.z), phi = atan2f(dir.y, dir.x);
    jj = orc_ang2pix_ring(P->roi_nside, phi, theta);
    ii = clampi(ii, 0, RNX * RNY + RNY * RNZ + RNZ * RNX - 1);
    jj = clampi(jj, 0, 12 * P->roi_nside * P->roi_nside - 1);
    addf(&roi_save[(long)ii * 12 * P->roi_nside * P->roi_nside + jj], photons);
}

/* Surface(): kernel_ASOC_aux.c:912-940 */
static void surface(const nav_t *N, v3 *p, const v3 *d) {
    float dx, dy, dz;
    if (d->x > 0.0f) dx = (p->x < 0.0f) ? (S_PEPS - p->x) / d->x : -1.0e10f;
    else             dx = (p->x > N->nx) ? (N->nx - S_PEPS - p->x) / d->x : -1.0e10f;
    if (d->y > 0.0f) dy = (p->y < 0.0f) ? (S_PEPS - p->y) / d->y : -1.0e10f;
    else             dy = (p->y > N->ny) ? (N->ny - S_PEPS - p->y) / d->y : -1.0e10f;
    if (d->z > 0.0f) dz = (p->z < 0.0f) ? (S_PEPS - p->z) / d->z : -1.0e10f;
    else             dz = (p->z > N->nz) ? (N->nz - S_PEPS - p->z) / d->z : -1.0e10f;
    dx = fmaxf_(dx, fmaxf_(dy, dz));
    p->x += dx * d->x; p->y += dx * d->y; p->z += dx * d->z;
}

/* Mirror(): kernel_ASOC_aux.c:1054-1083, restated literally: `if (c) a ; b ;` -- the direction component of every
 * enabled border is negated whether or not that border was crossed (a second enabled border of the same axis undoes
 * the reflection). */
static void mirror(const nav_t *N, int mask, int exact, v3 *p, v3 *d, int *level, int *ind) {
    if (exact) {          /* geometrically intended behaviour: reflect at the border that was crossed, nothing else */
        if ((mask & 1)  && p->x < 0.0f)  { p->x = S_EPS;         d->x = -d->x; }
        if ((mask & 2)  && p->x > N->nx) { p->x = N->nx - S_EPS; d->x = -d->x; }
        if ((mask & 4)  && p->y < 0.0f)  { p->y = S_EPS;         d->y = -d->y; }
        if ((mask & 8)  && p->y > N->ny) { p->y = N->ny - S_EPS; d->y = -d->y; }
        if ((mask & 16) && p->z < 0.0f)  { p->z = S_EPS;         d->z = -d->z; }
        if ((mask & 32) && p->z > N->nz) { p->z = N->nz - S_EPS; d->z = -d->z; }
        index_g(N, p, level, ind);
        return;
    }
    if (mask & 1)  { if (p->x < 0.0f)  p->x = S_EPS;          d->x = -d->x; index_g(N, p, level, ind); }
    if (mask & 2)  { if (p->x > N->nx) p->x = N->nx - S_EPS;  d->x = -d->x; index_g(N, p, level, ind); }
    if (mask & 4)  { if (p->y < 0.0f)  p->y = S_EPS;          d->y = -d->y; index_g(N, p, level, ind); }
    if (mask & 8)  { if (p->y > N->ny) p->y = N->ny - S_EPS;  d->y = -d->y; index_g(N, p, level, ind); }
    if (mask & 16) { if (p->z < 0.0f)  p->z = S_EPS;          d->z = -d->z; index_g(N, p, level, ind); }
    if (mask & 32) { if (p->z > N->nz) p->z = N->nz - S_EPS;  d->z = -d->z; index_g(N, p, level, ind); }
}

/* Deflect(): kernel_ASOC_aux.c:499-533 */
static void deflect(v3 *d, float cos_theta, float phi) {
    float cx = d->x, cy = d->y, cz = d->z;
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    float sin_phi = sinf(phi), cos_phi = cosf(phi);
    float ox = sin_theta * cos_phi, oy = sin_theta * sin_phi, oz = cos_theta;
    float theta0 = acosf(cz / sqrtf(cx * cx + cy * cy + cz * cz + S_DEPS));
    float phi0 = acosf(cx / sqrtf(cx * cx + cy * cy + S_DEPS));
    if (d->y < 0.0f) phi0 = TWOPI - phi0;
    theta0 = -theta0; phi0 = -phi0;
    float st = sinf(theta0), ct = cosf(theta0), sp = sinf(phi0), cp = cosf(phi0);
    d->x = +ox * ct * cp + oy * sp - oz * st * cp;
    d->y = -ox * ct * sp + oy * cp + oz * st * sp;
    d->z = +ox * st + oz * ct;
}
static inline void dir_fix(v3 *d) {       /* kernel_ASOC.c:508-511 */
    if (fabsf(d->x) < S_DEPS) d->x = S_DEPS;
    if (fabsf(d->y) < S_DEPS) d->y = S_DEPS;
    if (fabsf(d->z) < S_DEPS) d->z = S_DEPS;
    *d = v3_norm(*d);
}
/* Scatter(): kernel_ASOC_aux.c:540-561 (HG_TEST==0 branch) */
static void scatter(v3 *d, const float *csc, int bins, rng_t *r) {
    float cos_theta = csc[clampi((int)floorf(rnd(r) * bins), 0, bins - 1)];
    deflect(d, cos_theta, TWOPI * rnd(r));
    dir_fix(d);
}

/* Healpix RING: kernel_ASOC_aux.c:945-1026 */
int orc_ang2pix_ring(int nside, float phi, float theta) {
    int nl2, nl4, ncap, npix, jp, jm, ipix1, ir, ip, kshift;
    float z, za, tt, tp, tmp;
    if (theta < 0.0f || theta > PI_F) return -1;
    z = cosf(theta); za = fabsf(z);
    if (phi >= TWOPI) phi -= TWOPI;
    if (phi < 0.0f) phi += TWOPI;
    tt = phi / PIHALF;
    nl2 = 2 * nside; nl4 = 4 * nside; ncap = nl2 * (nside - 1); npix = 12 * nside * nside;
    if (za <= TWOTHIRD) {
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
void orc_pix2ang_ring(int nside, int ipix, float *phi, float *theta) {
    int nl2, nl4, npix, ncap, iring, iphi, ip, ipix1;
    float fact1, fact2, fodd, hip, fihip;
    npix = 12 * nside * nside; ipix1 = ipix + 1; nl2 = 2 * nside; nl4 = 4 * nside;
    ncap = 2 * nside * (nside - 1); fact1 = 1.5f * nside; fact2 = 3.0f * nside * nside;
    if (ipix1 <= ncap) {
        hip = ipix1 / 2.0f; fihip = (int)(hip);
        iring = (int)(sqrtf(hip - sqrtf(fihip))) + 1;
        iphi = ipix1 - 2 * iring * (iring - 1);
        *theta = acosf(1.0f - iring * iring / fact2);
        *phi = (iphi - 0.5f) * PI_F / (2.0f * iring);
    } else if (ipix1 <= nl2 * (5 * nside + 1)) {
        ip = ipix1 - ncap - 1;
        iring = (int)(ip / nl4) + nside;
        iphi = (ip % nl4) + 1;
        fodd = 0.5f * (1 + (iring + nside) % 2);
        *theta = acosf((nl2 - iring) / fact1);
        *phi = (iphi - fodd) * PI_F / (2.0f * nside);
    } else {
        ip = npix - ipix1 + 1;
        hip = ip / 2.0f; fihip = (int)(hip);
        iring = (int)(sqrtf(hip - sqrtf(fihip))) + 1;
        iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
        *theta = acosf(-1.0f + iring * iring / fact2);
        *phi = (iphi - 0.5f) * PI_F / (2.0f * iring);
    }
}

/* Parents: kernel_ASOC_aux.c:688-718 */
void orc_parents(const OrcParams *P, const OrcGrid *G) {
    int nxyz = P->nx * P->ny * P->nz;
    for (int level = 0; level < P->levels - 1; level++) {
        for (int ipar = 0; ipar < G->lcells[level]; ipar++) {
            float link = G->dens[G->off[level] + ipar];
            if (link < 1.0e-10f) {
                int ind = f2i(-link);
                for (int i = 0; i < 8; i++) G->par[G->off[level + 1] - nxyz + ind + i] = ipar;
            }
        }
    }
}

/* =================================================================================================
 * Absorption bookkeeping shared by SimRAM_PB / HP / CL (kernel_ASOC.c:586-612, 711-733)
 * ================================================================================================= */
static inline void addf(float *p, float v) {
    #pragma omp atomic
    *p += v;
}
typedef struct {
    const OrcParams *P; const OrcSimBufs *B; nav_t N; float tw; int use_int; OrcCounters c;
} sim_t;

static inline void cell_opacity(const sim_t *S, int oind, float *kabs, float *ksca) {
    if (S->P->with_abu) { *kabs = S->B->opt[2 * (long)oind]; *ksca = S->B->opt[2 * (long)oind + 1]; }
    else                { *kabs = S->B->abs; *ksca = S->B->sca; }
}
static inline void deposit(sim_t *S, int oind, float delta, const v3 *dir, int e_index) {
    const OrcSimBufs *B = S->B;
    if (S->P->with_ali == 1 && oind == e_index) addf(&B->xab[oind], delta * S->tw);          /* kernel_ASOC.c:1486-1494 */
    else                                        addf(&B->tabs[oind], delta * S->tw * S->P->adhoc);
    if (S->use_int) addf(&B->intens[oind], delta);
    if (S->P->save_intensity == 2) {
        addf(&B->intx[oind], delta * dir->x); addf(&B->inty[oind], delta * dir->y); addf(&B->intz[oind], delta * dir->z);
    }
    S->c.steps++;
}
/* free-path sampling incl. optional step weighting: kernel_ASOC.c:516-541 */
static inline float sample_free_path(const OrcParams *P, rng_t *r, float *photons) {
    float fp;
    if (P->step_weight <= 0) { fp = -logf(rnd(r)); }
    else if (P->step_weight == 1) {
        fp = -logf(rnd(r)) / P->sw_a;
        *photons *= expf(P->sw_a * fp - fp) / P->sw_a;
    } else {
        float a = P->sw_a, b = P->sw_b;
        fp = -logf((-b + sqrtf(b * b + 4.0f * rnd(r) * (1.0f - b))) / (2.0f - 2.0f * b)) / a;
        *photons *= 1.0f / (a * b * expf((1.0f - a) * fp) + 2.0f * a * (1.0f - b) * expf((1.0f - 2.0f * a) * fp));
    }
    return fp;
}

/* The propagation loop common to the three emission kernels.
 * kind 0: SimRAM_PB / SimRAM_HP (kernel_ASOC.c:556-820, 985-1202) -- failed-step nudge, scattering limit
 *         tested after the scatter.
 * kind 1: SimRAM_CL (kernel_ASOC.c:1448-1679) -- no nudge, limit tested before the partial-step absorption. */
static void propagate(sim_t *S, rng_t *r, v3 pos, v3 dir, int level, int ind, float photons, float free_path,
                      int kind, int e_index) {
    const nav_t *N = &S->N;
    int scatterings = 0, oind = 0, ind0 = -1, level0 = 0;
    float tau, dtau, tauA, ds, dx, delta, kabs, ksca;
    v3 pos0 = pos;
    const int rsave = S->P->with_roi_save > 0;
    int roi = -1, oroi = -1;
    if (rsave && ind >= 0) roi = oroi = in_roi(N, S->P->roi, level, ind);       /* kernel_ASOC.c:550, 1439 */
    while (ind >= 0) {
        tau = 0.0f;
        while (ind >= 0) {
            oroi = roi;
            oind = N->off[level] + ind; ind0 = ind; level0 = level; pos0 = pos;
            ds = get_step(N, &pos, &dir, &level, &ind);
            cell_opacity(S, oind, &kabs, &ksca);
            tauA = ds * N->dens[oind] * kabs;
            dtau = ds * N->dens[oind] * ksca;
            if (free_path < (tau + dtau)) { ind = ind0; break; }
            delta = (tauA > TAULIM) ? (photons * (1.0f - expf(-tauA))) : (photons * tauA * (1.0f - 0.5f * tauA));
            deposit(S, oind, delta, &dir, e_index);
            photons *= expf(-tauA);
            tau += dtau;
            if (rsave) {                                                   /* only at the end of a full step, :615-643 */
                roi = (ind >= 0 || level == 0) ? in_roi(N, S->P->roi, level, ind) : -1;
                if (roi >= 0 && oroi < 0) roi_save_add(S->P, N, S->B->roi_save, pos, dir, level, ind, photons);
            }
            if (kind == 0 && level == level0 && ind == ind0) {             /* kernel_ASOC.c:649-665 */
                pos.x += S_PEPS * dir.x; pos.y += S_PEPS * dir.y; pos.z += S_PEPS * dir.z;
            }
            if (S->P->mirror > 0 && ind < 0) mirror(N, S->P->mirror, S->P->mirror_exact, &pos, &dir, &level, &ind);   /* :686-688, 1540-1542 */
        }
        if (ind < 0) break;
        scatterings++;
        if (kind == 1 && scatterings > 20) break;                           /* kernel_ASOC.c:1552-1556 */
        dtau = free_path - tau;
        cell_opacity(S, oind, &kabs, &ksca);
        dx = dtau / (ksca * N->dens[oind]);
        tauA = dx * N->dens[oind] * kabs;
        delta = (tauA > TAULIM) ? (photons * (1.0f - expf(-tauA))) : (photons * tauA * (1.0f - 0.5f * tauA));
        deposit(S, oind, delta, &dir, e_index);
        dx = ldexpf(dx, level0);
        dx = fmaxf_(0.0f, dx - 2.0f * S_PEPS);
        pos.x = pos0.x + dx * dir.x; pos.y = pos0.y + dx * dir.y; pos.z = pos0.z + dx * dir.z;
        photons *= expf(-tauA);
        free_path = sample_free_path(S->P, r, &photons);
        ind = ind0; level = level0;
        if (S->P->with_msf > 0) {                                           /* kernel_ASOC.c:777-794 */
            float sum = S->B->opt[2 * (long)oind + 1];
            float u = 0.99999f * rnd(r);
            int idust;
            for (idust = 0; idust < S->P->ndust; idust++) {
                u -= S->B->abu[idust + (long)oind * S->P->ndust] * S->B->sca_v[idust] / sum;
                if (u <= 0.0) break;
            }
            if (idust >= S->P->ndust) idust = S->P->ndust - 1;
            scatter(&dir, S->B->csc + idust * S->P->bins, S->P->bins, r);
        } else scatter(&dir, S->B->csc, S->P->bins, r);
        S->c.scatterings++;
        if (kind == 0 && scatterings > 20) break;                           /* kernel_ASOC.c:801-804 */
    }
}

/* ROI background (SOURCE==3, WITH_ROI_LOAD): 100 work items per surface element of the loaded file, `packets` =
 * number of surface elements, batch = multiple of the Healpix pixel count (kernel_ASOC.c:97-179, 469-502; the same
 * code in kernel_ASOC_sca.c:530-612, 812-845) */
typedef struct { int ielem, iside; float rd, dx, dy, x0; } roi_src_t;
static void roi_src_init(const OrcParams *P, int id, int packets, int batch, roi_src_t *r) {
    const int32_t *RD = P->roi_dim;
    int iside;
    r->ielem = id % packets; iside = r->ielem;
    r->rd = P->nx / ((float)RD[0]);
    r->dx = r->dy = 0.0f;
    if (iside < RD[1] * RD[2]) { r->dx = ((iside % RD[1]) + 0.5f) * r->rd; r->dy = ((iside / RD[1]) + 0.5f) * r->rd; iside = 0; }
    else { iside -= RD[1] * RD[2];
    if (iside < RD[0] * RD[2]) { r->dx = ((iside % RD[0]) + 0.5f) * r->rd; r->dy = ((iside / RD[0]) + 0.5f) * r->rd; iside = 1; }
    else { iside -= RD[0] * RD[2];
    if (iside < RD[0] * RD[1]) { r->dx = ((iside % RD[0]) + 0.5f) * r->rd; r->dy = ((iside / RD[0]) + 0.5f) * r->rd; iside = 2; } } }
    r->iside = iside;
    r->x0 = (float)(P->roi_nside * P->roi_nside * 12.0 / (100.0 * batch));
}
/* returns 0 when the direction pixel is empty (no packet, no random numbers) */
static int roi_src_emit(const OrcParams *P, const OrcSimBufs *B, const nav_t *N, const roi_src_t *r, rng_t *rng, int III,
                        v3 *pos, v3 *dir, int *level, int *ind, float *photons) {
    const int npix = 12 * P->roi_nside * P->roi_nside, NX = P->nx, NY = P->ny, NZ = P->nz;
    const float rd = r->rd;
    *ind = III % npix;
    *photons = r->x0 * B->roi_load[(long)r->ielem * npix + *ind];
    if (*photons <= 0.0f) return 0;
    float v1, v2;
    orc_pix2ang_ring(P->roi_nside, *ind, &v1, &v2);
    v1 += (rnd(rng) - 0.5f) * 0.05f;
    v2 += (rnd(rng) - 0.5f) * 0.05f;
    dir->x = sinf(v2) * cosf(v1); dir->y = sinf(v2) * sinf(v1); dir->z = cosf(v2);
    if (r->iside == 0) { pos->y = r->dx + (-0.49f + 0.98f * rnd(rng)) * rd; pos->z = r->dy + (-0.49f + 0.98f * rnd(rng)) * rd; pos->x = (dir->x > 0.0f) ? S_PEPS : (NX - S_PEPS); }
    if (r->iside == 1) { pos->x = r->dx + (-0.49f + 0.98f * rnd(rng)) * rd; pos->z = r->dy + (-0.49f + 0.98f * rnd(rng)) * rd; pos->y = (dir->y > 0.0f) ? S_PEPS : (NY - S_PEPS); }
    if (r->iside == 2) { pos->x = r->dx + (-0.49f + 0.98f * rnd(rng)) * rd; pos->y = r->dy + (-0.49f + 0.98f * rnd(rng)) * rd; pos->z = (dir->z > 0.0f) ? S_PEPS : (NZ - S_PEPS); }
    index_g(N, pos, level, ind);
    return 1;
}

/* ---- SimRAM_PB: one work item (kernel_ASOC.c:15-824) ------------------------------------------ */
static void pb_item(sim_t *S, int id, int source, int packets, int batch, float seed, float bg) {
    const OrcParams *P = S->P; const OrcSimBufs *B = S->B; const nav_t *N = &S->N;
    const int NX = P->nx, NY = P->ny, NZ = P->nz;
    const int AREA = 2 * (NX * NY + NY * NZ + NZ * NX);
    rng_t rng; rng_seed(&rng, seed, (uint64_t)id);
    int ind = -1, level = 0, side = 0;
    float X0 = 0, Y0 = 0, Z0 = 0, DX = 1, DY = 1, DZ = 1, photons = 0.0f;
    v3 pos = { 0, 0, 0 }, dir = { 0, 0, 0 };
    if (source == 1 && id >= 8 * AREA) return;
    if (source == 3 && !(P->with_roi_load > 0)) return;
    roi_src_t rs;
    if (source == 3) { if (id >= 100 * packets) return; roi_src_init(P, id, packets, batch, &rs); }
    if (source == 1) {                                                       /* kernel_ASOC.c:109-138 */
        ind = id % AREA;
        if (ind < NY * NZ) { side = 0; X0 = S_PEPS; Y0 = ind % NY; Z0 = ind / NY; DX = 0.0f; }
        else { ind -= NY * NZ;
        if (ind < NY * NZ) { side = 1; X0 = NX - S_PEPS; Y0 = ind % NY; Z0 = ind / NY; DX = 0.0f; }
        else { ind -= NY * NZ;
        if (ind < NX * NZ) { side = 2; Y0 = S_PEPS; X0 = ind % NX; Z0 = ind / NX; DY = 0.0f; }
        else { ind -= NX * NZ;
        if (ind < NX * NZ) { side = 3; Y0 = NY - S_PEPS; X0 = ind % NX; Z0 = ind / NX; DY = 0.0f; }
        else { ind -= NX * NZ;
        if (ind < NX * NY) { side = 4; Z0 = S_PEPS; X0 = ind % NX; Y0 = ind / NX; DZ = 0.0f; }
        else { ind -= NX * NY; side = 5; Z0 = NZ - S_PEPS; X0 = ind % NX; Y0 = ind / NX; DZ = 0.0f; } } } } }
    }
    for (int III = 0; III < batch; III++) {
        if (source == 0) {                                                   /* kernel_ASOC.c:202-434 */
            float phi = TWOPI * rnd(&rng);
            float cos_theta = 0.999997f - 1.999995f * rnd(&rng);
            float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
            dir.x = sin_theta * cosf(phi); dir.y = sin_theta * sinf(phi); dir.z = cos_theta;
            int ips = III % P->no_ps;
            photons = B->ps[ips];
            v3 src = { B->pspos[3 * ips], B->pspos[3 * ips + 1], B->pspos[3 * ips + 2] };
            pos = src;
            index_g(N, &pos, &level, &ind);
            if (ind < 0 || ind >= P->cells) {                                 /* external source */
                float v1, v2;
                switch (P->ps_method) {
                case 0:
                    surface(N, &pos, &dir); index_g(N, &pos, &level, &ind); break;
                case 1:
                    pos = src;
                    if (pos.z > NZ) { if (dir.z > 0.0f) dir.z = -dir.z; }
                    else if (pos.z < 0.0f) { if (dir.z < 0.0f) dir.z = -dir.z; }
                    else if (pos.x > NX) { if (dir.x > 0.0f) dir.x = -dir.x; }
                    else if (pos.x < 0.0f) { if (dir.x < 0.0f) dir.x = -dir.x; }
                    else if (pos.y > NY) { if (dir.y > 0.0f) dir.y = -dir.y; }
                    else if (pos.y < 0.0f) { if (dir.y < 0.0f) dir.y = -dir.y; }
                    surface(N, &pos, &dir);
                    photons *= 0.5f;
                    index_g(N, &pos, &level, &ind); break;
                case 2: {
                    pos = src;
                    ind = (int)floorf(rnd(&rng) * B->xps_nside[ips] * 0.999999f);
                    photons /= B->xps_area[3 * ips + ind];
                    ind = B->xps_side[3 * ips + ind];
                    float a = rnd(&rng), b = rnd(&rng);
                    if (ind == 0) { pos.x = NX - S_PEPS; pos.y = a * NY; pos.z = b * NZ; b = NY * NZ; }
                    if (ind == 1) { pos.x = S_PEPS;      pos.y = a * NY; pos.z = b * NZ; b = NY * NZ; }
                    if (ind == 2) { pos.y = NY - S_PEPS; pos.x = a * NX; pos.z = b * NZ; b = NX * NZ; }
                    if (ind == 3) { pos.y = S_PEPS;      pos.x = a * NX; pos.z = b * NZ; b = NX * NZ; }
                    if (ind == 4) { pos.z = NZ - S_PEPS; pos.x = a * NX; pos.y = b * NY; b = NX * NY; }
                    if (ind == 5) { pos.z = S_PEPS;      pos.x = a * NX; pos.y = b * NY; b = NX * NY; }
                    v3 dd = { pos.x - src.x, pos.y - src.y, pos.z - src.z };
                    v1 = sqrtf(dd.x * dd.x + dd.y * dd.y + dd.z * dd.z);
                    dir = v3_norm(dd);
                    v2 = (ind < 2) ? fabsf(dir.x) : ((ind < 4) ? fabsf(dir.y) : fabsf(dir.z));
                    photons *= v2 * b / (4.0f * PI_F * v1 * v1);
                    index_g(N, &pos, &level, &ind); break; }
                case 4: {
                    v1 = src.z - NZ;
                    cos_theta = v1 / sqrtf(v1 * v1 + 0.25f * NX * NX + 0.25f * NY * NY);
                    photons *= 0.5f * (1.0f - cos_theta);
                    cos_theta = 1.0f - rnd(&rng) * (1.0f - cos_theta);
                    v1 = TWOPI * rnd(&rng);
                    dir.x = sqrtf(1.0f - cos_theta * cos_theta) * cosf(v1);
                    dir.y = sqrtf(1.0f - cos_theta * cos_theta) * sinf(v1);
                    dir.z = -cos_theta;
                    surface(N, &pos, &dir); index_g(N, &pos, &level, &ind); break; }
                case 5: {
                    cos_theta = B->xps_area[3 * ips];
                    photons *= 0.5f * (1.0f - cos_theta);
                    cos_theta = 1.0f - rnd(&rng) * (1.0f - cos_theta);
                    v1 = TWOPI * rnd(&rng);
                    int s = B->xps_side[3 * ips];
                    float st = sqrtf(1.0f - cos_theta * cos_theta);
                    if (s < 2)      { dir.y = st * cosf(v1); dir.z = st * sinf(v1); dir.x = (s == 0) ? -cos_theta : cos_theta; }
                    else if (s < 4) { dir.x = st * cosf(v1); dir.z = st * sinf(v1); dir.y = (s == 2) ? -cos_theta : cos_theta; }
                    else            { dir.x = st * cosf(v1); dir.y = st * sinf(v1); dir.z = (s == 4) ? -cos_theta : cos_theta; }
                    surface(N, &pos, &dir); index_g(N, &pos, &level, &ind); break; }
                default: break;
                }
            }
        }
        if (source == 1) {                                                   /* kernel_ASOC.c:439-464 */
            pos.x = clampf(X0 + DX * rnd(&rng), S_PEPS, NX - S_PEPS);
            pos.y = clampf(Y0 + DY * rnd(&rng), S_PEPS, NY - S_PEPS);
            pos.z = clampf(Z0 + DZ * rnd(&rng), S_PEPS, NZ - S_PEPS);
            float cos_theta = sqrtf(rnd(&rng));
            float phi = TWOPI * rnd(&rng);
            float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
            float v1 = sin_theta * cosf(phi), v2 = sin_theta * sinf(phi);
            switch (side) {
            case 0: dir.x =  cos_theta; dir.y = v1; dir.z = v2; break;
            case 1: dir.x = -cos_theta; dir.y = v1; dir.z = v2; break;
            case 2: dir.y =  cos_theta; dir.x = v1; dir.z = v2; break;
            case 3: dir.y = -cos_theta; dir.x = v1; dir.z = v2; break;
            case 4: dir.z =  cos_theta; dir.x = v1; dir.y = v2; break;
            case 5: dir.z = -cos_theta; dir.x = v1; dir.y = v2; break;
            }
            photons = bg;
            index_g(N, &pos, &level, &ind);
        }
        if (source == 3 && !roi_src_emit(P, B, N, &rs, &rng, III, &pos, &dir, &level, &ind, &photons)) continue;   /* :469-502 */
        dir_fix(&dir);
        float free_path = sample_free_path(P, &rng, &photons);
        S->c.packets++;
        propagate(S, &rng, pos, dir, level, ind, photons, free_path, 0, -1);
        /* NOTE: `ind`, `level`, `dir` deliberately keep their last emission values between packets, like the
           reference's function-scope variables do (an external source that misses the cloud re-uses them). */
    }
}

void orc_sim_pb(const OrcParams *P, const OrcGrid *G, const OrcSimBufs *B, int global, int source, int packets,
                int batch, float seed, float bg, float tw, OrcCounters *C) {
    (void)packets;
    uint64_t np = 0, ns = 0, nsc = 0;
    #pragma omp parallel for schedule(runtime) reduction(+:np,ns,nsc)
    for (int id = 0; id < global; id++) {
        sim_t S; S.P = P; S.B = B; S.N = nav_make(P, G); S.tw = tw;
        S.use_int = (P->save_intensity == 1 || P->save_intensity == 2 || P->noabsorbed == 0);
        memset(&S.c, 0, sizeof(S.c));
        pb_item(&S, id, source, packets, batch, seed, bg);
        np += S.c.packets; ns += S.c.steps; nsc += S.c.scatterings;
    }
    if (C) { C->packets += np; C->steps += ns; C->scatterings += nsc; }
}

/* ---- SimRAM_HP: one work item (kernel_ASOC.c:831-1206) ----------------------------------------- */
static void hp_item(sim_t *S, int id, int batch, float seed) {
    const OrcParams *P = S->P; const OrcSimBufs *B = S->B; const nav_t *N = &S->N;
    const int NX = P->nx, NY = P->ny, NZ = P->nz;
    const int AREA = 2 * (NX * NY + NY * NZ + NZ * NX);
    rng_t rng; rng_seed(&rng, seed, (uint64_t)id);
    if (id >= 8 * AREA) return;
    int ind = -1, level = 0, ind0, level0;
    float photons, phi, theta, x, y, z, ds, v1, v2;
    v3 pos, dir;
    for (int III = 0; III < batch; III++) {
        if (P->hpbg_weighted < 1) {
            ind = clampi((int)(floorf(rnd(&rng) * 49152)), 0, 49151);
        } else {                                                             /* kernel_ASOC.c:890-907 */
            x = rnd(&rng); ind0 = 0; level0 = 49151;
            for (int i = 0; i < 10; i++) {
                ind = (ind0 + level0) / 2;
                if (B->hpbgp[ind] > x) level0 = ind; else ind0 = ind;
            }
            for (ind = ind0; ind <= level0; ind++) if (B->hpbgp[ind] >= x) break;
        }
        photons = B->hpbg[ind];
        orc_pix2ang_ring(64, ind, &phi, &theta);
        dir.x = +sinf(theta) * cosf(phi); dir.y = +sinf(theta) * sinf(phi); dir.z = -cosf(theta);
        dir_fix(&dir);
        x = fabsf(dir.x); y = fabsf(dir.y); z = fabsf(dir.z);
        ds = x + y + z; x /= ds; y /= ds; z /= ds;
        ds = rnd(&rng); v1 = rnd(&rng); v2 = rnd(&rng);
        if (ds < x)            { pos.y = v1 * NY; pos.z = v2 * NZ; pos.x = (dir.x > 0.0f) ? S_PEPS : (NX - S_PEPS); }
        else if (ds < (x + y)) { pos.x = v1 * NX; pos.z = v2 * NZ; pos.y = (dir.y > 0.0f) ? S_PEPS : (NY - S_PEPS); }
        else                   { pos.x = v1 * NX; pos.y = v2 * NY; pos.z = (dir.z > 0.0f) ? S_PEPS : (NZ - S_PEPS); }
        index_g(N, &pos, &level, &ind);
        float free_path = sample_free_path(P, &rng, &photons);
        S->c.packets++;
        propagate(S, &rng, pos, dir, level, ind, photons, free_path, 0, -1);
    }
}
void orc_sim_hp(const OrcParams *P, const OrcGrid *G, const OrcSimBufs *B, int global, int packets, int batch,
                float seed, float tw, OrcCounters *C) {
    (void)packets;
    uint64_t np = 0, ns = 0, nsc = 0;
    #pragma omp parallel for schedule(runtime) reduction(+:np,ns,nsc)
    for (int id = 0; id < global; id++) {
        sim_t S; S.P = P; S.B = B; S.N = nav_make(P, G); S.tw = tw;
        S.use_int = (P->save_intensity == 1 || P->save_intensity == 2 || P->noabsorbed == 0);
        memset(&S.c, 0, sizeof(S.c));
        hp_item(&S, id, batch, seed);
        np += S.c.packets; ns += S.c.steps; nsc += S.c.scatterings;
    }
    if (C) { C->packets += np; C->steps += ns; C->scatterings += nsc; }
}

/* ---- SimRAM_CL: one work item (kernel_ASOC.c:1223-1684, USE_EMWEIGHT 0/1) ---------------------- */
static void cl_item(sim_t *S, int id, int global, int batch_arg, float seed) {
    const OrcParams *P = S->P; const OrcSimBufs *B = S->B; const nav_t *N = &S->N;
    if (id >= P->cells) return;
    rng_t rng; rng_seed(&rng, seed, (uint64_t)id);
    int icell = id - global, iray = 0, batch = -1, ind, level;
    float pwei = 1.0f, X0, Y0, Z0;
    for (;;) {
        if (iray >= batch) {
            iray = 0; pwei = 1.0f;
            for (;;) {
                icell += global;
                if (icell >= P->cells) return;
                if (P->use_emweight > 0) {
                    pwei = B->emwei[icell];
                    if (pwei < 1e-10f || N->dens[icell] <= 0.0f) continue;
                    batch = (int)floorf(pwei);
                    if (batch < 1) { batch = 1; pwei = (float)(1.0 / (pwei + 1.0e-30f)); }
                    else           { pwei = (float)(1.0 / (batch + 1.0e-9f)); }
                } else {
                    batch = batch_arg;
                    pwei = 1.0f / (batch + 1.0e-9f);
                }
                break;
            }
        }
        ind = icell; iray++;
        for (level = 0; level < P->levels - 1; level++) {
            ind -= N->off[level + 1] - N->off[level];                       /* LCELLS[level] */
            if (ind < 0) { ind += N->off[level + 1] - N->off[level]; break; }
        }
        if (level == 0) { X0 = ind % P->nx; Y0 = (ind / P->nx) % P->ny; Z0 = ind / (P->nx * P->ny); }
        else { int sid = ind % 8; X0 = sid % 2; Y0 = ((sid % 4) > 1) ? 1.0f : 0.0f; Z0 = sid / 4; }
        float photons = B->emit[N->off[level] + ind] * pwei;
        v3 pos, dir;
        pos.x = X0 + rnd(&rng); pos.y = Y0 + rnd(&rng); pos.z = Z0 + rnd(&rng);
        float phi = TWOPI * rnd(&rng);
        float cos_theta = 0.999997f - 1.999995f * rnd(&rng);
        float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
        dir.x = sin_theta * cosf(phi); dir.y = sin_theta * sinf(phi); dir.z = cos_theta;
        int e_index = (P->with_ali > 0) ? N->off[level] + ind : -1;
        dir_fix(&dir);
        float free_path = sample_free_path(P, &rng, &photons);
        S->c.packets++;
        propagate(S, &rng, pos, dir, level, ind, photons, free_path, 1, e_index);
    }
}
void orc_sim_cl(const OrcParams *P, const OrcGrid *G, const OrcSimBufs *B, int global, int packets, int batch,
                float seed, float tw, OrcCounters *C) {
    (void)packets;
    uint64_t np = 0, ns = 0, nsc = 0;
    #pragma omp parallel for schedule(runtime) reduction(+:np,ns,nsc)
    for (int id = 0; id < global; id++) {
        sim_t S; S.P = P; S.B = B; S.N = nav_make(P, G); S.tw = tw;
        S.use_int = (P->save_intensity == 1 || P->save_intensity == 2 || P->noabsorbed == 0);
        memset(&S.c, 0, sizeof(S.c));
        cl_item(&S, id, global, batch, seed);
        np += S.c.packets; ns += S.c.steps; nsc += S.c.scatterings;
    }
    if (C) { C->packets += np; C->steps += ns; C->scatterings += nsc; }
}

/* =================================================================================================
 * Equilibrium temperature and emission (kernel_ASOC_aux.c:745-807)
 * ================================================================================================= */
void orc_eq_temperature(const OrcParams *P, const OrcGrid *G, int level, float adhoc, float kE, float Emin, int NE,
                        const float *ttt, const float *emit, float *tnew) {
    float scale = (6.62607e-27f * P->factor) / P->length;
    float oplgkE = 1.0f / log10f(kE);
    float beta = 1.0f;
    #pragma omp parallel for
    for (int i = 0; i < G->lcells[level]; i++) {
        int ind = G->off[level] + i;
        float Ein = (scale / adhoc) * emit[ind] * powf(8.0f, (float)level) / G->dens[ind];
        int iE = clampi((int)floorf(oplgkE * log10f((Ein / beta) / Emin)), 0, NE - 2);
        float wi = (Emin * powf(kE, (float)(iE + 1)) - (Ein / beta)) / (Emin * powf(kE, (float)iE) * (kE - 1.0f));
        tnew[ind] = (G->dens[ind] > 1.0e-7f) ? clampf(wi * ttt[iE] + (1.0f - wi) * ttt[iE + 1], 3.0f, 1600.0f) : 10.0f;
    }
}
void orc_emission(const OrcParams *P, const OrcGrid *G, float freq, float fabs_, const float *t, float *emit) {
    (void)G;
    #pragma omp parallel for
    for (int i = 0; i < P->cells; i++)
        emit[i] = (2.79639459e-20f * P->factor) * fabs_ * (freq * freq / (expf(4.7995074e-11f * freq / t[i]) - 1.0f)) / P->length;
}

/* Emission2: kernel_ASOC_aux.c:862-888 -- cells [c0,c1[, all nfreq frequencies, EMIT[(icell-c0)*nfreq+ifreq] */
void orc_emission2(const OrcParams *P, int c0, int c1, int nfreq, const float *freq, const float *fabs_, const float *t, float *emit) {
    #pragma omp parallel for
    for (int icell = c0; icell < c1; icell++)
        for (int ifreq = 0; ifreq < nfreq; ifreq++)
            emit[(long)(icell - c0) * nfreq + ifreq] =
                (2.79639459e-20f * P->factor) * fabs_[ifreq] * (freq[ifreq] * freq[ifreq] / (expf(4.7995074e-11f * freq[ifreq] / t[icell]) - 1.0f)) / P->length;
}

/* split_absorbed: kernel_A2E_MABU_aux.c:3-24 (A2E_MABU.py:700-705) -- the hand-off of the absorbed file to the dust solver
 * of one species: OUT[cell, f] = IN[cell, f] * RABS[f, IDUST] / sum_d ABU[cell, d] * RABS[f, d].  RABS is double; `den` is a
 * float that takes the double products one after the other, the quotient is formed in double and stored as float. */
void orc_split_absorbed(int idust, int n, int nfreq, int ndust, const double *rabs, const float *abu, const float *in, float *out) {
    #pragma omp parallel for
    for (int icell = 0; icell < n; icell++)
        for (int ifreq = 0; ifreq < nfreq; ifreq++) {
            float den = 0.0f;
            for (int d = 0; d < ndust; d++) den += abu[(long)icell * ndust + d] * rabs[ifreq * ndust + d];
            out[(long)icell * nfreq + ifreq] = in[(long)icell * nfreq + ifreq] * rabs[ifreq * ndust + idust] / den;
        }
}

/* =================================================================================================
 * Map ray-tracer (kernel_ASOC_map.c:496-875; MAP_INTERPOLATION==0, ROI_MAP==0)
 * ================================================================================================= */
static void map_pixel(const OrcParams *P, const nav_t *N, int id, float map_dx, int npx, int npy, float *map,
                      const float *emit, v3 DIR, v3 RA, v3 DE, float abs_, float sca_, v3 CENTRE, v3 INTOBS,
                      const float *opt, float *savetau, int save_colden, uint64_t *steps) {
    const int NX = P->nx, NY = P->ny, NZ = P->nz;
    if (id >= npx * npy) return;
    float DTAU, TAU = 0.0f, PHOTONS = 0.0f, colden = 0.0f, sx, sy, sz, dens, em;
    v3 POS, TMP;
    int ind, level = 0, oind, olevel;
    int i = id % npx, j = id / npx;
    if (INTOBS.x > -1e10f) {                                                 /* kernel_ASOC_map.c:534-553 */
        float phi = M_TWOPI * i / (float)(npx);
        phi += M_PI_F;
        float pix = M_TWOPI / npx;
        float theta = pix * (j - (npy - 1) / 2);
        POS = INTOBS;
        TMP.x = cosf(theta) * cosf(phi); TMP.y = cosf(theta) * sinf(phi); TMP.z = sinf(theta);
        if (fabsf(TMP.x) < 1.0e-5f) TMP.x = 1.0e-5f;
        if (fabsf(TMP.y) < 1.0e-5f) TMP.y = 1.0e-5f;
        if (fabsf(TMP.z) < 1.0e-5f) TMP.z = 1.0e-5f;
        if (fmodf(POS.x, 1.0f) < 1.0e-5f) POS.x += 2.0e-5f;
        if (fmodf(POS.y, 1.0f) < 1.0e-5f) POS.y += 2.0e-5f;
        if (fmodf(POS.z, 1.0f) < 1.0e-5f) POS.z += 2.0e-5f;
    } else {                                                                 /* kernel_ASOC_map.c:554-640 */
        POS.x = CENTRE.x + (i - 0.5f * (npx - 1)) * map_dx * RA.x + (j - 0.5f * (npy - 1)) * map_dx * DE.x;
        POS.y = CENTRE.y + (i - 0.5f * (npx - 1)) * map_dx * RA.y + (j - 0.5f * (npy - 1)) * map_dx * DE.y;
        POS.z = CENTRE.z + (i - 0.5f * (npx - 1)) * map_dx * RA.z + (j - 0.5f * (npy - 1)) * map_dx * DE.z;
        float far_ = (float)(NX + NY + NZ);
        POS.x += far_ * DIR.x; POS.y += far_ * DIR.y; POS.z += far_ * DIR.z;
        if (NX < 200) {
            sx = ((DIR.x >= 0.0f) ? (NX - POS.x) : (0.0f - POS.x)) / (-DIR.x) + M_EPS;
            sy = ((DIR.y >= 0.0f) ? (NY - POS.y) : (0.0f - POS.y)) / (-DIR.y) + M_EPS;
            sz = ((DIR.z >= 0.0f) ? (NZ - POS.z) : (0.0f - POS.z)) / (-DIR.z) + M_EPS;
            TMP.x = POS.x - sx * DIR.x; TMP.y = POS.y - sx * DIR.y; TMP.z = POS.z - sx * DIR.z;
            if (TMP.x < 0.0f || TMP.x > NX || TMP.y < 0.0f || TMP.y > NY || TMP.z < 0.0f || TMP.z > NZ) sx = 1e10f;
            TMP.x = POS.x - sy * DIR.x; TMP.y = POS.y - sy * DIR.y; TMP.z = POS.z - sy * DIR.z;
            if (TMP.x < 0.0f || TMP.x > NX || TMP.y < 0.0f || TMP.y > NY || TMP.z < 0.0f || TMP.z > NZ) sy = 1e10f;
            TMP.x = POS.x - sz * DIR.x; TMP.y = POS.y - sz * DIR.y; TMP.z = POS.z - sz * DIR.z;
            if (TMP.x < 0.0f || TMP.x > NX || TMP.y < 0.0f || TMP.y > NY || TMP.z < 0.0f || TMP.z > NZ) sz = 1e10f;
            sx = fminf_(sx, fminf_(sy, sz));
            POS.x = POS.x - sx * DIR.x; POS.y = POS.y - sx * DIR.y; POS.z = POS.z - sx * DIR.z;
        } else {
            float ex = (DIR.x > 0.0f) ? -M_EPS : M_EPS, ey = (DIR.y > 0.0f) ? -M_EPS : M_EPS, ez = (DIR.z > 0.0f) ? -M_EPS : M_EPS;
            sx = ((DIR.x >= 0.0f) ? (NX - POS.x) : (0.0f - POS.x)) / (-DIR.x);
            sy = ((DIR.y >= 0.0f) ? (NY - POS.y) : (0.0f - POS.y)) / (-DIR.y);
            sz = ((DIR.z >= 0.0f) ? (NZ - POS.z) : (0.0f - POS.z)) / (-DIR.z);
            TMP.x = POS.x - sx * DIR.x; TMP.y = POS.y - sx * DIR.y; TMP.z = POS.z - sx * DIR.z;
            TMP.x += ex; TMP.y += ey; TMP.z += ez;
            if (TMP.x < 0.0f || TMP.x > NX || TMP.y < 0.0f || TMP.y > NY || TMP.z < 0.0f || TMP.z > NZ) sx = 1e10f;
            TMP.x = POS.x - sy * DIR.x; TMP.y = POS.y - sy * DIR.y; TMP.z = POS.z - sy * DIR.z;
            TMP.x += ex; TMP.y += ey; TMP.z += ez;
            if (TMP.x < 0.0f || TMP.x > NX || TMP.y < 0.0f || TMP.y > NY || TMP.z < 0.0f || TMP.z > NZ) sy = 1e10f;
            TMP.x = POS.x - sz * DIR.x; TMP.y = POS.y - sz * DIR.y; TMP.z = POS.z - sz * DIR.z;
            TMP.x += ex; TMP.y += ey; TMP.z += ez;
            if (TMP.x < 0.0f || TMP.x > NX || TMP.y < 0.0f || TMP.y > NY || TMP.z < 0.0f || TMP.z > NZ) sz = 1e10f;
            sx = fminf_(sx, fminf_(sy, sz));
            POS.x = POS.x - sx * DIR.x; POS.y = POS.y - sx * DIR.y; POS.z = POS.z - sx * DIR.z;
            POS.x += ex; POS.y += ey; POS.z += ez;
        }
        TMP.x = -DIR.x; TMP.y = -DIR.y; TMP.z = -DIR.z;
        if (fabsf(TMP.x) < 1.0e-5f) TMP.x = 1.0e-5f;
        if (fabsf(TMP.y) < 1.0e-5f) TMP.y = 1.0e-5f;
        if (fabsf(TMP.z) < 1.0e-5f) TMP.z = 1.0e-5f;
    }
    index_g_map(N, &POS, &level, &ind);
    /* MAP_INTERPOLATION: two directions perpendicular to the line of sight, kernel_ASOC_map.c:656-684 */
    const int MI = P->map_interpolation;
    v3 ADIR = { 0, 0, 0 }, BDIR = { 0, 0, 0 }, MPOS, POS0;
    int slevel, sind, level0, ind0;
    float a, b, Adens, Bdens, Aemit, Bemit, K;
    if (MI > 0) {
        if (fabsf(TMP.x) > fabsf(TMP.y)) {
            if (fabsf(TMP.z) > fabsf(TMP.x)) { ADIR.x = 0.0005f; ADIR.y = 1.0f; ADIR.z = -TMP.y / TMP.z; }
            else                             { ADIR.x = -TMP.z / TMP.x; ADIR.y = 0.0005f; ADIR.z = 1.0f; }
        } else {
            if (fabsf(TMP.z) > fabsf(TMP.y)) { ADIR.x = 0.0005f; ADIR.y = 1.0f; ADIR.z = -TMP.y / TMP.z; }
            else                             { ADIR.x = 1.0f; ADIR.y = -TMP.x / TMP.y; ADIR.z = 0.0005f; }
        }
        ADIR = v3_norm(ADIR);
        BDIR.x = TMP.y * ADIR.z - TMP.z * ADIR.y;
        BDIR.y = TMP.z * ADIR.x - TMP.x * ADIR.z;
        BDIR.z = TMP.x * ADIR.y - TMP.y * ADIR.x;
        BDIR = v3_norm(BDIR);
    }
    while (ind >= 0) {                                                        /* kernel_ASOC_map.c:688-861 */
        oind = N->off[level] + ind; olevel = level;
        if (MI > 0) { POS0 = POS; ind0 = ind; level0 = level; K = ldexpf(1.0f, -level0); }
        sx = get_step_map(N, &POS, &TMP, &level, &ind);
        dens = N->dens[oind]; em = emit[oind];
        if (MI > 0) {                                                         /* :706-805 */
            const float lim = (MI == 2) ? 0.52f : 0.502f;
            if (MI == 2) {
                a = 0.22f * K;
                if (sx > a) {
                    sx = a;
                    POS.x = POS0.x + 0.22f * TMP.x; POS.y = POS0.y + 0.22f * TMP.y; POS.z = POS0.z + 0.22f * TMP.z;
                    ind = ind0; level = level0;
                    index_map(N, &POS, &level, &ind);
                }
            }
            float h = 0.5f * sx / K;
            slevel = level0; sind = ind0;
            MPOS.x = POS0.x + h * TMP.x; MPOS.y = POS0.y + h * TMP.y; MPOS.z = POS0.z + h * TMP.z;
            a = get_step_map(N, &MPOS, &ADIR, &slevel, &sind);
            a /= K;
            if (a <= lim && sind >= 0) { Adens = N->dens[N->off[slevel] + sind]; Aemit = emit[N->off[slevel] + sind]; }
            else {
                slevel = level0; sind = ind0; ADIR.x *= -1.0f; ADIR.y *= -1.0f; ADIR.z *= -1.0f;
                MPOS.x = POS0.x + h * TMP.x; MPOS.y = POS0.y + h * TMP.y; MPOS.z = POS0.z + h * TMP.z;
                a = get_step_map(N, &MPOS, &ADIR, &slevel, &sind);
                a /= K;
                if (a <= lim && sind >= 0) { Adens = N->dens[N->off[slevel] + sind]; Aemit = emit[N->off[slevel] + sind]; }
                else { a = 0.5f; Adens = 0.0f; Aemit = 0.0f; }
            }
            slevel = level0; sind = ind0;
            MPOS.x = POS0.x + h * TMP.x; MPOS.y = POS0.y + h * TMP.y; MPOS.z = POS0.z + h * TMP.z;
            b = get_step_map(N, &MPOS, &BDIR, &slevel, &sind);
            b /= K;
            if (b <= lim && sind >= 0) { Bdens = N->dens[N->off[slevel] + sind]; Bemit = emit[N->off[slevel] + sind]; }
            else {
                slevel = level0; sind = ind0; BDIR.x *= -1.0f; BDIR.y *= -1.0f; BDIR.z *= -1.0f;
                MPOS.x = POS0.x + h * TMP.x; MPOS.y = POS0.y + h * TMP.y; MPOS.z = POS0.z + h * TMP.z;
                b = get_step_map(N, &MPOS, &BDIR, &slevel, &sind);
                if (MI == 1) b /= K;                                          /* sic: missing in the MAP_INTERPOLATION==2 branch, :750 */
                if (b <= lim && sind >= 0) { Bdens = N->dens[N->off[slevel] + sind]; Bemit = emit[N->off[slevel] + sind]; }
                else { b = 0.5f; Bdens = 0.0f; Bemit = 0.0f; }
            }
            if (MI == 2) {
                a = clampf(a, 0.0f, 0.51f); b = clampf(b, 0.0f, 0.51f);
                dens = (0.5f - a) * Adens + (0.5f - b) * Bdens + (a + b) * dens;
                em   = (0.5f - a) * Aemit + (0.5f - b) * Bemit + (a + b) * em;
            } else {
                a = 0.5f - a; b = 0.5f - b;
                dens = (1.0f - a - b) * dens + a * Adens + b * Bdens;
                em   = (1.0f - a - b) * em + a * Aemit + b * Bemit;
            }
        }
        if (P->with_abu) DTAU = sx * dens * (opt[2 * (long)oind] + opt[2 * (long)oind + 1]);
        else             DTAU = sx * dens * (sca_ + abs_);
        if ((P->level_threshold <= 0 || olevel >= P->level_threshold) &&
            (P->roi_map <= 0 || in_roi(N, P->roi, olevel, oind - N->off[olevel]) >= 0)) {   /* ROI_MAP: kernel_ASOC_map.c:37-56, 822 */
            if (DTAU < 1.0e-3f) PHOTONS += expf(-TAU) * (1.0f - 0.5f * DTAU) * sx * em * dens;
            else                PHOTONS += expf(-TAU) * ((1.0f - expf(-DTAU)) / DTAU) * sx * em * dens;
        }
        TAU += DTAU;
        if (save_colden > 0) colden += sx * dens;
        (*steps)++;
    }
    map[id] = PHOTONS;
    savetau[id] = save_colden ? colden * P->length : TAU;
}

void orc_mapping(const OrcParams *P, const OrcGrid *G, float map_dx, int npx, int npy, float *map, const float *emit,
                 const float *dir, const float *ra, const float *de, float abs_, float sca_, const float *centre,
                 const float *intobs, const float *opt, float *savetau, int save_colden, OrcCounters *C) {
    v3 D = { dir[0], dir[1], dir[2] }, R = { ra[0], ra[1], ra[2] }, E = { de[0], de[1], de[2] };
    v3 CE = { centre[0], centre[1], centre[2] }, IO = { intobs[0], intobs[1], intobs[2] };
    nav_t N = nav_make(P, G);
    uint64_t steps = 0;
    #pragma omp parallel for schedule(runtime) reduction(+:steps)
    for (int id = 0; id < npx * npy; id++) {
        uint64_t s = 0;
        map_pixel(P, &N, id, map_dx, npx, npy, map, emit, D, R, E, abs_, sca_, CE, IO, opt, savetau, save_colden, &s);
        steps += s;
    }
    if (C) C->steps += steps;
}

/* Per-level Mapping: kernel_ASOC_map_H.c:380-506 (its IndexG / Index / GetStep, :179-347, are the ones of
 * kernel_ASOC_map.c -- same EPS / PEPS, same z<=0 containment quirk, double Index for NX>100).  One image per
 * hierarchy level: MAP[ilev*npx*npy + id] = emission of the cells of level ilev along the line of sight, attenuated
 * by all the material in front.  The ray set-up differs from kernel_ASOC_map.c: the start point is found from behind
 * the cloud as the LAST face crossing (largest step that still lies inside), without the 1e-5 clamp of the direction
 * components; the perspective branch has the other sign / axis convention (:418-423).  colden may be NULL
 * (WITH_COLDEN == 0).  P->maph_literal = 1 restates the file's own Index(), which forgets to store the root coordinates
 * when a ray climbs out of a refined region into a root-grid leaf (:250) and then continues from the wrong place;
 * 0 (default) is the Index() of kernel_ASOC_map.c, which has the store -- the expectation for the product. */
static void map_levels_pixel(const OrcParams *P, const nav_t *N, int id, float map_dx, int npx, int npy, float *map,
                             const float *emit, v3 DIR, v3 RA, v3 DE, float abs_, float sca_, v3 CENTRE, v3 INTOBS,
                             const float *opt, float *colden_out) {
    const int NX = P->nx, NY = P->ny, NZ = P->nz, LEVELS = P->levels;
    float DTAU, TAU = 0.0f, colden = 0.0f, sx, sy, sz;
    float PHOTONS[16];
    for (int l = 0; l < LEVELS; l++) PHOTONS[l] = 0.0f;
    v3 POS, TMP;
    int ind, level = 0, oind, olevel;
    int i = id % npx, j = id / npx;
    if (INTOBS.x > -1e10f) {                                                  /* :403-430 */
        float phi = M_TWOPI * i / (float)(npx);
        phi += M_PI_F;
        float pix = M_TWOPI / npx;
        float theta = pix * (j - (npy - 1) / 2);
        POS = INTOBS;
        TMP.x = -cosf(theta) * sinf(phi); TMP.y = -cosf(theta) * cosf(phi); TMP.z = +sinf(theta);
        if (fabsf(TMP.x) < 1.0e-5f) TMP.x = 1.0e-5f;
        if (fabsf(TMP.y) < 1.0e-5f) TMP.y = 1.0e-5f;
        if (fabsf(TMP.z) < 1.0e-5f) TMP.z = 1.0e-5f;
        if (fmodf(POS.x, 1.0f) < 1.0e-5f) POS.x += 2.0e-5f;
        if (fmodf(POS.y, 1.0f) < 1.0e-5f) POS.y += 2.0e-5f;
        if (fmodf(POS.z, 1.0f) < 1.0e-5f) POS.z += 2.0e-5f;
    } else {                                                                  /* :431-457 */
        POS.x = CENTRE.x + (i - 0.5f * (npx - 1)) * map_dx * RA.x + (j - 0.5f * (npy - 1)) * map_dx * DE.x;
        POS.y = CENTRE.y + (i - 0.5f * (npx - 1)) * map_dx * RA.y + (j - 0.5f * (npy - 1)) * map_dx * DE.y;
        POS.z = CENTRE.z + (i - 0.5f * (npx - 1)) * map_dx * RA.z + (j - 0.5f * (npy - 1)) * map_dx * DE.z;
        float far_ = (float)(NX + NY + NZ);
        POS.x -= far_ * DIR.x; POS.y -= far_ * DIR.y; POS.z -= far_ * DIR.z;
        if (DIR.x >= 0.0f) sx = (NX - POS.x) / (DIR.x + 1.0e-10f) - M_EPS; else sx = (0.0f - POS.x) / DIR.x - M_EPS;
        if (DIR.y >= 0.0f) sy = (NY - POS.y) / (DIR.y + 1.0e-10f) - M_EPS; else sy = (0.0f - POS.y) / DIR.y - M_EPS;
        if (DIR.z >= 0.0f) sz = (NZ - POS.z) / (DIR.z + 1.0e-10f) - M_EPS; else sz = (0.0f - POS.z) / DIR.z - M_EPS;
        TMP.x = POS.x + sx * DIR.x; TMP.y = POS.y + sx * DIR.y; TMP.z = POS.z + sx * DIR.z;
        if (TMP.x <= 0.0f || TMP.x >= NX || TMP.y <= 0.0f || TMP.y >= NY || TMP.z <= 0.0f || TMP.z >= NZ) sx = -1e10f;
        TMP.x = POS.x + sy * DIR.x; TMP.y = POS.y + sy * DIR.y; TMP.z = POS.z + sy * DIR.z;
        if (TMP.x <= 0.0f || TMP.x >= NX || TMP.y <= 0.0f || TMP.y >= NY || TMP.z <= 0.0f || TMP.z >= NZ) sy = -1e10f;
        TMP.x = POS.x + sz * DIR.x; TMP.y = POS.y + sz * DIR.y; TMP.z = POS.z + sz * DIR.z;
        if (TMP.x <= 0.0f || TMP.x >= NX || TMP.y <= 0.0f || TMP.y >= NY || TMP.z <= 0.0f || TMP.z >= NZ) sz = -1e10f;
        sx = fmaxf(sx, fmaxf(sy, sz));
        POS.x = POS.x + sx * DIR.x; POS.y = POS.y + sx * DIR.y; POS.z = POS.z + sx * DIR.z;
        TMP.x = -DIR.x; TMP.y = -DIR.y; TMP.z = -DIR.z;
    }
    index_g_map(N, &POS, &level, &ind);
    while (ind >= 0) {                                                        /* :468-488 */
        oind = N->off[level] + ind; olevel = level;
        sx = get_step_any(N, &POS, &TMP, &level, &ind, M_PEPS, P->maph_literal ? 2 : 1);
        if (P->with_abu) DTAU = sx * N->dens[oind] * (opt[2 * (long)oind] + opt[2 * (long)oind + 1]);
        else             DTAU = sx * N->dens[oind] * (sca_ + abs_);
        if (DTAU < 1.0e-3f) PHOTONS[olevel] += expf(-TAU) * (1.0f - 0.5f * DTAU) * sx * emit[oind] * N->dens[oind];
        else                PHOTONS[olevel] += expf(-TAU) * ((1.0f - expf(-DTAU)) / DTAU) * sx * emit[oind] * N->dens[oind];
        TAU += DTAU;
        colden += sx * N->dens[oind];
    }
    for (int l = 0; l < LEVELS; l++) map[(long)l * npx * npy + id] = PHOTONS[l];
    if (colden_out) colden_out[id] = colden * P->length;
}

void orc_mapping_levels(const OrcParams *P, const OrcGrid *G, float map_dx, int npx, int npy, float *map, const float *emit,
                        const float *dir, const float *ra, const float *de, float abs_, float sca_, const float *centre,
                        const float *intobs, const float *opt, float *colden) {
    v3 D = { dir[0], dir[1], dir[2] }, R = { ra[0], ra[1], ra[2] }, E = { de[0], de[1], de[2] };
    v3 CE = { centre[0], centre[1], centre[2] }, IO = { intobs[0], intobs[1], intobs[2] };
    nav_t N = nav_make(P, G);
    #pragma omp parallel for schedule(runtime)
    for (int id = 0; id < npx * npy; id++)
        map_levels_pixel(P, &N, id, map_dx, npx, npy, map, emit, D, R, E, abs_, sca_, CE, IO, opt, colden);
}

/* PSTau: kernel_ASOC_map.c:1545-1599 -- optical depth and column density from every point source towards the observer */
void orc_ps_tau(const OrcParams *P, const OrcGrid *G, int no, const float *pspos, const float *dir, float abs_, float sca_,
                const float *opt, float *pscolden, float *pstau) {
    nav_t N = nav_make(P, G);
    v3 D = { dir[0], dir[1], dir[2] };
    for (int id = 0; id < no; id++) {
        float TAU = 0.0f, colden = 0.0f, sx, DTAU;
        v3 POS = { pspos[3 * id], pspos[3 * id + 1], pspos[3 * id + 2] };
        int ind, level = 0, oind;
        index_g_map(&N, &POS, &level, &ind);
        while (ind >= 0) {
            oind = N.off[level] + ind;
            sx = get_step_map(&N, &POS, &D, &level, &ind);
            if (P->with_abu) DTAU = sx * N.dens[oind] * (opt[2 * (long)oind] + opt[2 * (long)oind + 1]);
            else             DTAU = sx * N.dens[oind] * (sca_ + abs_);
            TAU += DTAU;
            colden += sx * N.dens[oind];
        }
        pscolden[id] = colden * P->length;
        pstau[id] = TAU;
    }
}

/* HealpixMapping: kernel_ASOC_map.c:890-966 (the NSIDE macro of the map program = nside here) */
static void map_pix2ang(int nside, int ipix, float *phi, float *theta) {      /* kernel_ASOC_map.c:101-140 */
    int nl2, nl4, npix, ncap, iring, iphi, ip, ipix1;
    float fact1, fact2, fodd, hip, fihip;
    npix = 12 * nside * nside; ipix1 = ipix + 1; nl2 = 2 * nside; nl4 = 4 * nside;
    ncap = 2 * nside * (nside - 1); fact1 = 1.5f * nside; fact2 = 3.0f * nside * nside;
    if (ipix1 <= ncap) {
        hip = ipix1 / 2.0f; fihip = (int)(hip);
        iring = (int)(sqrtf(hip - sqrtf(fihip))) + 1;
        iphi = ipix1 - 2 * iring * (iring - 1);
        *theta = acosf(1.0f - iring * iring / fact2);
        *phi = (iphi - 0.5f) * M_PI_F / (2.0f * iring);
    } else if (ipix1 <= nl2 * (5 * nside + 1)) {
        ip = ipix1 - ncap - 1;
        iring = (int)(ip / nl4) + nside;
        iphi = (ip % nl4) + 1;
        fodd = 0.5f * (1 + (iring + nside) % 2);
        *theta = acosf((nl2 - iring) / fact1);
        *phi = (iphi - fodd) * M_PI_F / (2.0f * nside);
    } else {
        ip = npix - ipix1 + 1;
        hip = ip / 2.0f; fihip = (int)(hip);
        iring = (int)(sqrtf(hip - sqrtf(fihip))) + 1;
        iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
        *theta = acosf(-1.0f + iring * iring / fact2);
        *phi = (iphi - 0.5f) * M_PI_F / (2.0f * iring);
    }
}
void orc_healpix_mapping(const OrcParams *P, const OrcGrid *G, int nside, float *map, const float *emit, float abs_,
                         float sca_, const float *intobs, const float *opt, float *savetau, int save_colden) {
    nav_t N = nav_make(P, G);
    #pragma omp parallel for schedule(runtime)
    for (int id = 0; id < 12 * nside * nside; id++) {
        float DTAU, TAU = 0.0f, PHOTONS = 0.0f, colden = 0.0f, dx, theta, phi;
        v3 POS, TMP; int ind, level = 0, oind;
        map_pix2ang(nside, id, &phi, &theta);
        TMP.x = -sinf(theta) * cosf(phi); TMP.y = -sinf(theta) * sinf(phi); TMP.z = cosf(theta);
        if (fabsf(TMP.x) < 1.0e-5f) TMP.x = 1.0e-5f;
        if (fabsf(TMP.y) < 1.0e-5f) TMP.y = 1.0e-5f;
        if (fabsf(TMP.z) < 1.0e-5f) TMP.z = 1.0e-5f;
        POS.x = intobs[0]; POS.y = intobs[1]; POS.z = intobs[2];
        if (fmodf(POS.x, 1.0f) < 1.0e-5f || fmodf(POS.x, 1.0f) < 0.99999f) POS.x += 2.0e-5f;   /* sic */
        if (fmodf(POS.y, 1.0f) < 1.0e-5f || fmodf(POS.y, 1.0f) < 0.99999f) POS.y += 2.0e-5f;
        if (fmodf(POS.z, 1.0f) < 1.0e-5f || fmodf(POS.z, 1.0f) < 0.99999f) POS.z += 2.0e-5f;
        index_g_map(&N, &POS, &level, &ind);
        while (ind >= 0) {
            oind = N.off[level] + ind;
            const int olevel = level;
            dx = get_step_map(&N, &POS, &TMP, &level, &ind);
            if (P->with_abu) DTAU = dx * N.dens[oind] * (opt[2 * (long)oind] + opt[2 * (long)oind + 1]);
            else             DTAU = dx * N.dens[oind] * (sca_ + abs_);
            if (P->roi_map > 0 && in_roi(&N, P->roi, olevel, oind - N.off[olevel]) < 0) ;        /* ROI_MAP, kernel_ASOC_map.c:947-957 */
            else if (DTAU < 1.0e-3f) PHOTONS += expf(-TAU) * (1.0f - 0.5f * DTAU) * dx * emit[oind] * N.dens[oind];
            else                PHOTONS += expf(-TAU) * ((1.0f - expf(-DTAU)) / DTAU) * dx * emit[oind] * N.dens[oind];
            TAU += DTAU;
            colden += dx * N.dens[oind];
        }
        map[id] = PHOTONS;
        savetau[id] = (save_colden > 0) ? colden * P->length : TAU;
    }
}

/* =================================================================================================
 * Scattered light with peel-off (kernel_ASOC_sca.c).  MAX_SCATTERINGS 30, no Russian roulette (:5-6).
 * flavour 0 = SimRAM_PS (:1462-1937: native_* maths, expm1, float log argument),
 * flavour 1 = SimRAM_PB (:471-1088: double-precision log argument).
 * ================================================================================================= */
typedef struct { const OrcParams *P; const OrcSimBufs *B; const OrcScaBufs *O; nav_t N; OrcCounters c; } sca_t;

/* WITH_MSF: pick the scattering dust species of cell `oind` with probability ABU*SCA / sum (kernel_ASOC_sca.c:339-346,
   992-999, 1052-1061 ...; kernel_ASOC.c:777-794).  Draws one random number; returns 0 without a draw when MSF is off. */
/* kernel_ASOC_sca.c:349-355, 392-398, 1343-1349, 1387-1393: the `#ifdef HG_TEST` branch that is live in the shipped
   SimRAM_HP / SimRAM_CL (HG_TEST is #defined, as 0, in kernel_ASOC_aux.c:1) */
static inline float hg_test_fraction(float cos_theta) {
    const float G = 0.65f;
    return (1.0f / (4.0f * PI_F)) * (1.0f - G * G) / powf(1.0f + G * G - 2.0f * G * cos_theta, 1.5f);
}
static int msf_pick(sca_t *S, rng_t *r, int oind) {
    const OrcParams *P = S->P; const OrcSimBufs *B = S->B;
    if (P->with_msf <= 0) return 0;
    float dx = B->opt[2 * (long)oind + 1];
    float ds = 0.99999f * rnd(r);
    int idust;
    for (idust = 0; idust < P->ndust; idust++) {
        ds -= B->abu[idust + (long)P->ndust * oind] * B->sca_v[idust] / dx;
        if (ds <= 0.0f) break;
    }
    return idust < P->ndust ? idust : P->ndust - 1;   /* the sca kernels do not clamp (they would read past DSC); kernel_ASOC.c:789 does */
}

static void sca_propagate(sca_t *S, rng_t *r, v3 pos, v3 dir, int level, int ind, float photons, int flavour) {
    const OrcParams *P = S->P; const OrcSimBufs *B = S->B; const OrcScaBufs *O = S->O; const nav_t *N = &S->N;
    int ind0, level0, oind = 0, scatterings = 0;
    float tau, dtau, ds, dx, delta, free_path, W, kabs, ksca;
    v3 pos0;
    /* forced first scattering: kernel_ASOC_sca.c:888-910 / 1720-1750 */
    if (P->ffs > 0) {
        pos0 = pos; ind0 = ind; level0 = level; tau = 0.0f;
        while (ind0 >= 0) {
            oind = N->off[level0] + ind0;
            ds = get_step(N, &pos0, &dir, &level0, &ind0);
            ksca = P->with_abu ? B->opt[2 * (long)oind + 1] : B->sca;
            tau += ds * N->dens[oind] * ksca;
            S->c.steps++;
        }
        if (flavour >= 2) {                       /* SimRAM_HP :279-287, SimRAM_CL :1257-1264: no draw for an empty line of sight */
            if (tau < 1.0e-22f) return;
            W = 1.0f - expf(-tau); free_path = (float)(-log(1.0 - W * rnd(r)));
        } else {
            if (tau < 1.0e-22f) ind = -1;
            if (flavour == 0) { W = -expm1f(-tau); free_path = -logf(1.0f - W * rnd(r)); }
            else              { W = 1.0f - expf(-tau); free_path = (float)(-log(1.0 - W * rnd(r))); }
        }
        photons *= W;
    } else {
        free_path = -logf(rnd(r));
    }
    while (ind >= 0) {
        tau = 0.0f;
        while (ind >= 0) {
            ind0 = ind; level0 = level; pos0 = pos; oind = N->off[level0] + ind0;
            ds = get_step(N, &pos, &dir, &level, &ind);
            ksca = P->with_abu ? B->opt[2 * (long)oind + 1] : B->sca;
            dtau = ds * N->dens[oind] * ksca;
            S->c.steps++;
            if (free_path < (tau + dtau)) { ind = ind0; break; }
            tau += dtau;
            if (P->mirror > 0 && ind < 0) mirror(N, P->mirror, P->mirror_exact, &pos, &dir, &level, &ind);   /* kernel_ASOC_sca.c:280, 940, 1280, 1780 */
        }
        if (ind < 0) break;
        scatterings++; S->c.scatterings++;
        dtau = free_path - tau;
        if (P->with_abu) { kabs = B->opt[2 * (long)oind]; ksca = B->opt[2 * (long)oind + 1]; }
        else             { kabs = B->abs; ksca = B->sca; }
        dx = dtau / (ksca * N->dens[oind]);
        /* sic: the reference converts with the level of the *next* cell (kernel_ASOC_sca.c:958,1797), which displaces
           the scattering point at refinement boundaries; sca_exact_level=1 is the geometrically exact variant that the
           tests use as the expectation for the production CUDA kernel */
        dx = ldexpf(dx, P->sca_exact_level ? level0 : level);
        pos0.x += dx * dir.x; pos0.y += dx * dir.y; pos0.z += dx * dir.z;
        photons *= expf(-free_path * kabs / ksca);
        const float cclamp = (flavour >= 2) ? 0.9999f : 0.999f;            /* :356,1329 vs :991,1830 */
        if (O->ndir < 0) {
            /* Healpix image seen by an observer at ODIRS[0] (root-grid units), NSIDE = -NDIR:
               kernel_ASOC_sca.c:312-378 (HP), 968-1008 (PB), 1309-1352 (CL), 1807-1847 (PS) */
            v3 p = pos0, q = pos0, od;
            int pind = ind0, plevel = level0, po;
            root_pos(N, &q, level0, ind0);
            od.x = O->odirs[0] - q.x; od.y = O->odirs[1] - q.y; od.z = O->odirs[2] - q.z;
            dx = sqrtf(od.x * od.x + od.y * od.y + od.z * od.z);
            delta = 1.0f / (dx * dx);
            od = v3_norm(od);
            tau = 0.0f;
            while (dx > 0 && pind >= 0) {
                po = N->off[plevel] + pind;
                ds = get_step(N, &p, &od, &plevel, &pind);
                if (flavour == 1) ds = (float)(fminf_(dx, ds) + 1.0e-6);      /* sic: double literal in SimRAM_PB :982 */
                else              ds = fminf_(dx, ds) + 1.0e-6f;
                dx -= ds;
                if (P->with_abu) tau += ds * N->dens[po] * (B->opt[2 * (long)po] + B->opt[2 * (long)po + 1]);
                else             tau += ds * N->dens[po] * (B->abs + B->sca);
                S->c.steps++;
            }
            S->c.peels++;
            float cos_theta = clampf(dir.x * od.x + dir.y * od.y + dir.z * od.z, -cclamp, +cclamp);
            int idust = msf_pick(S, r, N->off[level0] + ind0);
            if (flavour >= 2 && P->hg_test) delta *= photons * hg_test_fraction(cos_theta) * ((tau > TAULIM) ? (1.0f - expf(-tau)) : (tau * (1.0f - 0.5f * tau)));
            else delta *= photons * expf(-tau) * B->dsc[idust * P->bins + clampi((int)(P->bins * (1.0f + cos_theta) * 0.5f), 0, P->bins - 1)];
            float theta = acosf(-od.z), phi = atan2f(od.y, od.x);
            int ipix = orc_ang2pix_ring(-O->ndir, phi, theta);
            if (ipix >= 0) addf(&O->out[ipix], delta);
        } else
        /* peel-off towards every observer (orthographic maps): kernel_ASOC_sca.c:1010-1046 / 1849-1885 */
        for (int idir = 0; idir < O->ndir; idir++) {
            v3 p = pos0, od = { O->odirs[3 * idir], O->odirs[3 * idir + 1], O->odirs[3 * idir + 2] };
            int pind = ind0, plevel = level0, po;
            tau = 0.0f;
            while (pind >= 0) {
                po = N->off[plevel] + pind;
                ds = get_step(N, &p, &od, &plevel, &pind);
                if (P->with_abu) tau += ds * N->dens[po] * (B->opt[2 * (long)po] + B->opt[2 * (long)po + 1]);
                else             tau += ds * N->dens[po] * (B->abs + B->sca);
                S->c.steps++;
            }
            S->c.peels++;
            float cos_theta = clampf(dir.x * od.x + dir.y * od.y + dir.z * od.z, -cclamp, +cclamp);
            int idust = msf_pick(S, r, N->off[level0] + ind0);
            if (flavour >= 2 && P->hg_test) delta = photons * hg_test_fraction(cos_theta) * ((tau > TAULIM) ? (1.0f - expf(-tau)) : (tau * (1.0f - 0.5f * tau)));
            else delta = photons * expf(-tau) * B->dsc[idust * P->bins + clampi((int)(P->bins * (1.0f + cos_theta) * 0.5f), 0, P->bins - 1)];
            p.x -= O->centre[0]; p.y -= O->centre[1]; p.z -= O->centre[2];
            const float *ra = O->ora + 3 * idir, *de = O->ode + 3 * idir;
            int i = (int)((0.5f * O->npix_x - 0.00005f) + (p.x * ra[0] + p.y * ra[1] + p.z * ra[2]) / O->map_dx);
            int j = (int)((0.5f * O->npix_y - 0.00005f) + (p.x * de[0] + p.y * de[1] + p.z * de[2]) / O->map_dx);
            if (i >= 0 && j >= 0 && i < O->npix_x && j < O->npix_y)
                addf(&O->out[i + idir * O->npix_x * O->npix_y + j * O->npix_x], delta);
        }
        pos = pos0; ind = ind0; level = level0;
        scatter(&dir, B->csc + msf_pick(S, r, N->off[level0] + ind0) * P->bins, P->bins, r);
        free_path = -logf(rnd(r));
        if (scatterings == 30) break;
    }
}

/* point sources inside or outside the cloud: emission part of kernel_ASOC_sca.c:1520-1717 (== :627-808) */
static int emit_ps(const OrcParams *P, const OrcSimBufs *B, const nav_t *N, rng_t *rng, int III, v3 *pos, v3 *dir,
                   int *level, int *ind, float *photons) {
    const int NX = P->nx, NY = P->ny, NZ = P->nz;
    float phi = TWOPI * rnd(rng);
    float cos_theta = 0.999997f - 1.999995f * rnd(rng);
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    dir->x = sin_theta * cosf(phi); dir->y = sin_theta * sinf(phi); dir->z = cos_theta;
    int ips = III % P->no_ps;
    *photons = B->ps[ips];
    v3 src = { B->pspos[3 * ips], B->pspos[3 * ips + 1], B->pspos[3 * ips + 2] };
    *pos = src;
    index_g(N, pos, level, ind);
    if (*ind < 0 || *ind >= P->cells) {
        if (P->ps_method == 0) { surface(N, pos, dir); index_g(N, pos, level, ind); }
        else if (P->ps_method == 1) {
            *pos = src;
            if (pos->z > NZ) { if (dir->z > 0.0f) dir->z = -dir->z; }
            else if (pos->z < 0.0f) { if (dir->z < 0.0f) dir->z = -dir->z; }
            else if (pos->x > NX) { if (dir->x > 0.0f) dir->x = -dir->x; }
            else if (pos->x < 0.0f) { if (dir->x < 0.0f) dir->x = -dir->x; }
            else if (pos->y > NY) { if (dir->y > 0.0f) dir->y = -dir->y; }
            else if (pos->y < 0.0f) { if (dir->y < 0.0f) dir->y = -dir->y; }
            surface(N, pos, dir); *photons *= 0.5f; index_g(N, pos, level, ind);
        }
        /* methods 2/4/5 in the scattering kernels read XPS_* through float pointers although the host uploads
           int32 (kernel_ASOC_sca.c:1486-1488 vs ASOCS.py:267-269): not restated, rejected by the host. */
    }
    return 0;
}

void orc_sca_ps(const OrcParams *P, const OrcGrid *G, const OrcSimBufs *B, const OrcScaBufs *O, int global,
                int packets, int batch, float seed, OrcCounters *C) {
    (void)packets;
    uint64_t np = 0, ns = 0, nsc = 0, npl = 0;
    #pragma omp parallel for schedule(runtime) reduction(+:np,ns,nsc,npl)
    for (int id = 0; id < global; id++) {
        sca_t S; S.P = P; S.B = B; S.O = O; S.N = nav_make(P, G); memset(&S.c, 0, sizeof(S.c));
        rng_t rng; rng_seed(&rng, seed, (uint64_t)id);
        int ind = -1, level = 0; v3 pos = { 0, 0, 0 }, dir = { 0, 0, 0 }; float photons = 0.0f;
        for (int III = 0; III < batch; III++) {
            emit_ps(P, B, &S.N, &rng, III, &pos, &dir, &level, &ind, &photons);
            dir_fix(&dir);
            S.c.packets++;
            sca_propagate(&S, &rng, pos, dir, level, ind, photons, 0);
        }
        np += S.c.packets; ns += S.c.steps; nsc += S.c.scatterings; npl += S.c.peels;
    }
    if (C) { C->packets += np; C->steps += ns; C->scatterings += nsc; C->peels += npl; }
}

void orc_sca_pb(const OrcParams *P, const OrcGrid *G, const OrcSimBufs *B, const OrcScaBufs *O, int global,
                int source, int packets, int batch, float seed, float bg, OrcCounters *C) {
    const int NX = P->nx, NY = P->ny, NZ = P->nz;
    const int AREA = 2 * (NX * NY + NY * NZ + NZ * NX);
    uint64_t np = 0, ns = 0, nsc = 0, npl = 0;
    #pragma omp parallel for schedule(runtime) reduction(+:np,ns,nsc,npl)
    for (int id = 0; id < global; id++) {
        if (source == 1 && id >= 8 * AREA) continue;
        if (source == 3 && (!(P->with_roi_load > 0) || id >= 100 * packets)) continue;
        sca_t S; S.P = P; S.B = B; S.O = O; S.N = nav_make(P, G); memset(&S.c, 0, sizeof(S.c));
        rng_t rng; rng_seed(&rng, seed, (uint64_t)id);
        roi_src_t rs;
        if (source == 3) roi_src_init(P, id, packets, batch, &rs);
        int ind = -1, level = 0, side = 0;
        float X0 = 0, Y0 = 0, Z0 = 0, DX = 1, DY = 1, DZ = 1, photons = 0.0f;
        v3 pos = { 0, 0, 0 }, dir = { 0, 0, 0 };
        if (source == 1) {                                                   /* kernel_ASOC_sca.c:540-569 */
            ind = id % AREA;
            if (ind < NY * NZ) { side = 0; X0 = S_PEPS; Y0 = ind % NY; Z0 = ind / NY; DX = 0.0f; }
            else { ind -= NY * NZ;
            if (ind < NY * NZ) { side = 1; X0 = NX - S_PEPS; Y0 = ind % NY; Z0 = ind / NY; DX = 0.0f; }
            else { ind -= NY * NZ;
            if (ind < NX * NZ) { side = 2; Y0 = S_PEPS; X0 = ind % NX; Z0 = ind / NX; DY = 0.0f; }
            else { ind -= NX * NZ;
            if (ind < NX * NZ) { side = 3; Y0 = NY - S_PEPS; X0 = ind % NX; Z0 = ind / NX; DY = 0.0f; }
            else { ind -= NX * NZ;
            if (ind < NX * NY) { side = 4; Z0 = S_PEPS; X0 = ind % NX; Y0 = ind / NX; DZ = 0.0f; }
            else { ind -= NX * NY; side = 5; Z0 = NZ - S_PEPS; X0 = ind % NX; Y0 = ind / NX; DZ = 0.0f; } } } } }
        }
        for (int III = 0; III < batch; III++) {
            if (source == 0) emit_ps(P, B, &S.N, &rng, III, &pos, &dir, &level, &ind, &photons);
            if (source == 1) {
                pos.x = clampf(X0 + DX * rnd(&rng), S_PEPS, NX - S_PEPS);
                pos.y = clampf(Y0 + DY * rnd(&rng), S_PEPS, NY - S_PEPS);
                pos.z = clampf(Z0 + DZ * rnd(&rng), S_PEPS, NZ - S_PEPS);
                float cos_theta = sqrtf(rnd(&rng));
                float phi = TWOPI * rnd(&rng);
                float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
                float v1 = sin_theta * cosf(phi), v2 = sin_theta * sinf(phi);
                switch (side) {
                case 0: dir.x =  cos_theta; dir.y = v1; dir.z = v2; break;
                case 1: dir.x = -cos_theta; dir.y = v1; dir.z = v2; break;
                case 2: dir.y =  cos_theta; dir.x = v1; dir.z = v2; break;
                case 3: dir.y = -cos_theta; dir.x = v1; dir.z = v2; break;
                case 4: dir.z =  cos_theta; dir.x = v1; dir.y = v2; break;
                case 5: dir.z = -cos_theta; dir.x = v1; dir.y = v2; break;
                }
                photons = bg;
                index_g(&S.N, &pos, &level, &ind);
            }
            if (source == 3 && !roi_src_emit(P, B, &S.N, &rs, &rng, III, &pos, &dir, &level, &ind, &photons)) continue;
            dir_fix(&dir);
            S.c.packets++;
            sca_propagate(&S, &rng, pos, dir, level, ind, photons, 1);
        }
        np += S.c.packets; ns += S.c.steps; nsc += S.c.scatterings; npl += S.c.peels;
    }
    if (C) { C->packets += np; C->steps += ns; C->scatterings += nsc; C->peels += npl; }
}

/* ---- scattered light, Healpix background: SimRAM_HP, kernel_ASOC_sca.c:40-470 --------------------- */
void orc_sca_hp(const OrcParams *P, const OrcGrid *G, const OrcSimBufs *B, const OrcScaBufs *O, int global,
                int packets, int batch, float seed, OrcCounters *C) {
    (void)packets;
    const int NX = P->nx, NY = P->ny, NZ = P->nz;
    uint64_t np = 0, ns = 0, nsc = 0, npl = 0;
    #pragma omp parallel for schedule(runtime) reduction(+:np,ns,nsc,npl)
    for (int id = 0; id < global; id++) {
        sca_t S; S.P = P; S.B = B; S.O = O; S.N = nav_make(P, G); memset(&S.c, 0, sizeof(S.c));
        rng_t rng; rng_seed(&rng, seed, (uint64_t)id);
        const float Rout = 0.5f * sqrtf(1.0f * NX * NX + NY * NY + NZ * NZ);
        int ind = -1, level = 0, ind0, level0;
        float photons, phi, theta, x, ds, dx;
        v3 pos, pos0, dir;
        for (int III = 0; III < batch; III++) {
            if (P->hpbg_weighted < 1) {
                ind = clampi((int)(floorf(rnd(&rng) * 49152)), 0, 49151);
            } else {                                                         /* :117-129, 12 halvings here */
                x = rnd(&rng); ind0 = 0; level0 = 49151;
                for (int i = 0; i < 12; i++) {
                    ind = (ind0 + level0) / 2;
                    if (B->hpbgp[ind] > x) level0 = ind; else ind0 = ind;
                }
                for (ind = ind0; ind <= level0; ind++) if (B->hpbgp[ind] >= x) break;
            }
            photons = B->hpbg[ind];
            orc_pix2ang_ring(64, ind, &phi, &theta);
            dir.x = +sinf(theta) * cosf(phi); dir.y = +sinf(theta) * sinf(phi); dir.z = -cosf(theta);
            dir_fix(&dir);
            /* entry point: a disc of radius Rout facing the Healpix pixel, :149-217 */
            ds = 2.0f * PI_F * rnd(&rng);
            dx = sqrtf(rnd(&rng));
            pos.x = dx * cosf(ds); pos.y = dx * sinf(ds); pos.z = sqrtf(1.001f - dx * dx);
            pos0.x = pos.x * cosf(theta) + pos.z * sinf(theta);
            pos0.y = pos.y;
            pos0.z = -pos.x * sinf(theta) + pos.z * cosf(theta);
            pos.x = pos0.x * cosf(PI_F - phi) + pos0.y * sinf(PI_F - phi);
            pos.y = -pos0.x * sinf(PI_F - phi) + pos0.y * cosf(PI_F - phi);
            pos.z = pos0.z;
            pos.x = 0.5f * NX + Rout * pos.x; pos.y = 0.5f * NY + Rout * pos.y; pos.z = 0.5f * NZ + Rout * pos.z;
            surface(&S.N, &pos, &dir);
            index_g(&S.N, &pos, &level, &ind);
            S.c.packets++;
            if (ind < 0) continue;
            sca_propagate(&S, &rng, pos, dir, level, ind, photons, 2);
        }
        np += S.c.packets; ns += S.c.steps; nsc += S.c.scatterings; npl += S.c.peels;
    }
    if (C) { C->packets += np; C->steps += ns; C->scatterings += nsc; C->peels += npl; }
}

/* ---- scattered light, emission from the cells: SimRAM_CL, kernel_ASOC_sca.c:1098-1461 ---------------- */
void orc_sca_cl(const OrcParams *P, const OrcGrid *G, const OrcSimBufs *B, const OrcScaBufs *O, int global,
                int packets, int batch_arg, float seed, OrcCounters *C) {
    (void)packets;
    uint64_t np = 0, ns = 0, nsc = 0, npl = 0;
    #pragma omp parallel for schedule(runtime) reduction(+:np,ns,nsc,npl)
    for (int id = 0; id < global; id++) {
        if (id >= P->cells) continue;
        sca_t S; S.P = P; S.B = B; S.O = O; S.N = nav_make(P, G); memset(&S.c, 0, sizeof(S.c));
        const nav_t *N = &S.N;
        rng_t rng; rng_seed(&rng, seed, (uint64_t)id);
        int icell = id - global, iray = 0, batch = -1, ind, level, done = 0;
        float pwei = 1.0f, X0, Y0, Z0;
        while (!done) {
            if (iray >= batch) {
                iray = 0; pwei = 1.0f;
                for (;;) {
                    icell += global;
                    if (icell >= P->cells) { done = 1; break; }
                    if (P->use_emweight > 0) {
                        pwei = B->emwei[icell];
                        if (pwei < 1e-10f || N->dens[icell] <= 0.0f) continue;
                        batch = (int)floorf(pwei);
                        if (batch < 1) { batch = 1; pwei = (float)(1.0 / (pwei + 1.0e-30f)); }
                        else           { pwei = (float)(1.0 / (batch + 1.0e-9f)); }
                    } else {
                        batch = batch_arg;
                        pwei = 1.0f / (batch + 1.0e-9f);
                    }
                    break;
                }
                if (done) break;
            }
            ind = icell; iray++;
            for (level = 0; level < P->levels - 1; level++) {
                ind -= G->lcells[level];
                if (ind < 0) { ind += G->lcells[level]; break; }
            }
            if (level == 0) { X0 = ind % P->nx; Y0 = (ind / P->nx) % P->ny; Z0 = ind / (P->nx * P->ny); }
            else { int sid = ind % 8; X0 = sid % 2; Y0 = ((sid % 4) > 1) ? 1.0f : 0.0f; Z0 = sid / 4; }
            float photons = B->emit[N->off[level] + ind] * pwei;
            v3 pos, dir;
            pos.x = X0 + rnd(&rng); pos.y = Y0 + rnd(&rng); pos.z = Z0 + rnd(&rng);
            float phi = TWOPI * rnd(&rng);
            float cos_theta = 0.999997f - 1.999995f * rnd(&rng);
            float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
            dir.x = sin_theta * cosf(phi); dir.y = sin_theta * sinf(phi); dir.z = cos_theta;
            dir_fix(&dir);
            S.c.packets++;
            sca_propagate(&S, &rng, pos, dir, level, ind, photons, 3);
        }
        np += S.c.packets; ns += S.c.steps; nsc += S.c.scatterings; npl += S.c.peels;
    }
    if (C) { C->packets += np; C->steps += ns; C->scatterings += nsc; C->peels += npl; }
}
