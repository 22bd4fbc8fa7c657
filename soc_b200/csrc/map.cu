// soc_b200 -- emission-map ray tracer.
//
// Replaces the reference kernels Mapping (kernel_ASOC_map.c:496-875, MAP_INTERPOLATION==0, no ROI) and
// HealpixMapping (kernel_ASOC_map.c:890-966).  One thread integrates one line of sight away from the
// observer:  I += exp(-tau) * (1-exp(-dtau))/dtau * s * emit * n.  The ray set-up and the cell stepping keep
// the map kernel's own constants (EPS 2.5e-4, PEPS 5e-4, direction clamp 1e-5) and its order of single
// precision operations, so that maps agree with the reference to rounding (contract: 1e-5 relative).
// Per pixel-step: DENS[cell] + EMIT[cell] gathers (8 B; +8 B with per-cell opacities).
#include "map.cuh"

namespace {

__device__ __forceinline__ void clamp_dir(vec3 &d) {
    if (fabsf(d.x) < 1.0e-5f) d.x = 1.0e-5f;
    if (fabsf(d.y) < 1.0e-5f) d.y = 1.0e-5f;
    if (fabsf(d.z) < 1.0e-5f) d.z = 1.0e-5f;
}
__device__ __forceinline__ bool outside_closed(const vec3 &p, float NX, float NY, float NZ) {
    return p.x < 0.0f || p.x > NX || p.y < 0.0f || p.y > NY || p.z < 0.0f || p.z > NZ;
}
__device__ __forceinline__ vec3 back(const vec3 &p, float s, const vec3 &d) {      // p - s*d, no contraction
    vec3 r = { xsub(p.x, xmul(s, d.x)), xsub(p.y, xmul(s, d.y)), xsub(p.z, xmul(s, d.z)) };
    return r;
}

// exp() correctly rounded to float: 1-exp(-dtau) cancels for dtau just above 1e-3, where one ulp of expf is
// 6e-5 of the term -- beyond the 1e-5 contract.  The map kernel has only NPIX rays, so fp64 here is free.
__device__ __forceinline__ float exp_cr(float x) { return (float)exp((double)x); }

// the integration loop shared by both kernels
template <bool OCT, bool DBL, bool HEALPIX>
__device__ __forceinline__ void integrate(const MapArgs &M, vec3 POS, const vec3 &TMP, int id, unsigned long long &steps) {
    const GridDesc &G = M.G;
    int level = 0, ind;
    float rho = 0.0f, TAU = 0.0f, PHOTONS = 0.0f, colden = 0.0f;
    index_global<OCT, true>(G, POS, level, ind, rho);
    // MAP_INTERPOLATION (kernel_ASOC_map.c:656-684): two unit vectors perpendicular to the line of sight
    const int MI = HEALPIX ? 0 : M.map_interpolation;
    vec3 ADIR = { 0.0f, 0.0f, 0.0f }, BDIR = { 0.0f, 0.0f, 0.0f };
    if (MI > 0) {
        if (fabsf(TMP.x) > fabsf(TMP.y)) {
            if (fabsf(TMP.z) > fabsf(TMP.x)) { ADIR.x = 0.0005f; ADIR.y = 1.0f; ADIR.z = xdiv(-TMP.y, TMP.z); }
            else                             { ADIR.x = xdiv(-TMP.z, TMP.x); ADIR.y = 0.0005f; ADIR.z = 1.0f; }
        } else {
            if (fabsf(TMP.z) > fabsf(TMP.y)) { ADIR.x = 0.0005f; ADIR.y = 1.0f; ADIR.z = xdiv(-TMP.y, TMP.z); }
            else                             { ADIR.x = 1.0f; ADIR.y = xdiv(-TMP.x, TMP.y); ADIR.z = 0.0005f; }
        }
        ADIR = normalize3(ADIR);
        BDIR.x = xsub(xmul(TMP.y, ADIR.z), xmul(TMP.z, ADIR.y));
        BDIR.y = xsub(xmul(TMP.z, ADIR.x), xmul(TMP.x, ADIR.z));
        BDIR.z = xsub(xmul(TMP.x, ADIR.y), xmul(TMP.y, ADIR.x));
        BDIR = normalize3(BDIR);
    }
    while (ind >= 0) {
        int oind = OCT ? G.off[level] + ind : ind, olevel = level;
        float dens = rho;
        float em = M.emit[oind];
        float kext;
        if (M.with_abu) { float2 o = reinterpret_cast<const float2 *>(M.opt)[oind]; kext = xadd(o.x, o.y); }
        else            kext = xadd(M.ksca, M.kabs);
        const vec3 POS0 = POS;
        const int ind0 = ind, level0 = level;
        float sx = get_step<OCT, DBL, true>(G, POS, TMP, level, ind, rho);
        if (MI > 0) {                                                          // kernel_ASOC_map.c:706-805
            const float K = ldexpf(1.0f, -level0);
            const float lim = (MI == 2) ? 0.52f : 0.502f;
            if (MI == 2) {
                const float amax = xmul(0.22f, K);
                if (sx > amax) {                                               // the step is cut to 0.22 cells
                    sx = amax;
                    POS.x = xadd(POS0.x, xmul(0.22f, TMP.x)); POS.y = xadd(POS0.y, xmul(0.22f, TMP.y)); POS.z = xadd(POS0.z, xmul(0.22f, TMP.z));
                    ind = ind0; level = level0;
                    if (OCT) { if (DBL) index_octree<double, true>(G, POS, level, ind, rho); else index_octree<float, true>(G, POS, level, ind, rho); }
                    else {
                        if (!(POS.x > 0.0f && POS.x < G.nx && POS.y > 0.0f && POS.y < G.ny && POS.z > 0.0f && POS.z < G.nz)) ind = -1;
                        else { ind = (int)floorf(POS.z) * G.nx * G.ny + (int)floorf(POS.y) * G.nx + (int)floorf(POS.x); rho = G.dens[ind]; }
                    }
                }
            }
            const float h = xdiv(xmul(0.5f, sx), K);
            float a, b, Adens, Bdens, Aemit, Bemit, srho = 0.0f;
            int slevel = level0, sind = ind0;
            vec3 MPOS = { xadd(POS0.x, xmul(h, TMP.x)), xadd(POS0.y, xmul(h, TMP.y)), xadd(POS0.z, xmul(h, TMP.z)) };
            a = xdiv(get_step<OCT, DBL, true>(G, MPOS, ADIR, slevel, sind, srho), K);
            if (a <= lim && sind >= 0) { Adens = srho; Aemit = M.emit[(OCT ? G.off[slevel] : 0) + sind]; }
            else {
                slevel = level0; sind = ind0; ADIR.x = -ADIR.x; ADIR.y = -ADIR.y; ADIR.z = -ADIR.z;
                MPOS.x = xadd(POS0.x, xmul(h, TMP.x)); MPOS.y = xadd(POS0.y, xmul(h, TMP.y)); MPOS.z = xadd(POS0.z, xmul(h, TMP.z));
                a = xdiv(get_step<OCT, DBL, true>(G, MPOS, ADIR, slevel, sind, srho), K);
                if (a <= lim && sind >= 0) { Adens = srho; Aemit = M.emit[(OCT ? G.off[slevel] : 0) + sind]; }
                else { a = 0.5f; Adens = 0.0f; Aemit = 0.0f; }
            }
            slevel = level0; sind = ind0;
            MPOS.x = xadd(POS0.x, xmul(h, TMP.x)); MPOS.y = xadd(POS0.y, xmul(h, TMP.y)); MPOS.z = xadd(POS0.z, xmul(h, TMP.z));
            b = xdiv(get_step<OCT, DBL, true>(G, MPOS, BDIR, slevel, sind, srho), K);
            if (b <= lim && sind >= 0) { Bdens = srho; Bemit = M.emit[(OCT ? G.off[slevel] : 0) + sind]; }
            else {
                slevel = level0; sind = ind0; BDIR.x = -BDIR.x; BDIR.y = -BDIR.y; BDIR.z = -BDIR.z;
                MPOS.x = xadd(POS0.x, xmul(h, TMP.x)); MPOS.y = xadd(POS0.y, xmul(h, TMP.y)); MPOS.z = xadd(POS0.z, xmul(h, TMP.z));
                b = get_step<OCT, DBL, true>(G, MPOS, BDIR, slevel, sind, srho);
                if (MI == 1) b = xdiv(b, K);                                   // sic: not converted in the MAP_INTERPOLATION==2 branch (:750)
                if (b <= lim && sind >= 0) { Bdens = srho; Bemit = M.emit[(OCT ? G.off[slevel] : 0) + sind]; }
                else { b = 0.5f; Bdens = 0.0f; Bemit = 0.0f; }
            }
            if (MI == 2) {
                a = clampf(a, 0.0f, 0.51f); b = clampf(b, 0.0f, 0.51f);
                dens = xadd(xadd(xmul(xsub(0.5f, a), Adens), xmul(xsub(0.5f, b), Bdens)), xmul(xadd(a, b), dens));
                em   = xadd(xadd(xmul(xsub(0.5f, a), Aemit), xmul(xsub(0.5f, b), Bemit)), xmul(xadd(a, b), em));
            } else {
                a = xsub(0.5f, a); b = xsub(0.5f, b);
                dens = xadd(xadd(xmul(xsub(xsub(1.0f, a), b), dens), xmul(a, Adens)), xmul(b, Bdens));
                em   = xadd(xadd(xmul(xsub(xsub(1.0f, a), b), em), xmul(a, Aemit)), xmul(b, Bemit));
            }
        }
        float DTAU = xmul(xmul(sx, dens), kext);
        if ((HEALPIX || M.level_threshold <= 0 || olevel >= M.level_threshold) &&
            (!(M.roi.flags & 4) || in_roi<OCT>(G, M.roi, olevel, ind0))) {                 // ROI_MAP: kernel_ASOC_map.c:37-56, 822, 948
            float w = (DTAU < 1.0e-3f) ? xsub(1.0f, xmul(0.5f, DTAU)) : xdiv(xsub(1.0f, exp_cr(-DTAU)), DTAU);
            PHOTONS = xadd(PHOTONS, xmul(xmul(xmul(xmul(exp_cr(-TAU), w), sx), em), dens));
        }
        TAU = xadd(TAU, DTAU);
        if (HEALPIX || M.save_colden > 0) colden = xadd(colden, xmul(sx, dens));
        steps++;
    }
    M.map[id] = PHOTONS;
    M.savetau[id] = (M.save_colden > 0) ? xmul(colden, M.length) : TAU;
}

template <bool OCT, bool DBL>
__global__ void __launch_bounds__(128) mapping_kernel(const __grid_constant__ MapArgs M) {
    const GridDesc &G = M.G;
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long steps = 0;
    if (id < M.npx * M.npy) {
        const float NX = (float)G.nx, NY = (float)G.ny, NZ = (float)G.nz;
        const int i = id % M.npx, j = id / M.npx;
        vec3 POS, TMP;
        const vec3 DIR = M.dir;
        if (M.intobs.x > -1e10f) {                                              // kernel_ASOC_map.c:534-553
            float phi = xdiv(xmul(SOC_MAP_TWOPI, (float)i), (float)M.npx);
            phi = xadd(phi, SOC_MAP_PI);
            float pix = xdiv(SOC_MAP_TWOPI, (float)M.npx);
            float theta = xmul(pix, (float)(j - (M.npy - 1) / 2));
            POS = M.intobs;
            float st, ct, sp, cp;
            sincosf(theta, &st, &ct); sincosf(phi, &sp, &cp);
            TMP.x = xmul(ct, cp); TMP.y = xmul(ct, sp); TMP.z = st;
            clamp_dir(TMP);
            if (fmod1(POS.x) < 1.0e-5f) POS.x = xadd(POS.x, 2.0e-5f);
            if (fmod1(POS.y) < 1.0e-5f) POS.y = xadd(POS.y, 2.0e-5f);
            if (fmod1(POS.z) < 1.0e-5f) POS.z = xadd(POS.z, 2.0e-5f);
        } else {                                                                // kernel_ASOC_map.c:554-640
            float fi = xmul(xsub((float)i, xmul(0.5f, (float)(M.npx - 1))), M.map_dx);
            float fj = xmul(xsub((float)j, xmul(0.5f, (float)(M.npy - 1))), M.map_dx);
            POS.x = xadd(xadd(M.centre.x, xmul(fi, M.ra.x)), xmul(fj, M.de.x));
            POS.y = xadd(xadd(M.centre.y, xmul(fi, M.ra.y)), xmul(fj, M.de.y));
            POS.z = xadd(xadd(M.centre.z, xmul(fi, M.ra.z)), xmul(fj, M.de.z));
            float far_ = (float)(G.nx + G.ny + G.nz);
            POS.x = xadd(POS.x, xmul(far_, DIR.x)); POS.y = xadd(POS.y, xmul(far_, DIR.y)); POS.z = xadd(POS.z, xmul(far_, DIR.z));
            float sx = xdiv((DIR.x >= 0.0f) ? xsub(NX, POS.x) : xsub(0.0f, POS.x), -DIR.x);
            float sy = xdiv((DIR.y >= 0.0f) ? xsub(NY, POS.y) : xsub(0.0f, POS.y), -DIR.y);
            float sz = xdiv((DIR.z >= 0.0f) ? xsub(NZ, POS.z) : xsub(0.0f, POS.z), -DIR.z);
            if (G.nx < 200) {
                sx = xadd(sx, SOC_MAP_EPS); sy = xadd(sy, SOC_MAP_EPS); sz = xadd(sz, SOC_MAP_EPS);
                if (outside_closed(back(POS, sx, DIR), NX, NY, NZ)) sx = 1e10f;
                if (outside_closed(back(POS, sy, DIR), NX, NY, NZ)) sy = 1e10f;
                if (outside_closed(back(POS, sz, DIR), NX, NY, NZ)) sz = 1e10f;
                sx = fminf(sx, fminf(sy, sz));
                POS = back(POS, sx, DIR);
            } else {
                const float ex = (DIR.x > 0.0f) ? -SOC_MAP_EPS : SOC_MAP_EPS, ey = (DIR.y > 0.0f) ? -SOC_MAP_EPS : SOC_MAP_EPS,
                            ez = (DIR.z > 0.0f) ? -SOC_MAP_EPS : SOC_MAP_EPS;
                vec3 t;
                t = back(POS, sx, DIR); t.x = xadd(t.x, ex); t.y = xadd(t.y, ey); t.z = xadd(t.z, ez);
                if (outside_closed(t, NX, NY, NZ)) sx = 1e10f;
                t = back(POS, sy, DIR); t.x = xadd(t.x, ex); t.y = xadd(t.y, ey); t.z = xadd(t.z, ez);
                if (outside_closed(t, NX, NY, NZ)) sy = 1e10f;
                t = back(POS, sz, DIR); t.x = xadd(t.x, ex); t.y = xadd(t.y, ey); t.z = xadd(t.z, ez);
                if (outside_closed(t, NX, NY, NZ)) sz = 1e10f;
                sx = fminf(sx, fminf(sy, sz));
                POS = back(POS, sx, DIR);
                POS.x = xadd(POS.x, ex); POS.y = xadd(POS.y, ey); POS.z = xadd(POS.z, ez);
            }
            TMP.x = -DIR.x; TMP.y = -DIR.y; TMP.z = -DIR.z;
            clamp_dir(TMP);
        }
        integrate<OCT, DBL, false>(M, POS, TMP, id, steps);
    }
    warp_add_counter(M.counters + 1, steps);
}

template <bool OCT, bool DBL>
__global__ void __launch_bounds__(128) healpix_mapping_kernel(const __grid_constant__ MapArgs M) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long steps = 0;
    if (id < 12 * M.nside * M.nside) {
        float phi, theta, st, ct, sp, cp;
        pix2ang_ring(M.nside, id, phi, theta, SOC_MAP_PI);
        sincosf(theta, &st, &ct); sincosf(phi, &sp, &cp);
        vec3 TMP = { -xmul(st, cp), -xmul(st, sp), ct };
        clamp_dir(TMP);
        vec3 POS = M.intobs;
        // sic (kernel_ASOC_map.c:929-931): the second test makes the offset practically unconditional
        if (fmod1(POS.x) < 1.0e-5f || fmod1(POS.x) < 0.99999f) POS.x = xadd(POS.x, 2.0e-5f);
        if (fmod1(POS.y) < 1.0e-5f || fmod1(POS.y) < 0.99999f) POS.y = xadd(POS.y, 2.0e-5f);
        if (fmod1(POS.z) < 1.0e-5f || fmod1(POS.z) < 0.99999f) POS.z = xadd(POS.z, 2.0e-5f);
        integrate<OCT, DBL, true>(M, POS, TMP, id, steps);
    }
    warp_add_counter(M.counters + 1, steps);
}

// Per-level Mapping (kernel_ASOC_map_H.c:380-506; ASOC.py:3320-3440, `mapping nx ny dx 999`): MAP[l*npix + id] =
// emission of the level-l cells along the line of sight behind all the material in front of them.  The ray set-up is
// that file's own (start point = last face crossing found from behind the cloud, no clamp of the direction components,
// its own perspective convention); stepping is the map kernel's Index -- the copy in kernel_ASOC_map_H.c:250 forgets to
// store the root coordinates when a ray climbs into a root-grid leaf and loses ~15 % of the flux of a refined cloud
// (DESIGN.md section 7), which is not reproduced.  Level sums stay in registers: the adds are predicated, not indexed.
#define SOC_MAP_MAXLEV 12
template <bool OCT, bool DBL, bool LIT>
__global__ void __launch_bounds__(128) mapping_levels_kernel(const __grid_constant__ MapArgs M) {
    const GridDesc &G = M.G;
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long steps = 0;
    if (id < M.npx * M.npy) {
        const float NX = (float)G.nx, NY = (float)G.ny, NZ = (float)G.nz;
        const int i = id % M.npx, j = id / M.npx;
        vec3 POS, TMP;
        const vec3 DIR = M.dir;
        if (M.intobs.x > -1e10f) {                                              // :403-430
            float phi = xdiv(xmul(SOC_MAP_TWOPI, (float)i), (float)M.npx);
            phi = xadd(phi, SOC_MAP_PI);
            float pix = xdiv(SOC_MAP_TWOPI, (float)M.npx);
            float theta = xmul(pix, (float)(j - (M.npy - 1) / 2));
            POS = M.intobs;
            float st, ct, sp, cp;
            sincosf(theta, &st, &ct); sincosf(phi, &sp, &cp);
            TMP.x = -xmul(ct, sp); TMP.y = -xmul(ct, cp); TMP.z = st;
            clamp_dir(TMP);
            if (fmod1(POS.x) < 1.0e-5f) POS.x = xadd(POS.x, 2.0e-5f);
            if (fmod1(POS.y) < 1.0e-5f) POS.y = xadd(POS.y, 2.0e-5f);
            if (fmod1(POS.z) < 1.0e-5f) POS.z = xadd(POS.z, 2.0e-5f);
        } else {                                                                // :431-457
            float fi = xmul(xsub((float)i, xmul(0.5f, (float)(M.npx - 1))), M.map_dx);
            float fj = xmul(xsub((float)j, xmul(0.5f, (float)(M.npy - 1))), M.map_dx);
            POS.x = xadd(xadd(M.centre.x, xmul(fi, M.ra.x)), xmul(fj, M.de.x));
            POS.y = xadd(xadd(M.centre.y, xmul(fi, M.ra.y)), xmul(fj, M.de.y));
            POS.z = xadd(xadd(M.centre.z, xmul(fi, M.ra.z)), xmul(fj, M.de.z));
            float far_ = (float)(G.nx + G.ny + G.nz);
            POS.x = xsub(POS.x, xmul(far_, DIR.x)); POS.y = xsub(POS.y, xmul(far_, DIR.y)); POS.z = xsub(POS.z, xmul(far_, DIR.z));
            float sx = (DIR.x >= 0.0f) ? xsub(xdiv(xsub(NX, POS.x), xadd(DIR.x, 1.0e-10f)), SOC_MAP_EPS) : xsub(xdiv(xsub(0.0f, POS.x), DIR.x), SOC_MAP_EPS);
            float sy = (DIR.y >= 0.0f) ? xsub(xdiv(xsub(NY, POS.y), xadd(DIR.y, 1.0e-10f)), SOC_MAP_EPS) : xsub(xdiv(xsub(0.0f, POS.y), DIR.y), SOC_MAP_EPS);
            float sz = (DIR.z >= 0.0f) ? xsub(xdiv(xsub(NZ, POS.z), xadd(DIR.z, 1.0e-10f)), SOC_MAP_EPS) : xsub(xdiv(xsub(0.0f, POS.z), DIR.z), SOC_MAP_EPS);
            vec3 t;
            t = back(POS, -sx, DIR); if (t.x <= 0.0f || t.x >= NX || t.y <= 0.0f || t.y >= NY || t.z <= 0.0f || t.z >= NZ) sx = -1e10f;
            t = back(POS, -sy, DIR); if (t.x <= 0.0f || t.x >= NX || t.y <= 0.0f || t.y >= NY || t.z <= 0.0f || t.z >= NZ) sy = -1e10f;
            t = back(POS, -sz, DIR); if (t.x <= 0.0f || t.x >= NX || t.y <= 0.0f || t.y >= NY || t.z <= 0.0f || t.z >= NZ) sz = -1e10f;
            sx = fmaxf(sx, fmaxf(sy, sz));
            POS = back(POS, -sx, DIR);
            TMP.x = -DIR.x; TMP.y = -DIR.y; TMP.z = -DIR.z;
        }
        int level = 0, ind;
        float rho = 0.0f, TAU = 0.0f, colden = 0.0f;
        float PH[SOC_MAP_MAXLEV];
        #pragma unroll
        for (int l = 0; l < SOC_MAP_MAXLEV; l++) PH[l] = 0.0f;
        index_global<OCT, true>(G, POS, level, ind, rho);
        while (ind >= 0) {                                                      // :468-488
            const int oind = OCT ? G.off[level] + ind : ind, olevel = level;
            const float dens = rho, em = M.emit[oind];
            float kext;
            if (M.with_abu) { float2 o = reinterpret_cast<const float2 *>(M.opt)[oind]; kext = xadd(o.x, o.y); }
            else            kext = xadd(M.ksca, M.kabs);
            const float sx = get_step<OCT, DBL, true, LIT>(G, POS, TMP, level, ind, rho);
            const float DTAU = xmul(xmul(sx, dens), kext);
            const float w = (DTAU < 1.0e-3f) ? xsub(1.0f, xmul(0.5f, DTAU)) : xdiv(xsub(1.0f, exp_cr(-DTAU)), DTAU);
            const float term = xmul(xmul(xmul(xmul(exp_cr(-TAU), w), sx), em), dens);
            #pragma unroll
            for (int l = 0; l < SOC_MAP_MAXLEV; l++) if (l == olevel) PH[l] = xadd(PH[l], term);
            TAU = xadd(TAU, DTAU);
            colden = xadd(colden, xmul(sx, dens));
            steps++;
        }
        const long long npix = (long long)M.npx * M.npy;
        #pragma unroll
        for (int l = 0; l < SOC_MAP_MAXLEV; l++) if (l < G.levels) M.map[l * npix + id] = PH[l];
        if (M.savetau != nullptr) M.savetau[id] = xmul(colden, M.length);
    }
    warp_add_counter(M.counters + 1, steps);
}

// PSTau (kernel_ASOC_map.c:1545-1599): tau and column density from each point source towards the observer
template <bool OCT, bool DBL>
__global__ void pstau_kernel(const __grid_constant__ MapArgs M, int no, const float *__restrict__ pspos, float *__restrict__ colden_out,
                             float *__restrict__ tau_out) {
    const GridDesc &G = M.G;
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= no) return;
    vec3 POS = { pspos[3 * id], pspos[3 * id + 1], pspos[3 * id + 2] };
    int level = 0, ind;
    float rho = 0.0f, TAU = 0.0f, colden = 0.0f;
    index_global<OCT, true>(G, POS, level, ind, rho);
    while (ind >= 0) {
        const int oind = OCT ? G.off[level] + ind : ind;
        const float dens = rho;
        float kext;
        if (M.with_abu) { float2 o = reinterpret_cast<const float2 *>(M.opt)[oind]; kext = xadd(o.x, o.y); }
        else            kext = xadd(M.ksca, M.kabs);
        const float sx = get_step<OCT, DBL, true>(G, POS, M.dir, level, ind, rho);
        TAU = xadd(TAU, xmul(xmul(sx, dens), kext));
        colden = xadd(colden, xmul(sx, dens));
    }
    colden_out[id] = xmul(colden, M.length);
    tau_out[id] = TAU;
}

}  // namespace

void launch_pstau(const MapArgs &M, int no, const float *pspos, float *colden, float *tau, cudaStream_t stream) {
    const bool oct = M.G.levels > 1, dbl = M.G.dbl_map != 0;
    const int threads = 32, blocks = (no + threads - 1) / threads;
    if (!oct)      pstau_kernel<false, false><<<blocks, threads, 0, stream>>>(M, no, pspos, colden, tau);
    else if (!dbl) pstau_kernel<true, false><<<blocks, threads, 0, stream>>>(M, no, pspos, colden, tau);
    else           pstau_kernel<true, true><<<blocks, threads, 0, stream>>>(M, no, pspos, colden, tau);
}

void launch_mapping_levels(const MapArgs &M, cudaStream_t stream) {
    const bool oct = M.G.levels > 1, dbl = M.G.dbl_map != 0;
    const int n = M.npx * M.npy, threads = 128, blocks = (n + threads - 1) / threads;
    if (!oct)      mapping_levels_kernel<false, false, false><<<blocks, threads, 0, stream>>>(M);
    else if (M.maph_literal) {
        if (!dbl) mapping_levels_kernel<true, false, true><<<blocks, threads, 0, stream>>>(M);
        else      mapping_levels_kernel<true, true, true><<<blocks, threads, 0, stream>>>(M);
    } else {
        if (!dbl) mapping_levels_kernel<true, false, false><<<blocks, threads, 0, stream>>>(M);
        else      mapping_levels_kernel<true, true, false><<<blocks, threads, 0, stream>>>(M);
    }
}

void launch_mapping(const MapArgs &M, bool healpix, cudaStream_t stream) {
    const bool oct = M.G.levels > 1, dbl = M.G.dbl_map != 0;
    const int n = healpix ? 12 * M.nside * M.nside : M.npx * M.npy;
    const int threads = 128, blocks = (n + threads - 1) / threads;
    if (healpix) {
        if (!oct)      healpix_mapping_kernel<false, false><<<blocks, threads, 0, stream>>>(M);
        else if (!dbl) healpix_mapping_kernel<true, false><<<blocks, threads, 0, stream>>>(M);
        else           healpix_mapping_kernel<true, true><<<blocks, threads, 0, stream>>>(M);
    } else {
        if (!oct)      mapping_kernel<false, false><<<blocks, threads, 0, stream>>>(M);
        else if (!dbl) mapping_kernel<true, false><<<blocks, threads, 0, stream>>>(M);
        else           mapping_kernel<true, true><<<blocks, threads, 0, stream>>>(M);
    }
}
