// soc_b200 -- photon-packet emission shared by the absorption kernels (sim.cu) and the scattered-light
// kernels (sca.cu): point sources and the isotropic background.  ARGS is SimArgs or ScaArgs (same field names).
#pragma once
#include "common.cuh"

struct Packet {
    vec3 pos, dir;
    float photons, free_path, tau, rho;
    int level, ind, scat, eidx, nstep;
    int roi;            // WITH_ROI_SAVE: the packet was inside ROI after its last full step
};


template <bool OCT>
__device__ __forceinline__ void locate(const GridDesc &G, Packet &pk) {
    index_global<OCT, false>(G, pk.pos, pk.level, pk.ind, pk.rho);
}

// isotropic direction of the reference: kernel_ASOC.c:202-206
template <class RNG>
__device__ __forceinline__ void isotropic(RNG &rng, vec3 &d) {
    float phi = SOC_TWOPI * rng.uniform();
    float cos_theta = 0.999997f - 1.999995f * rng.uniform();
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    float s, c;
    sincosf(phi, &s, &c);
    d.x = sin_theta * c; d.y = sin_theta * s; d.z = cos_theta;
}

// ---- emission: point sources (kernel_ASOC.c:202-434) ---------------------------------------------------------
template <class ARGS, class RNG, bool OCT>
__device__ void emit_ps(const ARGS &A, RNG &rng, int III, Packet &pk) {
    const GridDesc &G = A.G;
    const float NX = (float)G.nx, NY = (float)G.ny, NZ = (float)G.nz;
    isotropic(rng, pk.dir);
    int ips = III % A.no_ps;
    pk.photons = A.ps[ips];
    vec3 src = { A.pspos[3 * ips], A.pspos[3 * ips + 1], A.pspos[3 * ips + 2] };
    pk.pos = src;
    locate<OCT>(G, pk);
    if (pk.ind >= 0) return;
    // source outside the cloud
    float v1, v2, cos_theta;
    switch (A.ps_method) {
    case 0:
        to_surface(G, pk.pos, pk.dir); locate<OCT>(G, pk); break;
    case 1:
        pk.pos = src;
        if (pk.pos.z > NZ) { if (pk.dir.z > 0.0f) pk.dir.z = -pk.dir.z; }
        else if (pk.pos.z < 0.0f) { if (pk.dir.z < 0.0f) pk.dir.z = -pk.dir.z; }
        else if (pk.pos.x > NX) { if (pk.dir.x > 0.0f) pk.dir.x = -pk.dir.x; }
        else if (pk.pos.x < 0.0f) { if (pk.dir.x < 0.0f) pk.dir.x = -pk.dir.x; }
        else if (pk.pos.y > NY) { if (pk.dir.y > 0.0f) pk.dir.y = -pk.dir.y; }
        else if (pk.pos.y < 0.0f) { if (pk.dir.y < 0.0f) pk.dir.y = -pk.dir.y; }
        to_surface(G, pk.pos, pk.dir);
        pk.photons *= 0.5f;
        locate<OCT>(G, pk); break;
    case 2: {
        pk.pos = src;
        int k = (int)floorf(rng.uniform() * A.xps_nside[ips] * 0.999999f);
        pk.photons /= A.xps_area[3 * ips + k];
        k = A.xps_side[3 * ips + k];
        float a = rng.uniform(), b = rng.uniform();
        if (k == 0) { pk.pos.x = NX - SOC_PEPS; pk.pos.y = a * NY; pk.pos.z = b * NZ; b = NY * NZ; }
        if (k == 1) { pk.pos.x = SOC_PEPS;      pk.pos.y = a * NY; pk.pos.z = b * NZ; b = NY * NZ; }
        if (k == 2) { pk.pos.y = NY - SOC_PEPS; pk.pos.x = a * NX; pk.pos.z = b * NZ; b = NX * NZ; }
        if (k == 3) { pk.pos.y = SOC_PEPS;      pk.pos.x = a * NX; pk.pos.z = b * NZ; b = NX * NZ; }
        if (k == 4) { pk.pos.z = NZ - SOC_PEPS; pk.pos.x = a * NX; pk.pos.y = b * NY; b = NX * NY; }
        if (k == 5) { pk.pos.z = SOC_PEPS;      pk.pos.x = a * NX; pk.pos.y = b * NY; b = NX * NY; }
        vec3 dd = { xsub(pk.pos.x, src.x), xsub(pk.pos.y, src.y), xsub(pk.pos.z, src.z) };
        v1 = sqrtf(xadd(xadd(xmul(dd.x, dd.x), xmul(dd.y, dd.y)), xmul(dd.z, dd.z)));
        pk.dir = normalize3(dd);
        v2 = (k < 2) ? fabsf(pk.dir.x) : ((k < 4) ? fabsf(pk.dir.y) : fabsf(pk.dir.z));
        pk.photons *= v2 * b / (4.0f * SOC_PI * v1 * v1);
        locate<OCT>(G, pk); break; }
    case 4: {
        v1 = src.z - NZ;
        cos_theta = v1 / sqrtf(v1 * v1 + 0.25f * NX * NX + 0.25f * NY * NY);
        pk.photons *= 0.5f * (1.0f - cos_theta);
        cos_theta = 1.0f - rng.uniform() * (1.0f - cos_theta);
        v1 = SOC_TWOPI * rng.uniform();
        float s, c; sincosf(v1, &s, &c);
        pk.dir.x = sqrtf(1.0f - cos_theta * cos_theta) * c;
        pk.dir.y = sqrtf(1.0f - cos_theta * cos_theta) * s;
        pk.dir.z = -cos_theta;
        to_surface(G, pk.pos, pk.dir); locate<OCT>(G, pk); break; }
    case 5: {
        cos_theta = A.xps_area[3 * ips];
        pk.photons *= 0.5f * (1.0f - cos_theta);
        cos_theta = 1.0f - rng.uniform() * (1.0f - cos_theta);
        v1 = SOC_TWOPI * rng.uniform();
        int sd = A.xps_side[3 * ips];
        float st = sqrtf(1.0f - cos_theta * cos_theta), s, c;
        sincosf(v1, &s, &c);
        if (sd < 2)      { pk.dir.y = st * c; pk.dir.z = st * s; pk.dir.x = (sd == 0) ? -cos_theta : cos_theta; }
        else if (sd < 4) { pk.dir.x = st * c; pk.dir.z = st * s; pk.dir.y = (sd == 2) ? -cos_theta : cos_theta; }
        else             { pk.dir.x = st * c; pk.dir.y = st * s; pk.dir.z = (sd == 4) ? -cos_theta : cos_theta; }
        to_surface(G, pk.pos, pk.dir); locate<OCT>(G, pk); break; }
    default: break;
    }
}

// ---- emission: isotropic background from surface element id % AREA (kernel_ASOC.c:109-138, 439-464) -----------
template <class ARGS, class RNG, bool OCT>
__device__ void emit_bg(const ARGS &A, RNG &rng, int id, Packet &pk) {
    const GridDesc &G = A.G;
    const int NX = G.nx, NY = G.ny, NZ = G.nz;
    int e = id % G.area, side;
    float X0 = 0.0f, Y0 = 0.0f, Z0 = 0.0f, DX = 1.0f, DY = 1.0f, DZ = 1.0f;
    if (e < NY * NZ) { side = 0; X0 = SOC_PEPS; Y0 = e % NY; Z0 = e / NY; DX = 0.0f; }
    else { e -= NY * NZ;
    if (e < NY * NZ) { side = 1; X0 = NX - SOC_PEPS; Y0 = e % NY; Z0 = e / NY; DX = 0.0f; }
    else { e -= NY * NZ;
    if (e < NX * NZ) { side = 2; Y0 = SOC_PEPS; X0 = e % NX; Z0 = e / NX; DY = 0.0f; }
    else { e -= NX * NZ;
    if (e < NX * NZ) { side = 3; Y0 = NY - SOC_PEPS; X0 = e % NX; Z0 = e / NX; DY = 0.0f; }
    else { e -= NX * NZ;
    if (e < NX * NY) { side = 4; Z0 = SOC_PEPS; X0 = e % NX; Y0 = e / NX; DZ = 0.0f; }
    else { e -= NX * NY; side = 5; Z0 = NZ - SOC_PEPS; X0 = e % NX; Y0 = e / NX; DZ = 0.0f; } } } } }
    pk.pos.x = clampf(xadd(X0, xmul(DX, rng.uniform())), SOC_PEPS, NX - SOC_PEPS);
    pk.pos.y = clampf(xadd(Y0, xmul(DY, rng.uniform())), SOC_PEPS, NY - SOC_PEPS);
    pk.pos.z = clampf(xadd(Z0, xmul(DZ, rng.uniform())), SOC_PEPS, NZ - SOC_PEPS);
    float cos_theta = sqrtf(rng.uniform());
    float phi = SOC_TWOPI * rng.uniform();
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    float s, c;
    sincosf(phi, &s, &c);
    float v1 = sin_theta * c, v2 = sin_theta * s;
    switch (side) {
    case 0: pk.dir.x =  cos_theta; pk.dir.y = v1; pk.dir.z = v2; break;
    case 1: pk.dir.x = -cos_theta; pk.dir.y = v1; pk.dir.z = v2; break;
    case 2: pk.dir.y =  cos_theta; pk.dir.x = v1; pk.dir.z = v2; break;
    case 3: pk.dir.y = -cos_theta; pk.dir.x = v1; pk.dir.z = v2; break;
    case 4: pk.dir.z =  cos_theta; pk.dir.x = v1; pk.dir.y = v2; break;
    default: pk.dir.z = -cos_theta; pk.dir.x = v1; pk.dir.y = v2; break;
    }
    pk.photons = A.bg;
    locate<OCT>(G, pk);
}


// ---- emission: Healpix background (kernel_ASOC.c:885-947) ---------------------------------------------------
template <class ARGS, class RNG, bool OCT>
__device__ void emit_hp(const ARGS &A, RNG &rng, Packet &pk) {
    const GridDesc &G = A.G;
    const float NX = (float)G.nx, NY = (float)G.ny, NZ = (float)G.nz;
    int ipix;
    if (A.hpbg_weighted < 1) {
        ipix = clampi((int)floorf(rng.uniform() * 49152), 0, 49151);
    } else {
        float x = rng.uniform();
        int lo = 0, hi = 49151;
        for (int i = 0; i < 10; i++) {
            ipix = (lo + hi) / 2;
            if (A.hpbgp[ipix] > x) hi = ipix; else lo = ipix;
        }
        for (ipix = lo; ipix <= hi; ipix++) if (A.hpbgp[ipix] >= x) break;
    }
    pk.photons = A.hpbg[ipix];
    float phi, theta, st, ct, sp, cp;
    pix2ang_ring(64, ipix, phi, theta, SOC_PI);
    sincosf(theta, &st, &ct); sincosf(phi, &sp, &cp);
    pk.dir.x = st * cp; pk.dir.y = st * sp; pk.dir.z = -ct;
    fix_direction(pk.dir);
    float x = fabsf(pk.dir.x), y = fabsf(pk.dir.y), z = fabsf(pk.dir.z);
    float ds = xadd(xadd(x, y), z);
    x = xdiv(x, ds); y = xdiv(y, ds); z = xdiv(z, ds);
    ds = rng.uniform();
    float v1 = rng.uniform(), v2 = rng.uniform();
    if (ds < x)                { pk.pos.y = v1 * NY; pk.pos.z = v2 * NZ; pk.pos.x = (pk.dir.x > 0.0f) ? SOC_PEPS : (NX - SOC_PEPS); }
    else if (ds < xadd(x, y))  { pk.pos.x = v1 * NX; pk.pos.z = v2 * NZ; pk.pos.y = (pk.dir.y > 0.0f) ? SOC_PEPS : (NY - SOC_PEPS); }
    else                       { pk.pos.x = v1 * NX; pk.pos.y = v2 * NY; pk.pos.z = (pk.dir.z > 0.0f) ? SOC_PEPS : (NZ - SOC_PEPS); }
    locate<OCT>(G, pk);
}

// ---- emission: one ray from cell `icell` (kernel_ASOC.c:1323-1393) ------------------------------------------
template <class ARGS, class RNG>
__device__ void emit_cl(const ARGS &A, RNG &rng, int icell, float pwei, Packet &pk) {
    const GridDesc &G = A.G;
    int ind = icell, level;
    for (level = 0; level < G.levels - 1; level++) {
        ind -= G.lcells[level];
        if (ind < 0) { ind += G.lcells[level]; break; }
    }
    float X0, Y0, Z0;
    if (level == 0) { X0 = ind % G.nx; Y0 = (ind / G.nx) % G.ny; Z0 = ind / (G.nx * G.ny); }
    else { int sid = ind & 7; X0 = sid & 1; Y0 = (sid >> 1) & 1; Z0 = sid >> 2; }
    pk.photons = A.emit[icell] * pwei;
    pk.pos.x = xadd(X0, rng.uniform()); pk.pos.y = xadd(Y0, rng.uniform()); pk.pos.z = xadd(Z0, rng.uniform());
    isotropic(rng, pk.dir);
    pk.level = level; pk.ind = ind; pk.rho = G.dens[icell];
    pk.eidx = A.with_ali ? icell : -1;
}

// Packets-per-cell rule of SimRAM_CL (kernel_ASOC.c:1293-1316).  Returns the number of rays (0 = skip the cell).
template <class ARGS>
__device__ __forceinline__ int cl_rays(const ARGS &A, int icell, float &pwei) {
    if (A.use_emweight > 0) {
        pwei = A.emwei[icell];
        if (pwei < 1e-10f || A.G.dens[icell] <= 0.0f) return 0;
        int batch = (int)floorf(pwei);
        if (batch < 1) { batch = 1; pwei = (float)(1.0 / (double)(pwei + 1.0e-30f)); }
        else           { pwei = (float)(1.0 / (double)((float)batch + 1.0e-9f)); }
        return batch;
    }
    pwei = 1.0f / ((float)A.batch + 1.0e-9f);
    return A.batch;
}


// ---- emission: Healpix background of the scattered-light run (kernel_ASOC_sca.c:104-217) ------------------------
// Unlike the absorption run's SimRAM_HP the packet enters through a disc of radius Rout facing the sky pixel and
// is stepped onto the cloud surface; packets that miss the cloud are dropped (ind < 0).
template <class ARGS, class RNG, bool OCT>
__device__ void emit_hp_sca(const ARGS &A, RNG &rng, Packet &pk) {
    const GridDesc &G = A.G;
    const float NX = (float)G.nx, NY = (float)G.ny, NZ = (float)G.nz;
    const float Rout = xmul(0.5f, sqrtf(xadd(xadd(xmul(xmul(1.0f, NX), NX), xmul(NY, NY)), xmul(NZ, NZ))));
    int ipix;
    if (A.hpbg_weighted < 1) {
        ipix = clampi((int)floorf(rng.uniform() * 49152), 0, 49151);
    } else {
        float x = rng.uniform();
        int lo = 0, hi = 49151;
        for (int i = 0; i < 12; i++) {
            ipix = (lo + hi) / 2;
            if (A.hpbgp[ipix] > x) hi = ipix; else lo = ipix;
        }
        for (ipix = lo; ipix <= hi; ipix++) if (A.hpbgp[ipix] >= x) break;
    }
    pk.photons = A.hpbg[ipix];
    float phi, theta, st, ct, sp, cp;
    pix2ang_ring(64, ipix, phi, theta, SOC_PI);
    sincosf(theta, &st, &ct); sincosf(phi, &sp, &cp);
    pk.dir.x = xmul(st, cp); pk.dir.y = xmul(st, sp); pk.dir.z = -ct;
    fix_direction(pk.dir);
    float ds = xmul(xmul(2.0f, SOC_PI), rng.uniform());
    float dx = sqrtf(rng.uniform());
    float sd, cd;
    sincosf(ds, &sd, &cd);
    vec3 p = { xmul(dx, cd), xmul(dx, sd), sqrtf(xsub(1.001f, xmul(dx, dx))) };
    vec3 q = { xadd(xmul(p.x, ct), xmul(p.z, st)), p.y, xadd(xmul(-p.x, st), xmul(p.z, ct)) };
    float s2, c2;
    sincosf(xsub(SOC_PI, phi), &s2, &c2);
    p.x = xadd(xmul(q.x, c2), xmul(q.y, s2));
    p.y = xadd(xmul(-q.x, s2), xmul(q.y, c2));
    p.z = q.z;
    pk.pos.x = xadd(xmul(0.5f, NX), xmul(Rout, p.x)); pk.pos.y = xadd(xmul(0.5f, NY), xmul(Rout, p.y)); pk.pos.z = xadd(xmul(0.5f, NZ), xmul(Rout, p.z));
    to_surface(G, pk.pos, pk.dir);
    locate<OCT>(G, pk);
}

// ---- emission: external field stored on the surface of the model (SOURCE == 3, WITH_ROI_LOAD) -----------------
// kernel_ASOC.c:141-179, 469-502 (== kernel_ASOC_sca.c SimRAM_PB): work item id serves surface element id % nelem of
// the loaded file (100 work items per element), ray III comes from Healpix pixel III % npix of that element.
// Returns false when the pixel is empty (the reference skips it before drawing any random number).
template <class ARGS, class RNG, bool OCT>
__device__ bool emit_roi(const ARGS &A, RNG &rng, int id, int III, int nelem, Packet &pk) {
    const GridDesc &G = A.G;
    const int *RD = A.roi.dim;
    const int ielem = id % nelem;
    int iside = ielem;
    const float rd = xdiv((float)G.nx, (float)RD[0]);
    float DX = 0.0f, DY = 0.0f;
    if (iside < RD[1] * RD[2]) { DX = xmul((float)(iside % RD[1]) + 0.5f, rd); DY = xmul((float)(iside / RD[1]) + 0.5f, rd); iside = 0; }
    else { iside -= RD[1] * RD[2];
    if (iside < RD[0] * RD[2]) { DX = xmul((float)(iside % RD[0]) + 0.5f, rd); DY = xmul((float)(iside / RD[0]) + 0.5f, rd); iside = 1; }
    else { iside -= RD[0] * RD[2];
    if (iside < RD[0] * RD[1]) { DX = xmul((float)(iside % RD[0]) + 0.5f, rd); DY = xmul((float)(iside / RD[0]) + 0.5f, rd); iside = 2; } } }
    const int npix = 12 * A.roi.nside * A.roi.nside;
    const float X0 = (float)(A.roi.nside * A.roi.nside * 12.0 / (100.0 * A.batch));
    const int ipix = III % npix;
    pk.photons = X0 * A.roi.load[(size_t)ielem * npix + ipix];
    pk.ind = -1;
    if (pk.photons <= 0.0f) return false;
    float v1, v2;
    pix2ang_ring(A.roi.nside, ipix, v1, v2, SOC_PI);
    v1 = xadd(v1, xmul(xsub(rng.uniform(), 0.5f), 0.05f));
    v2 = xadd(v2, xmul(xsub(rng.uniform(), 0.5f), 0.05f));
    float s1, c1, s2, c2;
    sincosf(v1, &s1, &c1); sincosf(v2, &s2, &c2);
    pk.dir.x = xmul(s2, c1); pk.dir.y = xmul(s2, s1); pk.dir.z = c2;
    const float NX = (float)G.nx, NY = (float)G.ny, NZ = (float)G.nz;
    const float a = xadd(DX, xmul(xadd(-0.49f, xmul(0.98f, rng.uniform())), rd));
    const float b = xadd(DY, xmul(xadd(-0.49f, xmul(0.98f, rng.uniform())), rd));
    if (iside == 0)      { pk.pos.y = a; pk.pos.z = b; pk.pos.x = (pk.dir.x > 0.0f) ? SOC_PEPS : (NX - SOC_PEPS); }
    else if (iside == 1) { pk.pos.x = a; pk.pos.z = b; pk.pos.y = (pk.dir.y > 0.0f) ? SOC_PEPS : (NY - SOC_PEPS); }
    else                 { pk.pos.x = a; pk.pos.y = b; pk.pos.z = (pk.dir.z > 0.0f) ? SOC_PEPS : (NZ - SOC_PEPS); }
    locate<OCT>(G, pk);
    return true;
}
