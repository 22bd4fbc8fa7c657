// soc_b200 -- photon-packet emission shared by the absorption kernels (sim.cu) and the scattered-light
// kernels (sca.cu): point sources and the isotropic background.  ARGS is SimArgs or ScaArgs (same field names).
#pragma once
#include "common.cuh"

struct Packet {
    vec3 pos, dir;
    float photons, free_path, tau, rho;
    int level, ind, scat, eidx, nstep;
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

