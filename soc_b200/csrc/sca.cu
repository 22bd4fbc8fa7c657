// soc_b200 -- scattered-light kernels with peel-off (ASOCS.py).
//
// Replaces SimRAM_PS (kernel_ASOC_sca.c:1462-1937) and SimRAM_PB (:471-1088) for orthographic observers:
// optional forced first scattering, no absorption bookkeeping (the packet weight carries exp(-tau_abs)), and
// at every scattering one peel-off ray per observer direction to the cloud surface whose attenuated weight
// is added to the image OUT[idir, j, i].
//
// B200 layout: the three kinds of rays a packet needs -- the forced-first-scattering look-ahead, the random
// walk itself and the peel-off rays -- are all "step from cell to cell and sum n*s*kappa", so each lane runs
// a small state machine and every lane of the warp executes the same GetStep per iteration whatever its
// state; only the (rare) transitions diverge.  Per cell-step the kernel reads DENS[cell] (4 B; +8 B with
// per-cell opacities); per peel-off ray it does one red.global.add.f32 into the image.
#include "sca.cuh"
#include "emit.cuh"
#include "walk.cuh"
#include "linkwalk.cuh"

#define FULL 0xffffffffu

namespace {

enum RayMode { RAY_IDLE = 0, RAY_FFS = 1, RAY_MAIN = 2, RAY_PEEL = 3 };
enum { SRC_PS = 0, SRC_BG = 1, SRC_HP = 2, SRC_CL = 3, SRC_ROI = 4 };         // == SimKind

struct Ray { vec3 pos, dir; float rho, tau; int level, ind; };

struct Lane {
    Ray r;              // the ray being stepped
    vec3 kpos, kdir;    // packet position / direction kept while look-ahead or peel-off rays are traced
    float krho, photons, free_path;
    float dleft, dscale;      // Healpix observer: distance left to the observer, 1/d^2
    int klevel, kind_, mode, idir, scat, nstep;
};

struct ScaCounters { unsigned long long packets, steps, scat, stuck, peels; };

// cos(theta) clamp before the DSC look-up: SimRAM_PS/PB use 0.999, SimRAM_HP/CL 0.9999 (kernel_ASOC_sca.c:991,1830 / 356,1329)
// kernel_ASOC_sca.c:349-355, 392-398, 1343-1349, 1387-1393: the `#ifdef HG_TEST` branch that is live in the shipped SimRAM_HP /
// SimRAM_CL (HG_TEST is #defined, as 0, in kernel_ASOC_aux.c:1): analytic Henyey-Greenstein function with g = 0.65 and
// the scattered instead of the transmitted fraction.  Only on request (ScaArgs.hg_test).
__device__ __forceinline__ float hg_test_weight(float cos_theta, float tau) {
    const float g = 0.65f;
    const float frac = (1.0f / (4.0f * SOC_PI)) * (1.0f - g * g) / powf(1.0f + g * g - 2.0f * g * cos_theta, 1.5f);
    return frac * ((tau > SOC_TAULIM) ? (1.0f - expf(-tau)) : (tau * (1.0f - 0.5f * tau)));
}
__device__ __forceinline__ float cos_clamp(const ScaArgs &S) { return (S.flavour >= 2) ? 0.9999f : 0.999f; }

template <class RNG, bool OCT>
__device__ __forceinline__ void begin_packet(const ScaArgs &S, Lane &L, RNG &rng, const Packet &pk) {
    L.kpos = pk.pos; L.kdir = pk.dir; L.krho = pk.rho; L.klevel = pk.level; L.kind_ = pk.ind;
    L.photons = pk.photons; L.scat = 0; L.nstep = 0;
    L.r.pos = pk.pos; L.r.dir = pk.dir; L.r.rho = pk.rho; L.r.level = pk.level; L.r.ind = pk.ind; L.r.tau = 0.0f;
    if (S.flavour >= 2 && pk.ind < 0) { L.mode = RAY_IDLE; return; }       // SimRAM_HP: `if (ind<0) continue`, no draw
    if (S.ffs > 0) {
        L.mode = RAY_FFS;
        if (pk.ind < 0) {              // nothing to look ahead through: tau = 0, the draw still happens
            float W = 0.0f;
            L.free_path = (S.flavour == 0) ? -logf(1.0f - W * rng.uniform()) : (float)(-log(1.0 - (double)(W * rng.uniform())));
            L.mode = RAY_IDLE;
        }
    } else {
        L.free_path = -logf(rng.uniform());
        L.mode = (pk.ind >= 0) ? RAY_MAIN : RAY_IDLE;
    }
}

template <bool OCT>
__device__ __forceinline__ void start_peel(const ScaArgs &S, Lane &L) {
    L.r.pos = L.kpos; L.r.level = L.klevel; L.r.ind = L.kind_; L.r.rho = L.krho; L.r.tau = 0.0f;
    if (S.nside > 0) {
        // Healpix image seen from the position odir[0..2]: kernel_ASOC_sca.c:312-332, 968-988, 1309-1329, 1807-1827
        vec3 q = L.kpos;
        if (OCT) root_position(S.G, q, L.klevel, L.kind_);
        vec3 od = { xsub(S.odir[0], q.x), xsub(S.odir[1], q.y), xsub(S.odir[2], q.z) };
        const float dx = sqrtf(xadd(xadd(xmul(od.x, od.x), xmul(od.y, od.y)), xmul(od.z, od.z)));
        L.dleft = dx;
        L.dscale = xdiv(1.0f, xmul(dx, dx));
        L.r.dir.x = xdiv(od.x, dx); L.r.dir.y = xdiv(od.y, dx); L.r.dir.z = xdiv(od.z, dx);
        return;
    }
    L.r.dir.x = S.odir[3 * L.idir]; L.r.dir.y = S.odir[3 * L.idir + 1]; L.r.dir.z = S.odir[3 * L.idir + 2];
}

// One iteration of the lane state machine: a GetStep for whichever ray is active, then the transition if the
// ray ended.
template <class RNG, bool OCT, bool DBL>
__device__ __forceinline__ void advance(const ScaArgs &S, Lane &L, RNG &rng, ScaCounters &cnt) {
    const GridDesc &G = S.G;
    Ray &r = L.r;
    const int oind = OCT ? G.off[r.level] + r.ind : r.ind;
    const int ind0 = r.ind, level0 = r.level;
    const vec3 pos0 = r.pos;
    const float rho0 = r.rho;
    float ds = get_step<OCT, DBL, false>(G, r.pos, r.dir, r.level, r.ind, r.rho);
    float kabs = S.kabs, ksca = S.ksca;
    if (S.with_abu) { float2 o = reinterpret_cast<const float2 *>(S.opt)[oind]; kabs = o.x; ksca = o.y; }
    cnt.steps++;
    if (L.mode == RAY_PEEL) {
        if (S.nside > 0) {
            // sic: the step is cut at the observer and lengthened by 1e-6 (a double literal in SimRAM_PB, :982)
            ds = (S.flavour == 1) ? (float)((double)fminf(L.dleft, ds) + 1.0e-6) : xadd(fminf(L.dleft, ds), 1.0e-6f);
            L.dleft = xsub(L.dleft, ds);
            r.tau = xadd(r.tau, xmul(xmul(ds, rho0), xadd(kabs, ksca)));
            if (L.dleft > 0.0f && r.ind >= 0) return;
        } else {
            r.tau += ds * rho0 * (kabs + ksca);
            if (r.ind >= 0) return;
        }
        // the peel-off ray has reached the surface (or the observer): kernel_ASOC_sca.c:1010-1046 / 1849-1885
        cnt.peels++;
        const float cc = cos_clamp(S);
        float cos_theta = clampf(dot3(L.kdir, r.dir), -cc, +cc);
        const int ocell = OCT ? G.off[L.klevel] + L.kind_ : L.kind_;
        const float *dsc = S.dsc;
        if (S.with_msf) dsc += S.bins * msf_pick(S.abu, S.scav, S.ndust, S.opt[2 * (size_t)ocell + 1], ocell, rng.uniform());
        float delta = L.photons * expf(-r.tau) * dsc[clampi((int)xmul(xmul((float)S.bins, xadd(1.0f, cos_theta)), 0.5f), 0, S.bins - 1)];
        if (S.hg_test && S.flavour >= 2) delta = L.photons * hg_test_weight(cos_theta, r.tau);
        if (S.nside > 0) {
            delta *= L.dscale;
            const float theta = acosf(-r.dir.z), phi = atan2f(r.dir.y, r.dir.x);
            const int ipix = ang2pix_ring(S.nside, phi, theta);
            if ((unsigned)ipix < 12u * S.nside * S.nside) atomicAdd(&S.out[ipix], delta);
            L.idir = 1;
        } else {
            vec3 p = { xsub(r.pos.x, S.centre.x), xsub(r.pos.y, S.centre.y), xsub(r.pos.z, S.centre.z) };
            const vec3 ra = { S.ora[3 * L.idir], S.ora[3 * L.idir + 1], S.ora[3 * L.idir + 2] };
            const vec3 de = { S.ode[3 * L.idir], S.ode[3 * L.idir + 1], S.ode[3 * L.idir + 2] };
            int i = (int)xadd(xsub(xmul(0.5f, (float)S.npx), 0.00005f), xdiv(dot3(p, ra), S.map_dx));
            int j = (int)xadd(xsub(xmul(0.5f, (float)S.npy), 0.00005f), xdiv(dot3(p, de), S.map_dx));
            if (i >= 0 && j >= 0 && i < S.npx && j < S.npy) atomicAdd(&S.out[i + L.idir * S.npx * S.npy + j * S.npx], delta);
            L.idir++;
        }
        if (L.idir < S.ndir) { start_peel<OCT>(S, L); return; }
        // all observers done: scatter and continue the random walk from the scattering point
        r.pos = L.kpos; r.level = L.klevel; r.ind = L.kind_; r.rho = L.krho; r.dir = L.kdir; r.tau = 0.0f;
        const float *csc = S.csc;
        if (S.with_msf) csc += S.bins * msf_pick(S.abu, S.scav, S.ndust, S.opt[2 * (size_t)ocell + 1], ocell, rng.uniform());
        scatter_direction(r.dir, csc, S.bins, rng);
        L.free_path = -logf(rng.uniform());
        L.mode = (L.scat == 30) ? RAY_IDLE : RAY_MAIN;                     // MAX_SCATTERINGS, kernel_ASOC_sca.c:5
        return;
    }
    if (L.mode == RAY_FFS) {                                                // kernel_ASOC_sca.c:888-910 / 1720-1750 / 268-290 / 1246-1266
        r.tau += ds * rho0 * ksca;
        if (r.ind >= 0) return;
        float tau = r.tau, W;
        if (S.flavour >= 2 && tau < 1.0e-22f) { L.mode = RAY_IDLE; return; }    // SimRAM_HP / CL: next packet, no draw
        if (S.flavour == 0) { W = -expm1f(-tau); L.free_path = -logf(1.0f - W * rng.uniform()); }
        // 1-exp(-tau) in single precision is quantised to 6e-8 (the reference's SimRAM_PB does exactly this); the
        // correctly rounded exp keeps the quantisation identical to a host OpenCL/libm evaluation
        else                { W = 1.0f - (float)exp(-(double)tau); L.free_path = (float)(-log(1.0 - (double)(W * rng.uniform()))); }
        L.photons *= W;
        r.pos = L.kpos; r.level = L.klevel; r.ind = L.kind_; r.rho = L.krho; r.tau = 0.0f;
        L.mode = (tau < 1.0e-22f) ? RAY_IDLE : RAY_MAIN;
        return;
    }
    // RAY_MAIN
    float dtau = ds * rho0 * ksca;
    if (L.free_path < (r.tau + dtau)) {
        L.scat++; cnt.scat++;
        dtau = L.free_path - r.tau;
        float dx = dtau / (ksca * rho0);
        if (OCT) dx = ldexpf(dx, r.level);             // sic: level of the cell *after* the step (kernel_ASOC_sca.c:958,1797)
        L.kpos.x = xadd(pos0.x, xmul(dx, r.dir.x)); L.kpos.y = xadd(pos0.y, xmul(dx, r.dir.y)); L.kpos.z = xadd(pos0.z, xmul(dx, r.dir.z));
        L.kdir = r.dir; L.klevel = level0; L.kind_ = ind0; L.krho = rho0;
        L.photons *= expf(-L.free_path * kabs / ksca);
        L.idir = 0; L.mode = RAY_PEEL;
        start_peel<OCT>(S, L);
        return;
    }
    r.tau += dtau;
    if (S.mirror && r.ind < 0) mirror_literal<OCT>(G, S.mirror, r.pos, r.dir, r.level, r.ind, r.rho);      // :280, 940, 1280, 1780
    if (r.ind < 0) L.mode = RAY_IDLE;
}

// one packet of a point source / the isotropic background / the Healpix sky (kernel_ASOC_sca.c:520-808, 104-220)
// Returns false when nothing is emitted (empty direction of the stored ROI field, kernel_ASOC_sca.c:812-840).
template <class RNG, bool OCT>
__device__ __forceinline__ bool emit_packet(const ScaArgs &S, RNG &rng, int id, int III, Packet &pk) {
    if (S.kind == SRC_PS)       { emit_ps<ScaArgs, RNG, OCT>(S, rng, III, pk); fix_direction(pk.dir); }
    else if (S.kind == SRC_BG)  { emit_bg<ScaArgs, RNG, OCT>(S, rng, id, pk); fix_direction(pk.dir); }
    else if (S.kind == SRC_ROI) { if (!emit_roi<ScaArgs, RNG, OCT>(S, rng, id, III, S.roi_nelem, pk)) return false; fix_direction(pk.dir); }
    else                        emit_hp_sca<ScaArgs, RNG, OCT>(S, rng, pk);
    return true;
}

__device__ __forceinline__ void flush(const ScaArgs &S, const ScaCounters &c) {
    warp_add_counter(S.counters + 0, c.packets);
    warp_add_counter(S.counters + 1, c.steps);
    warp_add_counter(S.counters + 2, c.scat);
    warp_add_counter(S.counters + 3, c.stuck);
    warp_add_counter(S.counters + 4, c.peels);
}

// thread <-> reference work item, MWC64X streams (parity layout)
template <bool OCT, bool DBL>
__global__ void __launch_bounds__(128) sca_item_kernel(const __grid_constant__ ScaArgs S) {
    ScaCounters cnt = { 0, 0, 0, 0, 0 };
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long id = t * S.world + S.rank;
    bool have = id < S.nunits;
    if (S.kind == SRC_BG) have = have && id < 8LL * S.G.area;
    if (S.kind == SRC_CL) have = have && id < S.G.cells;
    if (S.kind == SRC_ROI) have = have && id < 100LL * S.roi_nelem;
    RngMwc rng;
    if (have) rng.seed(S.mwc, (unsigned long long)id);
    Lane L; L.mode = RAY_IDLE;
    Packet pk; pk.ind = -1; pk.level = 0; pk.rho = 0.0f;
    int III = 0, icell = (int)id - S.global, nray = 0;     // CL: icell advances by `global` per cell
    float pwei = 1.0f;
    for (;;) {
        if (L.mode == RAY_IDLE && have) {
            if (S.kind == SRC_CL) {                          // kernel_ASOC_sca.c:1150-1230
                while (III >= nray) {
                    long long nc = (long long)icell + S.global;
                    if (nc >= S.G.cells) { have = false; break; }
                    icell = (int)nc; III = 0;
                    nray = cl_rays(S, icell, pwei);
                }
                if (have) {
                    III++;
                    emit_cl(S, rng, icell, pwei, pk);
                    fix_direction(pk.dir);
                    cnt.packets++;
                    begin_packet<RngMwc, OCT>(S, L, rng, pk);
                }
            } else if (III < S.batch) {
                if (emit_packet<RngMwc, OCT>(S, rng, (int)id, III, pk)) {
                    cnt.packets++;
                    begin_packet<RngMwc, OCT>(S, L, rng, pk);
                }
                III++;
            } else have = false;
        }
        if (!__any_sync(FULL, L.mode != RAY_IDLE || have)) break;
        if (L.mode != RAY_IDLE) {
            advance<RngMwc, OCT, DBL>(S, L, rng, cnt);
            if (++L.nstep > S.max_steps) { L.mode = RAY_IDLE; cnt.stuck++; }
        }
    }
    flush(S, cnt);
}

// persistent warps, one Philox stream per packet (CL: per cell and ray), idle lanes refilled from the work counter
template <bool OCT, bool DBL>
__global__ void __launch_bounds__(128) sca_stream_kernel(const __grid_constant__ ScaArgs S) {
    ScaCounters cnt = { 0, 0, 0, 0, 0 };
    const int lane = threadIdx.x & 31;
    const long long nlocal = (S.nunits - S.rank + S.world - 1) / S.world;
    RngPhilox rng;
    Lane L; L.mode = RAY_IDLE;
    Packet pk; pk.ind = -1; pk.level = 0; pk.rho = 0.0f;
    bool more = true;
    int icell = 0, iray = 0, nray = 0;
    float pwei = 1.0f;
    for (;;) {
        unsigned idle = __ballot_sync(FULL, L.mode == RAY_IDLE);
        if (idle == FULL || (__popc(idle) >= 8 && __any_sync(FULL, more))) {
            bool need = L.mode == RAY_IDLE && more && iray >= nray;
            unsigned nm = __ballot_sync(FULL, need);
            if (nm) {
                int leader = __ffs(nm) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(S.work, (unsigned long long)__popc(nm));
                base = __shfl_sync(FULL, base, leader);
                if (need) {
                    long long u = (long long)base + __popc(nm & ((1u << lane) - 1u));
                    if (u >= nlocal) more = false;
                    else {
                        unsigned long long q = (unsigned long long)u * S.world + S.rank;
                        if (S.kind == SRC_CL) { icell = (int)q; iray = 0; nray = cl_rays(S, icell, pwei); }
                        else {
                            rng.seed(S.phx, q);
                            if (emit_packet<RngPhilox, OCT>(S, rng, (int)(q / (unsigned)S.batch), (int)(q % (unsigned)S.batch), pk)) {
                                cnt.packets++;
                                begin_packet<RngPhilox, OCT>(S, L, rng, pk);
                            }
                        }
                    }
                }
            }
            if (S.kind == SRC_CL && L.mode == RAY_IDLE && iray < nray) {
                rng.seed(S.phx, (unsigned long long)(unsigned)icell | ((unsigned long long)(unsigned)iray << 32));
                iray++;
                emit_cl(S, rng, icell, pwei, pk);
                fix_direction(pk.dir);
                cnt.packets++;
                begin_packet<RngPhilox, OCT>(S, L, rng, pk);
            }
            if (!__any_sync(FULL, L.mode != RAY_IDLE || more || iray < nray)) break;
        }
        if (L.mode != RAY_IDLE) {
            advance<RngPhilox, OCT, DBL>(S, L, rng, cnt);
            if (++L.nstep > S.max_steps) { L.mode = RAY_IDLE; cnt.stuck++; }
        }
    }
    flush(S, cnt);
}

// =================================================================================================================
// Production kernel: the three kinds of rays walk with the incremental octree walker of walk.cuh (regular grids
// are the LEVELS == 1 case), one hop per iteration.  The packet's scattering point is kept as (cell, fractional
// position, direction); look-ahead and peel-off rays are started from it with walker_set_direction().  The image
// pixel of a peel-off ray follows from the scattering point alone: the exit point differs from it by a multiple of
// the observer direction, which is perpendicular to the image axes RA and DE.
// =================================================================================================================
struct ScatterPoint {
    float fx, fy, fz;          // fractional position inside the cell
    vec3 dir;                  // packet direction (before the next scattering)
    vec3 gpos;                 // root-grid position (for the image pixel), advanced along the walk
    float rho;
    int level, ind, ix, iy, iz;
};

__device__ __forceinline__ void ray_from_point(Walker &w, const ScatterPoint &k, const vec3 &dir) {
    w.level = k.level; w.ind = k.ind; w.ix = k.ix; w.iy = k.iy; w.iz = k.iz; w.rho = k.rho;
    walker_set_direction(w, dir, k.fx, k.fy, k.fz);
}

__device__ __forceinline__ float uniform_fast_log(float u) { return -__logf(u); }

template <bool OCT>
__global__ void __launch_bounds__(256, 3) sca_walk_kernel(const __grid_constant__ ScaArgs S) {
    const GridDesc &G = S.G;
    ScaCounters cnt = { 0, 0, 0, 0, 0 };
    const int lane = threadIdx.x & 31;
    const long long nlocal = (S.nunits - S.rank + S.world - 1) / S.world;
    Walker w; w.ind = -1; w.level = 0;
    ScatterPoint k;
    int mode = RAY_IDLE, phase = WALK_LEAF, ax = 0, idir = 0, scat = 0, nstep = 0, nevent = 0;
    float photons = 0.0f, free_path = 0.0f, tau = 0.0f;
    float dleft = 0.0f, dscale = 1.0f;            // Healpix observer: distance left on the peel-off ray, 1/d^2
    unsigned long long rid = 0;
    bool more = true;
    int icell = 0, iray = 0, nray = 0;            // SimRAM_CL: cell of this lane and its rays
    float pwei = 1.0f;
    const float cclamp = cos_clamp(S);
    for (;;) {
        unsigned idle = __ballot_sync(FULL, mode == RAY_IDLE);
        if (idle == FULL || (__popc(idle) >= 8 && __any_sync(FULL, more))) {
            bool need = mode == RAY_IDLE && more && iray >= nray;
            unsigned nm = __ballot_sync(FULL, need);
            bool got = false;
            if (nm) {
                int leader = __ffs(nm) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(S.work, (unsigned long long)__popc(nm));
                base = __shfl_sync(FULL, base, leader);
                if (need) {
                    long long u = (long long)base + __popc(nm & ((1u << lane) - 1u));
                    if (u >= nlocal) more = false;
                    else {
                        rid = (unsigned long long)u * S.world + S.rank;
                        if (S.kind == SRC_CL) { icell = (int)rid; iray = 0; nray = cl_rays(S, icell, pwei); }
                        else got = true;
                    }
                }
            }
            if (S.kind == SRC_CL && mode == RAY_IDLE && iray < nray) {
                rid = (unsigned long long)(unsigned)icell | ((unsigned long long)(unsigned)iray << 32);
                iray++; got = true;
            }
            if (got) {
                RngPhilox rng; rng.seed(S.phx, rid);
                Packet pk; pk.ind = -1; pk.level = 0; pk.rho = 0.0f;
                bool emitted = true;
                if (S.kind == SRC_CL) { emit_cl(S, rng, icell, pwei, pk); fix_direction(pk.dir); }
                else emitted = emit_packet<RngPhilox, OCT>(S, rng, (int)(rid / (unsigned)S.batch), (int)(rid % (unsigned)S.batch), pk);
                if (emitted) cnt.packets++; else pk.ind = -1;
                photons = pk.photons; scat = 0; nstep = 0; nevent = 0; tau = 0.0f; phase = WALK_LEAF;
                if (pk.ind >= 0) {
                    k.level = pk.level; k.ind = pk.ind; k.rho = pk.rho; k.dir = pk.dir;
                    k.fx = pk.pos.x - floorf(pk.pos.x); k.fy = pk.pos.y - floorf(pk.pos.y); k.fz = pk.pos.z - floorf(pk.pos.z);
                    int root = pk.ind;
                    if (OCT) for (int l = pk.level; l > 0; l--) root = G.par[G.off[l] + root - G.nxyz];
                    k.ix = root % G.nx; k.iy = (root / G.nx) % G.ny; k.iz = root / (G.nx * G.ny);
                    k.gpos = pk.pos;
                    if (OCT) root_position(G, k.gpos, pk.level, pk.ind);
                    ray_from_point(w, k, k.dir);
                    if (S.ffs > 0) mode = RAY_FFS;
                    else { free_path = uniform_fast_log(rng.uniform()); mode = RAY_MAIN; }
                }
            }
            if (!__any_sync(FULL, mode != RAY_IDLE || more || iray < nray)) break;
        }
        // ---- rays that have reached the surface, several lanes at a time (the block is long and rare per lane) ----
        const unsigned em = __ballot_sync(FULL, mode != RAY_IDLE && phase == WALK_END);
        if (em && (__popc(em) >= S.ev_batch || !__any_sync(FULL, mode != RAY_IDLE && phase != WALK_END)))
        if (mode != RAY_IDLE && phase == WALK_END) {
            phase = WALK_LEAF; nstep = 0;
            if (mode == RAY_PEEL) {                                  // kernel_ASOC_sca.c:1010-1046 / 1849-1885
                cnt.peels++;
                const vec3 od = w.d;
                float cos_theta = clampf(k.dir.x * od.x + k.dir.y * od.y + k.dir.z * od.z, -cclamp, +cclamp);
                const int kcell = OCT ? G.off[k.level] + k.ind : k.ind;
                // WITH_MSF: one Philox block per peel-off ray / scattering for the dust species draws
                const float *dsc = S.dsc;
                if (S.with_msf) {
                    RngBlock rm(S.phx, rid, 0x20000u + (unsigned)(scat * 64 + idir));
                    dsc += S.bins * msf_pick(S.abu, S.scav, S.ndust, __ldg(S.opt + 2 * (size_t)kcell + 1), kcell, rm.uniform());
                }
                float delta = photons * __expf(-tau) * __ldg(dsc + clampi((int)(S.bins * (1.0f + cos_theta) * 0.5f), 0, S.bins - 1));
                if (S.hg_test && S.flavour >= 2) delta = photons * hg_test_weight(cos_theta, tau);
                if (S.nside > 0) {                                    // Healpix image seen from odir[0..2]
                    const int ipix = ang2pix_ring(S.nside, atan2f(od.y, od.x), acosf(clampf(-od.z, -1.0f, 1.0f)));
                    if ((unsigned)ipix < 12u * S.nside * S.nside) atomicAdd(&S.out[ipix], delta * dscale);
                    idir = S.ndir;
                } else {
                    vec3 p = { k.gpos.x - S.centre.x, k.gpos.y - S.centre.y, k.gpos.z - S.centre.z };
                    const float *ra = S.ora + 3 * idir, *de = S.ode + 3 * idir;
                    int i = (int)((0.5f * S.npx - 0.00005f) + (p.x * ra[0] + p.y * ra[1] + p.z * ra[2]) / S.map_dx);
                    int j = (int)((0.5f * S.npy - 0.00005f) + (p.x * de[0] + p.y * de[1] + p.z * de[2]) / S.map_dx);
                    if (i >= 0 && j >= 0 && i < S.npx && j < S.npy) atomicAdd(&S.out[i + idir * S.npx * S.npy + j * S.npx], delta);
                    idir++;
                }
                tau = 0.0f;
                if (idir < S.ndir) {
                    vec3 nd = { S.odir[3 * idir], S.odir[3 * idir + 1], S.odir[3 * idir + 2] };
                    ray_from_point(w, k, nd);
                } else if (scat == 30) mode = RAY_IDLE;               // MAX_SCATTERINGS, kernel_ASOC_sca.c:5
                else {
                    RngBlock rb(S.phx, rid, 0x10000u + (unsigned)(nevent++));
                    const float u_ct = rb.uniform(), u_phi = rb.uniform(), u_fp = rb.uniform();
                    const float *csc = S.csc;
                    if (S.with_msf) csc += S.bins * msf_pick(S.abu, S.scav, S.ndust, __ldg(S.opt + 2 * (size_t)kcell + 1), kcell, rb.uniform());
                    float ct = __ldg(csc + clampi((int)(u_ct * S.bins), 0, S.bins - 1));
                    vec3 nd = k.dir;
                    scatter_rotate(nd, ct, SOC_TWOPI * u_phi);
                    free_path = uniform_fast_log(u_fp);
                    k.dir = nd;
                    ray_from_point(w, k, nd);
                    mode = RAY_MAIN;
                }
            } else if (mode == RAY_FFS) {                            // kernel_ASOC_sca.c:888-910 / 1720-1750
                if (S.flavour >= 2 && tau < 1.0e-22f) mode = RAY_IDLE;         // SimRAM_HP / CL: nothing on the line of sight
                else {
                    RngBlock rb(S.phx, rid, 0x10000u + (unsigned)(nevent++));
                    float W;
                    if (S.flavour == 0) { W = -expm1f(-tau); free_path = -logf(1.0f - W * rb.uniform()); }
                    else                { W = 1.0f - (float)exp(-(double)tau); free_path = (float)(-log(1.0 - (double)(W * rb.uniform()))); }
                    photons *= W;
                    mode = (tau < 1.0e-22f) ? RAY_IDLE : RAY_MAIN;
                    tau = 0.0f;
                    ray_from_point(w, k, k.dir);
                }
            } else mode = RAY_IDLE;                                   // the packet itself has left the cloud
        }
        // ---- one cell of whichever ray the lane is tracing -----------------------------------------------------
        const bool ready = mode != RAY_IDLE && phase == WALK_LEAF;
        if (ready) {
            const int oind = OCT ? G.off[w.level] + w.ind : w.ind;
            const float tmin = fminf(w.tx, fminf(w.ty, w.tz));
            ax = (w.tx <= w.ty && w.tx <= w.tz) ? 0 : ((w.ty <= w.tz) ? 1 : 2);
            float ds = fmaxf(tmin, 0.0f);
            float kabs = S.kabs, ksca = S.ksca;
            if (S.with_abu) { float2 o = __ldg(reinterpret_cast<const float2 *>(S.opt) + oind); kabs = o.x; ksca = o.y; }
            cnt.steps++; nstep++;
            bool go = true;
            if (mode == RAY_PEEL) {
                if (S.nside > 0) {                                    // the ray ends at the observer
                    ds = fminf(ds, dleft);
                    dleft -= ds;
                    if (dleft <= 0.0f) { go = false; phase = WALK_END; }
                }
                tau += ds * w.rho * (kabs + ksca);
            }
            else if (mode == RAY_FFS) tau += ds * w.rho * ksca;
            else {
                const float dtau = ds * w.rho * ksca;
                if (free_path < tau + dtau) {
                    // scattering inside this cell: remember the point, start the peel-off rays
                    ds = fminf(ds, (free_path - tau) / (ksca * w.rho));
                    w.tx -= ds; w.ty -= ds; w.tz -= ds;
                    walker_fraction(w, k.fx, k.fy, k.fz);
                    k.level = w.level; k.ind = w.ind; k.ix = w.ix; k.iy = w.iy; k.iz = w.iz; k.rho = w.rho; k.dir = w.d;
                    k.gpos.x += ds * w.d.x; k.gpos.y += ds * w.d.y; k.gpos.z += ds * w.d.z;
                    photons *= __expf(-free_path * kabs / ksca);
                    scat++; cnt.scat++;
                    idir = 0; mode = RAY_PEEL; tau = 0.0f; nstep = 0;
                    vec3 od;
                    if (S.nside > 0) {
                        od.x = S.odir[0] - k.gpos.x; od.y = S.odir[1] - k.gpos.y; od.z = S.odir[2] - k.gpos.z;
                        const float d2 = od.x * od.x + od.y * od.y + od.z * od.z;
                        const float id_ = rsqrtf(fmaxf(d2, 1.0e-30f));
                        dleft = d2 * id_; dscale = id_ * id_;
                        od.x *= id_; od.y *= id_; od.z *= id_;
                        // the walker needs non-zero direction components
                        if (fabsf(od.x) < 1.0e-6f) od.x = 1.0e-6f;
                        if (fabsf(od.y) < 1.0e-6f) od.y = 1.0e-6f;
                        if (fabsf(od.z) < 1.0e-6f) od.z = 1.0e-6f;
                    } else { od.x = S.odir[0]; od.y = S.odir[1]; od.z = S.odir[2]; }
                    ray_from_point(w, k, od);
                    go = false;
                } else {
                    tau += dtau;
                    k.gpos.x += ds * w.d.x; k.gpos.y += ds * w.d.y; k.gpos.z += ds * w.d.z;
                }
            }
            if (go) { w.tx -= tmin; w.ty -= tmin; w.tz -= tmin; phase = WALK_CROSS; }
            if (nstep > S.max_steps) { mode = RAY_IDLE; phase = WALK_LEAF; cnt.stuck++; }
        }
        // ---- navigation: one hop of each kind ------------------------------------------------------------------
        #pragma unroll 1
        for (int hop = 0; hop < (OCT ? S.nav_hops : 1); hop++) {
            if (mode != RAY_IDLE) {
                if (OCT && phase == WALK_CLIMB) { nav_climb(G, w, ax); phase = WALK_CROSS; }
                if (phase == WALK_CROSS) {
                    // only the packet itself is reflected by a mirror border; look-ahead and peel-off rays leave
                    phase = nav_cross(G, w, ax, mode == RAY_MAIN ? S.mirror : 0);
                    if (w.ind < 0) phase = WALK_END;
                }
                if (OCT && phase == WALK_DESCEND) phase = nav_descend(G, w, ax);
            }
        }
    }
    flush(S, cnt);
}

struct LScatterPoint {
    float fx, fy, fz;
    vec3 dir, gpos;
    float rho;
    int level, cell, cx, cy, cz;
};
__device__ __forceinline__ void ray_from_point(LWalker &w, const LScatterPoint &k, const vec3 &dir) {
    w.level = k.level; w.cell = k.cell; w.cx = k.cx; w.cy = k.cy; w.cz = k.cz; w.rho = k.rho;
    lw_set_direction(w, dir, k.fx, k.fy, k.fz);
}

// The same kernel on the neighbour table of linkwalk.cuh (octrees): a face crossing is one table look-up, no climbs.
#ifndef SOC_SCA_REPS
#define SOC_SCA_REPS 2
#endif
template <bool OCT>
__global__ void __launch_bounds__(256, 3) sca_link_kernel(const __grid_constant__ ScaArgs S) {
    const GridDesc &G = S.G;
    // per-lane work counters in 32 bits (64-bit ones cost ten registers in the loop); moved to the global counters at the
    // end of the launch, steps and peel-off rays also whenever a lane passes 2^31
    unsigned c_packets = 0u, c_steps = 0u, c_scat = 0u, c_stuck = 0u, c_peels = 0u;
    const int lane = threadIdx.x & 31;
    const long long nlocal = (S.nunits - S.rank + S.world - 1) / S.world;
    LWalker w; w.cell = -1; w.level = 0;
    LScatterPoint k;
    const int *__restrict__ nbr = S.nbr;
    int mode = RAY_IDLE, phase = WALK_LEAF, ax = 0, idir = 0, scat = 0, nstep = 0, nevent = 0;
    float photons = 0.0f, free_path = 0.0f, tau = 0.0f;
    float dleft = 0.0f, dscale = 1.0f;            // Healpix observer: distance left on the peel-off ray, 1/d^2
    unsigned long long rid = 0;
    bool more = true;
    int icell = 0, iray = 0, nray = 0;            // SimRAM_CL: cell of this lane and its rays
    float pwei = 1.0f;
    const float cclamp = cos_clamp(S);
    for (;;) {
        unsigned idle = __ballot_sync(FULL, mode == RAY_IDLE);
        if (idle == FULL || (__popc(idle) >= 8 && __any_sync(FULL, more))) {
            bool need = mode == RAY_IDLE && more && iray >= nray;
            unsigned nm = __ballot_sync(FULL, need);
            bool got = false;
            if (nm) {
                int leader = __ffs(nm) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(S.work, (unsigned long long)__popc(nm));
                base = __shfl_sync(FULL, base, leader);
                if (need) {
                    long long u = (long long)base + __popc(nm & ((1u << lane) - 1u));
                    if (u >= nlocal) more = false;
                    else {
                        rid = (unsigned long long)u * S.world + S.rank;
                        if (S.kind == SRC_CL) { icell = (int)rid; iray = 0; nray = cl_rays(S, icell, pwei); }
                        else got = true;
                    }
                }
            }
            if (S.kind == SRC_CL && mode == RAY_IDLE && iray < nray) {
                rid = (unsigned long long)(unsigned)icell | ((unsigned long long)(unsigned)iray << 32);
                iray++; got = true;
            }
            if (got) {
                RngPhilox rng; rng.seed(S.phx, rid);
                Packet pk; pk.ind = -1; pk.level = 0; pk.rho = 0.0f;
                bool emitted = true;
                if (S.kind == SRC_CL) { emit_cl(S, rng, icell, pwei, pk); fix_direction(pk.dir); }
                else emitted = emit_packet<RngPhilox, OCT>(S, rng, (int)(rid / (unsigned)S.batch), (int)(rid % (unsigned)S.batch), pk);
                if (emitted) c_packets++; else pk.ind = -1;
                if ((c_steps | c_peels) & 0x80000000u) {
                    atomicAdd(S.counters + 1, (unsigned long long)c_steps); atomicAdd(S.counters + 4, (unsigned long long)c_peels);
                    c_steps = 0u; c_peels = 0u;
                }
                photons = pk.photons; scat = 0; nstep = 0; nevent = 0; tau = 0.0f; phase = WALK_LEAF;
                if (pk.ind >= 0) {
                    lw_init(G, w, pk.pos, pk.dir, pk.level, pk.ind, pk.rho);
                    k.level = w.level; k.cell = w.cell; k.cx = w.cx; k.cy = w.cy; k.cz = w.cz; k.rho = pk.rho; k.dir = pk.dir;
                    k.fx = pk.pos.x - floorf(pk.pos.x); k.fy = pk.pos.y - floorf(pk.pos.y); k.fz = pk.pos.z - floorf(pk.pos.z);
                    k.gpos = pk.pos;
                    if (OCT) root_position(G, k.gpos, pk.level, pk.ind);
                    ray_from_point(w, k, k.dir);
                    if (S.ffs > 0) mode = RAY_FFS;
                    else { free_path = uniform_fast_log(rng.uniform()); mode = RAY_MAIN; }
                }
            }
            if (!__any_sync(FULL, mode != RAY_IDLE || more || iray < nray)) break;
        }
        // ---- rays that have reached the surface, several lanes at a time (the block is long and rare per lane) ----
        const unsigned em = __ballot_sync(FULL, mode != RAY_IDLE && phase == WALK_END);
        if (em && (__popc(em) >= S.ev_batch || !__any_sync(FULL, mode != RAY_IDLE && phase != WALK_END)))
        if (mode != RAY_IDLE && phase == WALK_END) {
            phase = WALK_LEAF; nstep = 0;
            if (mode == RAY_PEEL) {                                  // kernel_ASOC_sca.c:1010-1046 / 1849-1885
                c_peels++;
                const vec3 od = lw_dir(w);
                float cos_theta = clampf(k.dir.x * od.x + k.dir.y * od.y + k.dir.z * od.z, -cclamp, +cclamp);
                const int kcell = k.cell;
                // WITH_MSF: one Philox block per peel-off ray / scattering for the dust species draws
                const float *dsc = S.dsc;
                if (S.with_msf) {
                    RngBlock rm(S.phx, rid, 0x20000u + (unsigned)(scat * 64 + idir));
                    dsc += S.bins * msf_pick(S.abu, S.scav, S.ndust, __ldg(S.opt + 2 * (size_t)kcell + 1), kcell, rm.uniform());
                }
                float delta = photons * __expf(-tau) * __ldg(dsc + clampi((int)(S.bins * (1.0f + cos_theta) * 0.5f), 0, S.bins - 1));
                if (S.hg_test && S.flavour >= 2) delta = photons * hg_test_weight(cos_theta, tau);
                if (S.nside > 0) {                                    // Healpix image seen from odir[0..2]
                    const int ipix = ang2pix_ring(S.nside, atan2f(od.y, od.x), acosf(clampf(-od.z, -1.0f, 1.0f)));
                    if ((unsigned)ipix < 12u * S.nside * S.nside) atomicAdd(&S.out[ipix], delta * dscale);
                    idir = S.ndir;
                } else {
                    vec3 p = { k.gpos.x - S.centre.x, k.gpos.y - S.centre.y, k.gpos.z - S.centre.z };
                    const float *ra = S.ora + 3 * idir, *de = S.ode + 3 * idir;
                    int i = (int)((0.5f * S.npx - 0.00005f) + (p.x * ra[0] + p.y * ra[1] + p.z * ra[2]) / S.map_dx);
                    int j = (int)((0.5f * S.npy - 0.00005f) + (p.x * de[0] + p.y * de[1] + p.z * de[2]) / S.map_dx);
                    if (i >= 0 && j >= 0 && i < S.npx && j < S.npy) atomicAdd(&S.out[i + idir * S.npx * S.npy + j * S.npx], delta);
                    idir++;
                }
                tau = 0.0f;
                if (idir < S.ndir) {
                    vec3 nd = { S.odir[3 * idir], S.odir[3 * idir + 1], S.odir[3 * idir + 2] };
                    ray_from_point(w, k, nd);
                } else if (scat == 30) mode = RAY_IDLE;               // MAX_SCATTERINGS, kernel_ASOC_sca.c:5
                else {
                    RngBlock rb(S.phx, rid, 0x10000u + (unsigned)(nevent++));
                    const float u_ct = rb.uniform(), u_phi = rb.uniform(), u_fp = rb.uniform();
                    const float *csc = S.csc;
                    if (S.with_msf) csc += S.bins * msf_pick(S.abu, S.scav, S.ndust, __ldg(S.opt + 2 * (size_t)kcell + 1), kcell, rb.uniform());
                    float ct = __ldg(csc + clampi((int)(u_ct * S.bins), 0, S.bins - 1));
                    vec3 nd = k.dir;
                    scatter_rotate(nd, ct, SOC_TWOPI * u_phi);
                    free_path = uniform_fast_log(u_fp);
                    k.dir = nd;
                    ray_from_point(w, k, nd);
                    mode = RAY_MAIN;
                }
            } else if (mode == RAY_FFS) {                            // kernel_ASOC_sca.c:888-910 / 1720-1750
                if (S.flavour >= 2 && tau < 1.0e-22f) mode = RAY_IDLE;         // SimRAM_HP / CL: nothing on the line of sight
                else {
                    RngBlock rb(S.phx, rid, 0x10000u + (unsigned)(nevent++));
                    float W;
                    if (S.flavour == 0) { W = -expm1f(-tau); free_path = -logf(1.0f - W * rb.uniform()); }
                    else                { W = 1.0f - (float)exp(-(double)tau); free_path = (float)(-log(1.0 - (double)(W * rb.uniform()))); }
                    photons *= W;
                    mode = (tau < 1.0e-22f) ? RAY_IDLE : RAY_MAIN;
                    tau = 0.0f;
                    ray_from_point(w, k, k.dir);
                }
            } else mode = RAY_IDLE;                                   // the packet itself has left the cloud
        }
        #pragma unroll
        for (int rep = 0; rep < SOC_SCA_REPS; rep++) {           // cells per refill / ray-end check
        // ---- one cell of whichever ray the lane is tracing -----------------------------------------------------
        const bool ready = mode != RAY_IDLE && phase == WALK_LEAF;
        int2 e2 = make_int2(0, 0);
        if (ready) {
            const int oind = w.cell;
            const float tmin = fminf(w.tx, fminf(w.ty, w.tz));
            ax = (w.tx <= w.ty && w.tx <= w.tz) ? 0 : ((w.ty <= w.tz) ? 1 : 2);
            e2 = lw_entry(nbr, w, ax);                                // the table entry behind the exit face, used after the physics
            float ds = fmaxf(tmin, 0.0f);
            float kabs = S.kabs, ksca = S.ksca;
            if (S.with_abu) { float2 o = __ldg(reinterpret_cast<const float2 *>(S.opt) + oind); kabs = o.x; ksca = o.y; }
            c_steps++; nstep++;
            bool go = true;
            if (mode == RAY_PEEL) {
                if (S.nside > 0) {                                    // the ray ends at the observer
                    ds = fminf(ds, dleft);
                    dleft -= ds;
                    if (dleft <= 0.0f) { go = false; phase = WALK_END; }
                }
                tau += ds * w.rho * (kabs + ksca);
            }
            else if (mode == RAY_FFS) tau += ds * w.rho * ksca;
            else {
                const float dtau = ds * w.rho * ksca;
                if (free_path < tau + dtau) {
                    // scattering inside this cell: remember the point, start the peel-off rays
                    ds = fminf(ds, (free_path - tau) * rcp_approx(ksca * w.rho));
                    w.tx -= ds; w.ty -= ds; w.tz -= ds;
                    lw_fraction(w, k.fx, k.fy, k.fz);
                    k.level = w.level; k.cell = w.cell; k.cx = w.cx; k.cy = w.cy; k.cz = w.cz; k.rho = w.rho; k.dir = lw_dir(w);
                    k.gpos.x += ds * k.dir.x; k.gpos.y += ds * k.dir.y; k.gpos.z += ds * k.dir.z;
                    photons *= __expf(-free_path * kabs / ksca);
                    scat++; c_scat++;
                    idir = 0; mode = RAY_PEEL; tau = 0.0f; nstep = 0;
                    vec3 od;
                    if (S.nside > 0) {
                        od.x = S.odir[0] - k.gpos.x; od.y = S.odir[1] - k.gpos.y; od.z = S.odir[2] - k.gpos.z;
                        const float d2 = od.x * od.x + od.y * od.y + od.z * od.z;
                        const float id_ = rsqrtf(fmaxf(d2, 1.0e-30f));
                        dleft = d2 * id_; dscale = id_ * id_;
                        od.x *= id_; od.y *= id_; od.z *= id_;
                        // the walker needs non-zero direction components
                        if (fabsf(od.x) < 1.0e-6f) od.x = 1.0e-6f;
                        if (fabsf(od.y) < 1.0e-6f) od.y = 1.0e-6f;
                        if (fabsf(od.z) < 1.0e-6f) od.z = 1.0e-6f;
                    } else { od.x = S.odir[0]; od.y = S.odir[1]; od.z = S.odir[2]; }
                    ray_from_point(w, k, od);
                    go = false;
                } else {
                    tau += dtau;
                    const vec3 wd = lw_dir(w);
                    k.gpos.x += ds * wd.x; k.gpos.y += ds * wd.y; k.gpos.z += ds * wd.z;
                }
            }
            if (go) { w.tx -= tmin; w.ty -= tmin; w.tz -= tmin; phase = WALK_CROSS; }
            if (nstep > S.max_steps) { mode = RAY_IDLE; phase = WALK_LEAF; c_stuck++; }
        }
        // ---- navigation: table look-up per crossing, one descent per iteration while the cell entered is refined --------
        if (mode != RAY_IDLE && phase == WALK_DESCEND) phase = lw_descend(G, w, ax) ? WALK_LEAF : WALK_DESCEND;
        if (mode != RAY_IDLE && phase == WALK_CROSS) {
            // only the packet itself is reflected by a mirror border; look-ahead and peel-off rays leave
            phase = lw_cross(G, e2, w, ax, mode == RAY_MAIN ? S.mirror : 0) ? WALK_LEAF : WALK_DESCEND;
            if (w.cell < 0) phase = WALK_END;
            else if (phase == WALK_DESCEND) phase = lw_descend(G, w, ax) ? WALK_LEAF : WALK_DESCEND;
        }
        }
    }
    { const ScaCounters cnt = { c_packets, c_steps, c_scat, c_stuck, c_peels }; flush(S, cnt); }
}

}  // namespace

void launch_sca(const ScaArgs &S, int rng_mode, int blocks, int threads, cudaStream_t stream) {
    const bool oct = S.G.levels > 1, dbl = S.G.dbl_sim != 0;
    if (rng_mode == SOC_RNG_REFERENCE) {
        if (!oct)      sca_item_kernel<false, false><<<blocks, threads, 0, stream>>>(S);
        else if (!dbl) sca_item_kernel<true, false><<<blocks, threads, 0, stream>>>(S);
        else           sca_item_kernel<true, true><<<blocks, threads, 0, stream>>>(S);
    } else if (!S.ref_geometry) {
        if (!oct)       sca_walk_kernel<false><<<blocks, threads, 0, stream>>>(S);
        else if (S.nbr) sca_link_kernel<true><<<blocks, threads, 0, stream>>>(S);
        else            sca_walk_kernel<true><<<blocks, threads, 0, stream>>>(S);
    } else {
        if (!oct)      sca_stream_kernel<false, false><<<blocks, threads, 0, stream>>>(S);
        else if (!dbl) sca_stream_kernel<true, false><<<blocks, threads, 0, stream>>>(S);
        else           sca_stream_kernel<true, true><<<blocks, threads, 0, stream>>>(S);
    }
}

int sca_blocks_per_sm(bool octree, bool dbl, int threads) {
    int n = 0;
    if (threads == 256) {
        if (!octree) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sca_walk_kernel<false>, threads, 0);
        else         cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sca_walk_kernel<true>, threads, 0);
        return n > 0 ? n : 1;
    }
    if (!octree)   cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sca_stream_kernel<false, false>, threads, 0);
    else if (!dbl) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sca_stream_kernel<true, false>, threads, 0);
    else           cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sca_stream_kernel<true, true>, threads, 0);
    return n > 0 ? n : 1;
}
