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

#define FULL 0xffffffffu

namespace {

enum RayMode { RAY_IDLE = 0, RAY_FFS = 1, RAY_MAIN = 2, RAY_PEEL = 3 };

struct Ray { vec3 pos, dir; float rho, tau; int level, ind; };

struct Lane {
    Ray r;              // the ray being stepped
    vec3 kpos, kdir;    // packet position / direction kept while look-ahead or peel-off rays are traced
    float krho, photons, free_path;
    int klevel, kind_, mode, idir, scat, nstep;
};

struct ScaCounters { unsigned long long packets, steps, scat, stuck, peels; };

template <class RNG, bool OCT>
__device__ __forceinline__ void begin_packet(const ScaArgs &S, Lane &L, RNG &rng, const Packet &pk) {
    L.kpos = pk.pos; L.kdir = pk.dir; L.krho = pk.rho; L.klevel = pk.level; L.kind_ = pk.ind;
    L.photons = pk.photons; L.scat = 0; L.nstep = 0;
    L.r.pos = pk.pos; L.r.dir = pk.dir; L.r.rho = pk.rho; L.r.level = pk.level; L.r.ind = pk.ind; L.r.tau = 0.0f;
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

__device__ __forceinline__ void start_peel(const ScaArgs &S, Lane &L) {
    L.r.pos = L.kpos; L.r.level = L.klevel; L.r.ind = L.kind_; L.r.rho = L.krho; L.r.tau = 0.0f;
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
        r.tau += ds * rho0 * (kabs + ksca);
        if (r.ind >= 0) return;
        // the peel-off ray has reached the surface: kernel_ASOC_sca.c:1010-1046 / 1849-1885
        cnt.peels++;
        float cos_theta = clampf(dot3(L.kdir, r.dir), -0.999f, +0.999f);
        float delta = L.photons * expf(-r.tau) * S.dsc[clampi((int)xmul(xmul((float)S.bins, xadd(1.0f, cos_theta)), 0.5f), 0, S.bins - 1)];
        vec3 p = { xsub(r.pos.x, S.centre.x), xsub(r.pos.y, S.centre.y), xsub(r.pos.z, S.centre.z) };
        const vec3 ra = { S.ora[3 * L.idir], S.ora[3 * L.idir + 1], S.ora[3 * L.idir + 2] };
        const vec3 de = { S.ode[3 * L.idir], S.ode[3 * L.idir + 1], S.ode[3 * L.idir + 2] };
        int i = (int)xadd(xsub(xmul(0.5f, (float)S.npx), 0.00005f), xdiv(dot3(p, ra), S.map_dx));
        int j = (int)xadd(xsub(xmul(0.5f, (float)S.npy), 0.00005f), xdiv(dot3(p, de), S.map_dx));
        if (i >= 0 && j >= 0 && i < S.npx && j < S.npy) atomicAdd(&S.out[i + L.idir * S.npx * S.npy + j * S.npx], delta);
        L.idir++;
        if (L.idir < S.ndir) { start_peel(S, L); return; }
        // all observers done: scatter and continue the random walk from the scattering point
        r.pos = L.kpos; r.level = L.klevel; r.ind = L.kind_; r.rho = L.krho; r.dir = L.kdir; r.tau = 0.0f;
        scatter_direction(r.dir, S.csc, S.bins, rng);
        L.free_path = -logf(rng.uniform());
        L.mode = (L.scat == 30) ? RAY_IDLE : RAY_MAIN;                     // MAX_SCATTERINGS, kernel_ASOC_sca.c:5
        return;
    }
    if (L.mode == RAY_FFS) {                                                // kernel_ASOC_sca.c:888-910 / 1720-1750
        r.tau += ds * rho0 * ksca;
        if (r.ind >= 0) return;
        float tau = r.tau, W;
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
        start_peel(S, L);
        return;
    }
    r.tau += dtau;
    if (r.ind < 0) L.mode = RAY_IDLE;
}

template <class RNG, bool OCT>
__device__ __forceinline__ void emit_packet(const ScaArgs &S, RNG &rng, int id, int III, Packet &pk) {
    if (S.kind == 0) emit_ps<ScaArgs, RNG, OCT>(S, rng, III, pk);
    else             emit_bg<ScaArgs, RNG, OCT>(S, rng, id, pk);
    fix_direction(pk.dir);
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
    if (S.kind == 1) have = have && id < 8LL * S.G.area;
    RngMwc rng;
    if (have) rng.seed(S.mwc, (unsigned long long)id);
    Lane L; L.mode = RAY_IDLE;
    Packet pk; pk.ind = -1; pk.level = 0; pk.rho = 0.0f;
    int III = 0;
    for (;;) {
        if (L.mode == RAY_IDLE && have) {
            if (III < S.batch) {
                emit_packet<RngMwc, OCT>(S, rng, (int)id, III, pk);
                III++; cnt.packets++;
                begin_packet<RngMwc, OCT>(S, L, rng, pk);
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

// persistent warps, one Philox stream per packet, idle lanes refilled from the work counter
template <bool OCT, bool DBL>
__global__ void __launch_bounds__(128) sca_stream_kernel(const __grid_constant__ ScaArgs S) {
    ScaCounters cnt = { 0, 0, 0, 0, 0 };
    const int lane = threadIdx.x & 31;
    const long long nlocal = (S.nunits - S.rank + S.world - 1) / S.world;
    RngPhilox rng;
    Lane L; L.mode = RAY_IDLE;
    Packet pk; pk.ind = -1; pk.level = 0; pk.rho = 0.0f;
    bool more = true;
    for (;;) {
        unsigned idle = __ballot_sync(FULL, L.mode == RAY_IDLE);
        if (idle == FULL || (__popc(idle) >= 8 && __any_sync(FULL, more))) {
            bool need = L.mode == RAY_IDLE && more;
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
                        rng.seed(S.phx, q);
                        emit_packet<RngPhilox, OCT>(S, rng, (int)(q / (unsigned)S.batch), (int)(q % (unsigned)S.batch), pk);
                        cnt.packets++;
                        begin_packet<RngPhilox, OCT>(S, L, rng, pk);
                    }
                }
            }
            if (!__any_sync(FULL, L.mode != RAY_IDLE || more)) break;
        }
        if (L.mode != RAY_IDLE) {
            advance<RngPhilox, OCT, DBL>(S, L, rng, cnt);
            if (++L.nstep > S.max_steps) { L.mode = RAY_IDLE; cnt.stuck++; }
        }
    }
    flush(S, cnt);
}

}  // namespace

void launch_sca(const ScaArgs &S, int rng_mode, int blocks, int threads, cudaStream_t stream) {
    const bool oct = S.G.levels > 1, dbl = S.G.dbl_sim != 0;
    if (rng_mode == SOC_RNG_REFERENCE) {
        if (!oct)      sca_item_kernel<false, false><<<blocks, threads, 0, stream>>>(S);
        else if (!dbl) sca_item_kernel<true, false><<<blocks, threads, 0, stream>>>(S);
        else           sca_item_kernel<true, true><<<blocks, threads, 0, stream>>>(S);
    } else {
        if (!oct)      sca_stream_kernel<false, false><<<blocks, threads, 0, stream>>>(S);
        else if (!dbl) sca_stream_kernel<true, false><<<blocks, threads, 0, stream>>>(S);
        else           sca_stream_kernel<true, true><<<blocks, threads, 0, stream>>>(S);
    }
}

int sca_blocks_per_sm(bool octree, bool dbl, int threads) {
    int n = 0;
    if (!octree)   cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sca_stream_kernel<false, false>, threads, 0);
    else if (!dbl) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sca_stream_kernel<true, false>, threads, 0);
    else           cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sca_stream_kernel<true, true>, threads, 0);
    return n > 0 ? n : 1;
}
