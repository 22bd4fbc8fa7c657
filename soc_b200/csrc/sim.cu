// soc_b200 -- Monte Carlo photon-packet kernels of the emission / absorption run.
//
// Replaces the reference kernels SimRAM_PB (point sources + isotropic background), SimRAM_HP (Healpix
// background) and SimRAM_CL (emission from the cells): kernel_ASOC.c:15-824, 831-1206, 1223-1684.
//
// Two launch layouts share every line of the physics:
//   * item kernel   (SOC_RNG_REFERENCE): one thread = one work item of the reference launch, seeded with the
//     reference's MWC64X stream for that work item and drawing random numbers in the reference's order.  With
//     the same inputs the absorption counters agree with the CPU oracle to float rounding -- this is the
//     parity layout.
//   * stream kernel (SOC_RNG_PACKET): persistent warps; one Philox stream per photon packet keyed by the
//     global packet number, idle lanes are refilled from a global work counter so that warps stay full, and
//     packet q is simulated by rank q % world -- results do not depend on the number of GPUs except for the
//     order of the float additions.  This is the production layout.
//
// Per cell-step the kernels touch DENS[cell] (4 B gather, carried in a register from the step that entered
// the cell) and TABS[cell] (+INT[cell]) through red.global.add.f32; nothing else leaves the SM.
#include <cstdio>
#include <cstdlib>
#include "sim.cuh"
#include "emit.cuh"
#include "walk.cuh"
#include "linkwalk.cuh"

#define FULL 0xffffffffu

namespace {

struct Counters { unsigned packets, steps, scat, stuck; };      // per thread; widened when flushed

// ---- accumulation ------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add(float *p, float v) { atomicAdd(p, v); }   // result unused -> RED.E.ADD.F32

// Sum `v` over the lanes in `peers` (all of which call this with the same set); returns the total on the
// lowest lane of the set.  log2(|peers|) shuffle rounds.
__device__ __forceinline__ float reduce_peers(unsigned peers, float v, int lane) {
    int rel = __popc(peers & ((1u << lane) - 1u));            // my position within the set
    unsigned higher = peers & (0xfffffffeu << lane);
    while (__any_sync(peers, higher != 0u)) {
        int next = __ffs(higher);
        float t = __shfl_sync(peers, v, (next - 1) & 31);
        if (next) v += t;
        bool done = rel & 1;
        higher &= __ballot_sync(peers, !done);
        rel >>= 1;
    }
    return v;
}

struct Deposit {
    const SimArgs &A;
    float *tile;        // shared-memory tile or nullptr
    __device__ __forceinline__ Deposit(const SimArgs &a, float *t) : A(a), tile(t) {}

    __device__ __forceinline__ void one(int oind, float delta, const vec3 &dir, int eidx, int level, int ind) const {
        if (A.with_ali && oind == eidx) {                                                     // kernel_ASOC.c:1486-1499
            red_add(&A.xab[oind], delta * A.tw);
            if (A.use_int && A.use_acc) red_add(&A.inten[oind], delta);      // INT takes every absorption (the fold only sees acc)
        } else {
            float v = A.use_acc ? delta : delta * A.tw * A.adhoc;
            float *main = A.use_acc ? A.acc : A.tabs;
            bool in_tile = false;
            if (tile != nullptr && level == 0 && (unsigned)(ind - A.tile_lo) < (unsigned)A.tile_span) {
                int ix = ind % A.G.nx, iy = (ind / A.G.nx) % A.G.ny, iz = ind / (A.G.nx * A.G.ny);
                unsigned tx = (unsigned)(ix - A.tile_x0), ty = (unsigned)(iy - A.tile_y0), tz = (unsigned)(iz - A.tile_z0);
                if (tx < SOC_TILE_N && ty < SOC_TILE_N && tz < SOC_TILE_N) {
                    atomicAdd(&tile[(tz * SOC_TILE_N + ty) * SOC_TILE_N + tx], v);
                    in_tile = true;
                }
            }
            if (!in_tile) red_add(&main[oind], v);
        }
        if (A.use_int && !A.use_acc) red_add(&A.inten[oind], delta);
        if (A.save_int2) {
            red_add(&A.intx[oind], delta * dir.x); red_add(&A.inty[oind], delta * dir.y); red_add(&A.intz[oind], delta * dir.z);
        }
    }

    // Called by every lane of the warp; `dep` says whether this lane has something to add.
    __device__ __forceinline__ void all(bool dep, int oind, float delta, const vec3 &dir, int eidx, int level, int ind,
                                        bool aggregate) const {
        if (!aggregate || A.save_int2 || A.with_ali) {
            if (dep) one(oind, delta, dir, eidx, level, ind);
            return;
        }
        unsigned act = __ballot_sync(FULL, dep);
        if (!dep) return;
        int lane = threadIdx.x & 31;
        unsigned peers = __match_any_sync(act, oind);
        if (peers != (1u << lane)) {
            delta = reduce_peers(peers, delta, lane);
            if (lane != __ffs(peers) - 1) return;
        }
        one(oind, delta, dir, eidx, level, ind);
    }
};

// ---- sampling --------------------------------------------------------------------------------------------
// free path incl. the optional step weighting, kernel_ASOC.c:516-541
template <class RNG>
__device__ __forceinline__ float sample_free_path(const SimArgs &A, RNG &rng, float &photons) {
    float fp;
    if (A.step_weight <= 0) fp = -logf(rng.uniform());
    else if (A.step_weight == 1) {
        fp = -logf(rng.uniform()) / A.sw_a;
        photons *= expf(A.sw_a * fp - fp) / A.sw_a;
    } else {
        float a = A.sw_a, b = A.sw_b;
        fp = -logf((-b + sqrtf(b * b + 4.0f * rng.uniform() * (1.0f - b))) / (2.0f - 2.0f * b)) / a;
        photons *= 1.0f / (a * b * expf((1.0f - a) * fp) + 2.0f * a * (1.0f - b) * expf((1.0f - 2.0f * a) * fp));
    }
    return fp;
}

// one packet of a point source / the background / the Healpix sky / the stored ROI field; false = nothing emitted
template <class RNG, bool OCT>
__device__ __forceinline__ bool emit_source(const SimArgs &A, RNG &rng, int id, int III, Packet &pk) {
    if (A.kind == SIM_PS)       emit_ps<SimArgs, RNG, OCT>(A, rng, III, pk);
    else if (A.kind == SIM_BG)  emit_bg<SimArgs, RNG, OCT>(A, rng, id, pk);
    else if (A.kind == SIM_ROI) return emit_roi<SimArgs, RNG, OCT>(A, rng, id, III, A.roi_nelem, pk);
    else                        emit_hp<SimArgs, RNG, OCT>(A, rng, pk);
    return true;
}
template <bool OCT>
__device__ __forceinline__ void init_roi(const SimArgs &A, Packet &pk) {          // kernel_ASOC.c:550, 1439
    if ((A.roi.flags & 2) && pk.ind >= 0) pk.roi = in_roi<OCT>(A.G, A.roi, pk.level, pk.ind) ? 1 : 0;
}

template <class RNG>
__device__ __forceinline__ void start_packet(const SimArgs &A, RNG &rng, Packet &pk, bool fixdir) {
    if (fixdir) fix_direction(pk.dir);
    pk.free_path = sample_free_path(A, rng, pk.photons);
    pk.tau = 0.0f; pk.scat = 0; pk.nstep = 0; pk.roi = 0;
}

// ---- one cell-step of the propagation loop (kernel_ASOC.c:556-820 / 1448-1679) -------------------------------
// Part 1: geometry + optical depths; decides between a full step and a scattering inside the cell and
// returns the energy to deposit.  Part 2 (finish_step) updates the packet.  The deposit sits between the
// two so that the whole warp reaches it together.
struct StepTmp { vec3 pos0; float rho0, tauA, dx; int ind0, level0, oind; bool scatter; };

template <bool CL, bool OCT, bool DBL>
__device__ __forceinline__ bool begin_step(const SimArgs &A, Packet &pk, StepTmp &t, float &delta, Counters &cnt) {
    const GridDesc &G = A.G;
    t.oind = OCT ? G.off[pk.level] + pk.ind : pk.ind;
    t.ind0 = pk.ind; t.level0 = pk.level; t.pos0 = pk.pos; t.rho0 = pk.rho;
    float ds = get_step<OCT, DBL, false>(G, pk.pos, pk.dir, pk.level, pk.ind, pk.rho);
    float kabs = A.kabs, ksca = A.ksca;
    if (A.with_abu) { float2 o = reinterpret_cast<const float2 *>(A.opt)[t.oind]; kabs = o.x; ksca = o.y; }
    t.tauA = ds * t.rho0 * kabs;
    float dtau = ds * t.rho0 * ksca;
    t.scatter = pk.free_path < (pk.tau + dtau);
    if (t.scatter) {
        pk.scat++;
        if (CL && pk.scat > 20) return false;                      // kernel_ASOC.c:1552-1556: dies before the deposit
        dtau = pk.free_path - pk.tau;
        t.dx = dtau / (ksca * t.rho0);
        t.tauA = t.dx * t.rho0 * kabs;
        cnt.scat++;
    } else {
        pk.tau += dtau;
    }
    delta = (t.tauA > SOC_TAULIM) ? (pk.photons * (1.0f - expf(-t.tauA))) : (pk.photons * t.tauA * (1.0f - 0.5f * t.tauA));
    cnt.steps++;
    return true;
}

template <class RNG, bool CL, bool OCT>
__device__ __forceinline__ bool finish_step(const SimArgs &A, Packet &pk, const StepTmp &t, RNG &rng) {
    pk.photons *= expf(-t.tauA);
    if (t.scatter) {
        float dx = OCT ? ldexpf(t.dx, t.level0) : t.dx;
        dx = fmaxf(0.0f, dx - 2.0f * SOC_PEPS);
        pk.pos.x = xadd(t.pos0.x, xmul(dx, pk.dir.x)); pk.pos.y = xadd(t.pos0.y, xmul(dx, pk.dir.y)); pk.pos.z = xadd(t.pos0.z, xmul(dx, pk.dir.z));
        pk.free_path = sample_free_path(A, rng, pk.photons);
        pk.ind = t.ind0; pk.level = t.level0; pk.rho = t.rho0;
        const float *csc = A.csc;
        if (A.with_msf)                                                // kernel_ASOC.c:777-794
            csc += A.bins * msf_pick(A.abu, A.scav, A.ndust, A.opt[2 * (size_t)t.oind + 1], t.oind, rng.uniform());
        scatter_direction(pk.dir, csc, A.bins, rng);
        pk.tau = 0.0f;
        if (!CL && pk.scat > 20) return false;                     // kernel_ASOC.c:801-804
        return true;
    }
    if (A.roi.flags & 2) {                                          // WITH_ROI_SAVE: kernel_ASOC.c:615-643, 1508-1535
        const bool r = pk.ind >= 0 && in_roi<OCT>(A.G, A.roi, pk.level, pk.ind);
        if (r && !pk.roi) {
            vec3 rp = pk.pos;
            if (OCT) root_position(A.G, rp, pk.level, pk.ind);
            roi_save_add(A.roi, rp, pk.dir, pk.photons);
        }
        pk.roi = r;
    }
    if (!CL && pk.level == t.level0 && pk.ind == t.ind0) {         // failed step: kernel_ASOC.c:649-665
        pk.pos.x = xadd(pk.pos.x, xmul(SOC_PEPS, pk.dir.x)); pk.pos.y = xadd(pk.pos.y, xmul(SOC_PEPS, pk.dir.y)); pk.pos.z = xadd(pk.pos.z, xmul(SOC_PEPS, pk.dir.z));
    }
    if (A.mirror && pk.ind < 0) mirror_literal<OCT>(A.G, A.mirror, pk.pos, pk.dir, pk.level, pk.ind, pk.rho);   // :686, 1064, 1540
    return pk.ind >= 0;
}

__device__ __forceinline__ void flush_counters(const SimArgs &A, const Counters &c) {
    warp_add_counter(A.counters + 0, c.packets);
    warp_add_counter(A.counters + 1, c.steps);
    warp_add_counter(A.counters + 2, c.scat);
    warp_add_counter(A.counters + 3, c.stuck);
}

__device__ __forceinline__ float *tile_begin(const SimArgs &A, float *smem) {
    if (A.deposit != DEP_TILE) return nullptr;
    for (int i = threadIdx.x; i < SOC_TILE_CELLS; i += blockDim.x) smem[i] = 0.0f;
    __syncthreads();
    return smem;
}
__device__ __forceinline__ void tile_end(const SimArgs &A, float *tile) {
    if (tile == nullptr) return;
    __syncthreads();
    const GridDesc &G = A.G;
    for (int i = threadIdx.x; i < SOC_TILE_CELLS; i += blockDim.x) {
        float v = tile[i];
        if (v != 0.0f) {
            int tx = i % SOC_TILE_N, ty = (i / SOC_TILE_N) % SOC_TILE_N, tz = i / (SOC_TILE_N * SOC_TILE_N);
            int ix = A.tile_x0 + tx, iy = A.tile_y0 + ty, iz = A.tile_z0 + tz;
            if (ix < G.nx && iy < G.ny && iz < G.nz) red_add(&(A.use_acc ? A.acc : A.tabs)[(iz * G.ny + iy) * G.nx + ix], v);
        }
    }
}

// =================================================================================================================
// Item kernel: thread <-> reference work item.
// =================================================================================================================
template <class RNG, bool OCT, bool DBL>
__global__ void __launch_bounds__(128) sim_item_kernel(const __grid_constant__ SimArgs A) {
    __shared__ float smem[SOC_TILE_CELLS];
    float *tile = tile_begin(A, smem);
    Deposit dep(A, OCT ? nullptr : tile);
    Counters cnt = { 0, 0, 0, 0 };
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long id = t * A.world + A.rank;
    bool have = id < A.nunits;
    if (A.kind == SIM_BG || A.kind == SIM_HP) have = have && id < 8LL * A.G.area;      // kernel_ASOC.c:96, 878
    if (A.kind == SIM_CL) have = have && id < A.G.cells;
    if (A.kind == SIM_ROI) have = have && id < 100LL * A.roi_nelem;                    // kernel_ASOC.c:99
    RNG rng;
    if (have) rng.seed_item(A, (unsigned long long)id);
    Packet pk; pk.ind = -1; pk.level = 0; pk.eidx = -1; pk.rho = 0.0f;
    bool alive = false;
    int III = 0, icell = (int)id - A.global, nray = 0;        // CL: icell advances by `global` per cell
    float pwei = 1.0f;
    const bool agg = A.deposit != DEP_RED;
    for (;;) {
        if (!alive && have) {
            // next packet of this work item
            if (A.kind == SIM_CL) {
                while (III >= nray) {
                    long long nc = (long long)icell + A.global;
                    if (nc >= A.G.cells) { have = false; break; }
                    icell = (int)nc; III = 0;
                    nray = cl_rays(A, icell, pwei);
                }
                if (have) {
                    III++;
                    emit_cl(A, rng, icell, pwei, pk);
                    start_packet(A, rng, pk, true);
                    init_roi<OCT>(A, pk);
                    cnt.packets++; alive = true;
                }
            } else if (III < A.batch) {
                if (emit_source<RNG, OCT>(A, rng, (int)id, III, pk)) {
                    start_packet(A, rng, pk, A.kind != SIM_HP);
                    init_roi<OCT>(A, pk);
                    cnt.packets++;
                    alive = pk.ind >= 0;
                }
                III++;
            } else have = false;
        }
        if (!__any_sync(FULL, alive || have)) break;
        StepTmp st; float delta = 0.0f; bool d = false;
        if (alive) {
            if (A.kind == SIM_CL) d = begin_step<true, OCT, DBL>(A, pk, st, delta, cnt);
            else                  d = begin_step<false, OCT, DBL>(A, pk, st, delta, cnt);
            if (!d) alive = false;
        }
        dep.all(d, st.oind, delta, pk.dir, pk.eidx, st.level0, st.ind0, agg);
        if (alive) {
            if (A.kind == SIM_CL) alive = finish_step<RNG, true, OCT>(A, pk, st, rng);
            else                  alive = finish_step<RNG, false, OCT>(A, pk, st, rng);
            if (++pk.nstep > A.max_steps) { alive = false; cnt.stuck++; }
        }
    }
    tile_end(A, tile);
    flush_counters(A, cnt);
}

// =================================================================================================================
// Stream kernel: persistent warps, one Philox stream per packet, idle lanes refilled from a work counter.
// Units: PS/BG/HP one packet; CL one cell (all its rays on one lane, ray number in the Philox counter).
// =================================================================================================================
template <bool OCT, bool DBL>
__global__ void __launch_bounds__(256) sim_stream_kernel(const __grid_constant__ SimArgs A) {
    __shared__ float smem[SOC_TILE_CELLS];
    float *tile = tile_begin(A, smem);
    Deposit dep(A, OCT ? nullptr : tile);
    Counters cnt = { 0, 0, 0, 0 };
    const int lane = threadIdx.x & 31;
    const long long nlocal = (A.nunits - A.rank + A.world - 1) / A.world;
    RngPhilox rng;
    Packet pk; pk.ind = -1; pk.level = 0; pk.eidx = -1; pk.rho = 0.0f;
    bool alive = false, more = true;
    int icell = 0, iray = 0, nray = 0;
    float pwei = 1.0f;
    const bool agg = A.deposit != DEP_RED;
    const int refill = A.refill;
    for (;;) {
        unsigned idle = __ballot_sync(FULL, !alive);
        if (idle == FULL || (__popc(idle) >= refill && __any_sync(FULL, more))) {
            // hand new units to the lanes that need one
            bool need = !alive && more && iray >= nray;
            unsigned nm = __ballot_sync(FULL, need);
            if (nm) {
                int leader = __ffs(nm) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(A.work, (unsigned long long)__popc(nm));
                base = __shfl_sync(FULL, base, leader);
                if (need) {
                    long long u = (long long)base + __popc(nm & ((1u << lane) - 1u));
                    if (u >= nlocal) more = false;
                    else {
                        unsigned long long q = (unsigned long long)u * A.world + A.rank;
                        if (A.kind == SIM_CL) {
                            icell = (int)q; iray = 0;
                            nray = cl_rays(A, icell, pwei);
                        } else {
                            rng.seed(A.phx, q);
                            int id = (int)(q / (unsigned)A.batch), III = (int)(q % (unsigned)A.batch);
                            if (emit_source<RngPhilox, OCT>(A, rng, id, III, pk)) {
                                start_packet(A, rng, pk, A.kind != SIM_HP);
                                init_roi<OCT>(A, pk);
                                cnt.packets++;
                                alive = pk.ind >= 0;
                            }
                        }
                    }
                }
            }
            if (A.kind == SIM_CL && !alive && iray < nray) {
                rng.seed(A.phx, (unsigned long long)(unsigned)icell | ((unsigned long long)(unsigned)iray << 32));
                iray++;
                emit_cl(A, rng, icell, pwei, pk);
                start_packet(A, rng, pk, true);
                init_roi<OCT>(A, pk);
                cnt.packets++; alive = true;
            }
            if (!__any_sync(FULL, alive || more || iray < nray)) break;
        }
        StepTmp st; float delta = 0.0f; bool d = false;
        if (alive) {
            if (A.kind == SIM_CL) d = begin_step<true, OCT, DBL>(A, pk, st, delta, cnt);
            else                  d = begin_step<false, OCT, DBL>(A, pk, st, delta, cnt);
            if (!d) alive = false;
        }
        dep.all(d, st.oind, delta, pk.dir, pk.eidx, st.level0, st.ind0, agg && __any_sync(FULL, d && pk.nstep < A.agg_steps));
        if (alive) {
            if (A.kind == SIM_CL) alive = finish_step<RngPhilox, true, OCT>(A, pk, st, rng);
            else                  alive = finish_step<RngPhilox, false, OCT>(A, pk, st, rng);
            if (++pk.nstep > A.max_steps) { alive = false; cnt.stuck++; }
        }
    }
    tile_end(A, tile);
    flush_counters(A, cnt);
}

// =================================================================================================================
// Fast kernel: the production path on regular grids (LEVELS == 1).
//
// Same physics and the same Philox packet streams as the stream kernel, but the cell-to-cell stepping is an
// incremental 3-D DDA written for the SM instead of a restatement of GetStep/Index:
//   * the packet carries (ix,iy,iz), the distances tx,ty,tz along the ray to the next x/y/z face and their
//     increments 1/|d|; one step is min3 + three subtractions + one integer increment -- no division, no
//     fmod, no float->int conversion, and exact geometry (no PEPS overshoot, so no "failed step" nudge);
//   * the density of the next cell is requested at the top of the step, before the optical-depth and exp()
//     work of the current cell, so the gather latency overlaps a full step of arithmetic;
//   * the only random numbers inside the loop are drawn at a scattering: one Philox block keyed by
//     (packet, scattering number) -- no generator state lives in registers across steps;
//   * one red.global.add.f32 per step into the scratch accumulator (folded into TABS / INT afterwards).
// The position inside the cell is implied by (tx,ty,tz): frac = d>0 ? 1-t|d| : t|d|.
// =================================================================================================================
struct FastPk {
    float tx, ty, tz, rdx, rdy, rdz;
    vec3 dir;
    int ix, iy, iz, ind;
    float rho, photons, free_path, tau;
    float2 opt;                      // per-cell (kabs, ksca) of the current cell (WITH_ABU)
    int scat, nstep, eidx;
    unsigned long long rid;          // Philox stream id of the packet
};


__device__ __forceinline__ float face_distance(float frac, float d, float rd) {
    frac = fminf(fmaxf(frac, 0.0f), 1.0f);
    return ((d > 0.0f) ? (1.0f - frac) : frac) * rd;
}

__device__ __forceinline__ void fast_from_packet(const SimArgs &A, const Packet &pk, FastPk &f) {
    const GridDesc &G = A.G;
    f.ix = clampi((int)floorf(pk.pos.x), 0, G.nx - 1);
    f.iy = clampi((int)floorf(pk.pos.y), 0, G.ny - 1);
    f.iz = clampi((int)floorf(pk.pos.z), 0, G.nz - 1);
    f.ind = (f.iz * G.ny + f.iy) * G.nx + f.ix;
    f.dir = pk.dir;
    f.rdx = 1.0f / fabsf(pk.dir.x); f.rdy = 1.0f / fabsf(pk.dir.y); f.rdz = 1.0f / fabsf(pk.dir.z);
    f.tx = face_distance(pk.pos.x - (float)f.ix, pk.dir.x, f.rdx);
    f.ty = face_distance(pk.pos.y - (float)f.iy, pk.dir.y, f.rdy);
    f.tz = face_distance(pk.pos.z - (float)f.iz, pk.dir.z, f.rdz);
    f.rho = pk.rho; f.photons = pk.photons; f.free_path = pk.free_path; f.tau = 0.0f;
    f.scat = 0; f.nstep = 0; f.eidx = pk.eidx;
    if (A.with_abu) f.opt = __ldg(reinterpret_cast<const float2 *>(A.opt) + f.ind);
}

// free path of the production kernels: plain -log(u) through the SFU unless step weighting is on
template <class RNG>
__device__ __forceinline__ float free_path_fast(const SimArgs &A, RNG &rng, float &photons) {
    if (A.step_weight <= 0) return -__logf(rng.uniform());
    return sample_free_path(A, rng, photons);
}

// ---- domain queues (QPk, sim.cuh) ---------------------------------------------------------------------------------------
__device__ __forceinline__ void q_store(QPk *p, const QPk &v) {
    const float4 *s = reinterpret_cast<const float4 *>(&v);
    float4 *d = reinterpret_cast<float4 *>(p);
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
}
__device__ __forceinline__ QPk q_load(const QPk *p) {
    QPk v;
    const float4 *s = reinterpret_cast<const float4 *>(p);
    float4 *d = reinterpret_cast<float4 *>(&v);
    d[0] = __ldcs(s); d[1] = __ldcs(s + 1); d[2] = __ldcs(s + 2); d[3] = __ldcs(s + 3);      // read once: streaming
    return v;
}
__device__ __forceinline__ QPk q_load_shared(const QPk *p) {
    QPk v;
    const float4 *s = reinterpret_cast<const float4 *>(p);
    float4 *d = reinterpret_cast<float4 *>(&v);
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
    return v;
}
// parks the packets of a warp in the queues of the domains that hold their cells (ix,iy,iz): every lane calls, `have` = this lane
// has a packet; one atomic per domain and warp instead of one per packet -- 3e7 adds to a single counter would serialise in the L2
__device__ __forceinline__ void q_push_warp(const SimArgs &A, const QPk &v, bool have) {
    const unsigned act = __ballot_sync(FULL, have);
    if (!have) return;
    const int lane = threadIdx.x & 31;
    const int d = ((v.iz / A.dsize[2]) * A.dsplit[1] + v.iy / A.dsize[1]) * A.dsplit[0] + v.ix / A.dsize[0];
    const unsigned peers = __match_any_sync(act, d);
    const int leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(A.q_tail + d, (unsigned)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    q_store(A.q_base + (size_t)d * (size_t)A.q_cap + base + __popc(peers & ((1u << lane) - 1u)), v);
}

// Parked packets are staged per warp in shared memory and leave in groups: the slot request of q_push_warp is an atomic
// WITH a return value, i.e. a full round trip to the L2 for the whole warp -- paid once per ~1.5 parked packets it was 18 %
// of the stall samples of the domain kernels at 512^3 (ncu, SHFL after the ATOMG); staged, it is paid once per Q_STAGE_N.
#ifndef Q_STAGE_N
#define Q_STAGE_N 16
#endif
__device__ __forceinline__ void q_stage_flush(const SimArgs &A, const QPk *stage, int &n) {
    __syncwarp();
    const int lane = threadIdx.x & 31;
    const bool have = lane < n;
    QPk v; v.ix = v.iy = v.iz = 0;
    if (have) v = q_load_shared(stage + lane);
    q_push_warp(A, v, have);
    __syncwarp();
    n = 0;
}
// every lane of the warp calls; `n` (entries staged so far) is warp-uniform
__device__ __forceinline__ void q_stage_push(const SimArgs &A, QPk *stage, int &n, const QPk &v, bool have) {
    const unsigned pm = __ballot_sync(FULL, have);
    const int k = __popc(pm);
    if (k > Q_STAGE_N) { q_push_warp(A, v, have); return; }
    if (n + k > Q_STAGE_N) q_stage_flush(A, stage, n);
    if (have) q_store(stage + n + __popc(pm & ((1u << (threadIdx.x & 31)) - 1u)), v);
    n += k;
}

// DEP: accumulation engine (DepositMode).  GENERAL = false drops the per-cell opacities, the intensity vector,
// the ALI split and the cell-emission source from the loop (uniform tests the common runs never take).
template <int DEP, bool GENERAL>
__global__ void __launch_bounds__(256, 4) sim_fast_kernel(const __grid_constant__ SimArgs A) {
    __shared__ float smem[DEP == DEP_TILE ? SOC_TILE_CELLS : 1];
    float *tile = (DEP == DEP_TILE) ? tile_begin(A, smem) : nullptr;
    const GridDesc &G = A.G;
    Counters cnt = { 0, 0, 0, 0 };
    const int lane = threadIdx.x & 31;
    // A.dom: clean-up pass of the domain mode -- the work units are packets parked in a queue (q_in[0 .. nlocal)), resumed
    // here on the whole grid in the reference's cell order; their absorptions go straight to TABS / INT
    const long long nlocal = A.dom ? A.nlocal : (A.nunits - A.rank + A.world - 1) / A.world;
    const int sy_ = G.nx, sz_ = G.nx * G.ny;
    const bool cl = GENERAL && A.kind == SIM_CL;
    const bool abu = GENERAL && A.with_abu;
    FastPk f; f.ind = 0; f.rid = 0; f.eidx = -1;
    bool alive = false, more = true;
    bool in_roi_now = false;         // WITH_ROI_SAVE: the packet is inside the region of interest
    bool wsc = false;                // the packet sits at a scattering point and waits for the batched scatter block
    int icell = 0, iray = 0, nray = 0;
    float pwei = 1.0f;
    const int refill = A.refill;
    for (;;) {
        unsigned idle = __ballot_sync(FULL, !alive);
        if (idle == FULL || (__popc(idle) >= refill && __any_sync(FULL, more))) {
            bool need = !alive && more && iray >= nray;
            unsigned nm = __ballot_sync(FULL, need);
            unsigned long long q = 0;
            bool got = false;
            if (nm) {
                int leader = __ffs(nm) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(A.work, (unsigned long long)__popc(nm));
                base = __shfl_sync(FULL, base, leader);
                if (need) {
                    long long u = (long long)base + __popc(nm & ((1u << lane) - 1u));
                    if (u >= nlocal) more = false;
                    else { q = A.dom ? (unsigned long long)u : (unsigned long long)u * A.world + A.rank; got = true; }
                }
            }
            if (got && cl) { icell = (int)q; iray = 0; nray = cl_rays(A, icell, pwei); got = false; }
            if (cl && !alive && iray < nray) {
                q = (unsigned long long)(unsigned)icell | ((unsigned long long)(unsigned)iray << 32);
                iray++; got = true;
            }
            if (got && A.dom) {                      // resume a parked packet
                int part = 0;                        // which queue the unit lies in (all non-empty queues in one launch)
                while (part + 1 < A.q_nparts && (long long)q >= A.q_part[part + 1]) part++;
                const QPk s = q_load(A.q_base + (size_t)part * (size_t)A.q_cap + (q - (unsigned long long)A.q_part[part]));
                f.tx = s.tx; f.ty = s.ty; f.tz = s.tz; f.rdx = s.rdx; f.rdy = s.rdy; f.rdz = s.rdz;
                const float adx = rcp_approx(s.rdx), ady = rcp_approx(s.rdy), adz = rcp_approx(s.rdz);
                f.dir.x = (s.upm & 1u) ? adx : -adx; f.dir.y = (s.upm & 2u) ? ady : -ady; f.dir.z = (s.upm & 4u) ? adz : -adz;
                f.ix = s.ix; f.iy = s.iy; f.iz = s.iz; f.ind = (s.iz * G.ny + s.iy) * G.nx + s.ix;
                f.rho = __ldg(G.dens + f.ind);
                f.photons = s.photons; f.free_path = s.free_path; f.tau = s.tau;
                f.scat = (int)(s.sn >> 24); f.nstep = (int)(s.sn & 0xffffffu); f.eidx = -1;
                f.rid = (unsigned long long)s.u * A.world + A.rank;
                if (A.with_abu) f.opt = __ldg(reinterpret_cast<const float2 *>(A.opt) + f.ind);
                alive = true;
            } else if (got) {
                RngPhilox rng; rng.seed(A.phx, q);
                Packet pk; pk.ind = -1; pk.level = 0; pk.eidx = -1; pk.rho = 0.0f;
                bool emitted = true;
                if (GENERAL && cl) emit_cl(A, rng, icell, pwei, pk);
                else emitted = emit_source<RngPhilox, false>(A, rng, (int)(q / (unsigned)A.batch), (int)(q % (unsigned)A.batch), pk);
                if (emitted) {
                    start_packet(A, rng, pk, A.kind != SIM_HP);
                    cnt.packets++;
                    alive = pk.ind >= 0;
                    if (alive) {
                        fast_from_packet(A, pk, f); f.rid = q;
                        if (GENERAL && (A.roi.flags & 2)) in_roi_now = in_roi_xyz(A.roi, f.ix, f.iy, f.iz);
                    }
                }
            }
            if (!__any_sync(FULL, alive || more || iray < nray)) break;
        }
        // ---- one cell-step ---------------------------------------------------------------------------------
        bool d = false; float delta = 0.0f; int oind = f.ind;
        bool sc = false; float rho_n = 0.0f; int nind = 0; bool inb = false; float tmin = 0.0f; int ax = 0;
        float2 opt_n = make_float2(0.0f, 0.0f);
        const bool was_alive = alive;
        // ---- scatterings, several lanes at a time: the block is long (Philox, table look-up, rotation) and only
        // ~1 lane in 30 needs it in a given iteration, so lanes wait until A.sc_batch of them are at a scattering
        // point (or nothing else can run) and then scatter together
        {
            const unsigned sm = __ballot_sync(FULL, alive && wsc);
            if (sm && (__popc(sm) >= A.sc_batch || !__any_sync(FULL, alive && !wsc))) {
                if (alive && wsc) {
                    // position inside the cell from the face distances, then a new direction
                    float fx = (f.dir.x > 0.0f) ? 1.0f - f.tx * fabsf(f.dir.x) : f.tx * fabsf(f.dir.x);
                    float fy = (f.dir.y > 0.0f) ? 1.0f - f.ty * fabsf(f.dir.y) : f.ty * fabsf(f.dir.y);
                    float fz = (f.dir.z > 0.0f) ? 1.0f - f.tz * fabsf(f.dir.z) : f.tz * fabsf(f.dir.z);
                    RngBlock rb(A.phx, f.rid, 0x10000u + (unsigned)f.scat);
                    f.free_path = free_path_fast(A, rb, f.photons);
                    const float u_ct = rb.uniform(), u_phi = rb.uniform();
                    const float *csc = A.csc;
                    if (GENERAL && A.with_msf) csc += A.bins * msf_pick(A.abu, A.scav, A.ndust, f.opt.y, f.ind, rb.uniform());
                    float ct = __ldg(csc + clampi((int)(u_ct * A.bins), 0, A.bins - 1));
                    scatter_rotate(f.dir, ct, SOC_TWOPI * u_phi);
                    f.rdx = __fdividef(1.0f, fabsf(f.dir.x)); f.rdy = __fdividef(1.0f, fabsf(f.dir.y)); f.rdz = __fdividef(1.0f, fabsf(f.dir.z));
                    f.tx = face_distance(fx, f.dir.x, f.rdx);
                    f.ty = face_distance(fy, f.dir.y, f.rdy);
                    f.tz = face_distance(fz, f.dir.z, f.rdz);
                    f.tau = 0.0f;
                    wsc = false;
                }
            }
        }
        const bool run = alive && !wsc;
        if (run) {
            // which face comes first, and the cell behind it; its density is requested right away
            tmin = fminf(f.tx, fminf(f.ty, f.tz));
            ax = (f.tx <= f.ty && f.tx <= f.tz) ? 0 : ((f.ty <= f.tz) ? 1 : 2);
            int c = (ax == 0) ? f.ix : ((ax == 1) ? f.iy : f.iz);
            float dd = (ax == 0) ? f.dir.x : ((ax == 1) ? f.dir.y : f.dir.z);
            int lim = (ax == 0) ? G.nx : ((ax == 1) ? G.ny : G.nz);
            int stride = (ax == 0) ? 1 : ((ax == 1) ? sy_ : sz_);
            int sgn = dd > 0.0f ? 1 : -1;
            inb = (unsigned)(c + sgn) < (unsigned)lim;
            nind = f.ind + sgn * stride;
            if (inb) {
                rho_n = __ldg(G.dens + nind);
                if (abu) opt_n = __ldg(reinterpret_cast<const float2 *>(A.opt) + nind);
            }
            float kabs = A.kabs, ksca = A.ksca;
            if (abu) { kabs = f.opt.x; ksca = f.opt.y; }
            float dtau = tmin * f.rho * ksca, tauA;
            sc = f.free_path < f.tau + dtau;
            d = true;
            if (sc) {
                f.scat++;
                if (cl && f.scat > 20) { d = false; alive = false; }
                tmin = fminf(tmin, (f.free_path - f.tau) / (ksca * f.rho));
            } else f.tau += dtau;
            tauA = tmin * f.rho * kabs;
            float e = expf(-tauA);
            delta = (tauA > SOC_TAULIM) ? (f.photons * (1.0f - e)) : (f.photons * tauA * (1.0f - 0.5f * tauA));
            if (d) { f.photons *= e; f.nstep++; }
        }
        // ---- deposit: one red.global.add.f32 (or a shared-memory tile / warp-combined add) ------------------
        if (GENERAL && (A.save_int2 || A.with_ali)) {
            if (d) {
                if (A.with_ali && oind == f.eidx) {          // kernel_ASOC.c:1486-1499: XAB instead of TABS, INT as ever
                    red_add(&A.xab[oind], delta * A.tw);
                    if (A.use_int) red_add(&A.inten[oind], delta);
                } else red_add(&A.acc[oind], delta);
                if (A.save_int2) {
                    red_add(&A.intx[oind], delta * f.dir.x); red_add(&A.inty[oind], delta * f.dir.y); red_add(&A.intz[oind], delta * f.dir.z);
                }
            }
        } else if (DEP == DEP_RED) {
            if (d) {
                if (A.dom) { red_add(&A.tabs[oind], delta * A.tw * A.adhoc); if (A.use_int) red_add(&A.inten[oind], delta); }
                else red_add(&A.acc[oind], delta);
            }
        } else {
            bool comb = __any_sync(FULL, d && f.nstep < A.agg_steps);
            if (comb) {
                unsigned act = __ballot_sync(FULL, d);
                if (d) {
                    unsigned peers = __match_any_sync(act, oind);
                    if (peers != (1u << lane)) {
                        delta = reduce_peers(peers, delta, lane);
                        if (lane != __ffs(peers) - 1) d = false;
                    }
                }
            }
            if (d) {
                bool in_tile = false;
                if (DEP == DEP_TILE && (unsigned)(oind - A.tile_lo) < (unsigned)A.tile_span) {
                    unsigned tx = (unsigned)(f.ix - A.tile_x0), ty = (unsigned)(f.iy - A.tile_y0), tz = (unsigned)(f.iz - A.tile_z0);
                    if (tx < SOC_TILE_N && ty < SOC_TILE_N && tz < SOC_TILE_N) {
                        atomicAdd(&tile[(tz * SOC_TILE_N + ty) * SOC_TILE_N + tx], delta);
                        in_tile = true;
                    }
                }
                if (!in_tile) red_add(&A.acc[oind], delta);
            }
        }
        if (run && alive) {
            f.tx -= tmin; f.ty -= tmin; f.tz -= tmin;
            if (sc) {
                wsc = true;
                if (!cl && f.scat > 20) { alive = false; wsc = false; }
            } else {
                const float da = (ax == 0) ? f.dir.x : ((ax == 1) ? f.dir.y : f.dir.z);
                if (!inb && (A.mirror & ((da > 0.0f ? 2 : 1) << (2 * ax)))) {
                    // reflecting border: same cell, the crossed component of the direction changes sign
                    if (ax == 0)      { f.dir.x = -f.dir.x; f.tx = f.rdx; }
                    else if (ax == 1) { f.dir.y = -f.dir.y; f.ty = f.rdy; }
                    else              { f.dir.z = -f.dir.z; f.tz = f.rdz; }
                } else {
                    if (ax == 0)      { f.ix += (f.dir.x > 0.0f) ? 1 : -1; f.tx = f.rdx; }
                    else if (ax == 1) { f.iy += (f.dir.y > 0.0f) ? 1 : -1; f.ty = f.rdy; }
                    else              { f.iz += (f.dir.z > 0.0f) ? 1 : -1; f.tz = f.rdz; }
                    f.ind = nind; f.rho = rho_n;
                    if (abu) f.opt = opt_n;
                    alive = inb;
                    if (GENERAL && (A.roi.flags & 2) && inb) {           // WITH_ROI_SAVE: kernel_ASOC.c:615-643
                        const bool r = in_roi_xyz(A.roi, f.ix, f.iy, f.iz);
                        if (r && !in_roi_now) {
                            // entry point on the face just crossed; the other coordinates from the face distances
                            const float ax_ = f.tx * fabsf(f.dir.x), ay_ = f.ty * fabsf(f.dir.y), az_ = f.tz * fabsf(f.dir.z);
                            vec3 rp = { (float)f.ix + ((f.dir.x > 0.0f) ? 1.0f - ax_ : ax_), (float)f.iy + ((f.dir.y > 0.0f) ? 1.0f - ay_ : ay_),
                                        (float)f.iz + ((f.dir.z > 0.0f) ? 1.0f - az_ : az_) };
                            roi_save_add(A.roi, rp, f.dir, f.photons);
                        }
                        in_roi_now = r;
                    }
                }
            }
            if (f.nstep > A.max_steps) { alive = false; wsc = false; cnt.stuck++; }
        }
        if (was_alive && !alive) { cnt.steps += f.nstep; cnt.scat += min(f.scat, 20); }     // packet finished
    }
    tile_end(A, tile);
    flush_counters(A, cnt);
}

// =================================================================================================================
// Lean kernel: the fast kernel for the common run (scalar opacities, TABS/INT only, point-source / background /
// Healpix packets), rewritten for the two limits ncu showed on the fast kernel -- instruction issue (~200 warp
// instructions per cell-step iteration) and the L1 -> L2 request interface (one RED plus 0.74 gather misses per
// step):
//   * per-axis state is (face distance, crossings left before the packet leaves the grid, index increment): the
//     step is FMNMX3 + two compares + selects, no coordinate arithmetic, no bounds compare against the dimensions;
//   * the absorbed fraction comes from ex2.approx (>= 0.01) or a 3-term series (< 0.01, error < 5e-8 relative);
//     deposit = P*d and P -= deposit, so the energy of a packet is conserved to rounding;
//   * scatter distance through rcp.approx, all of it branch-free; work counters live in shared memory and are
//     touched once per packet (they were spilled to local memory and updated every iteration);
//   * BRICK: DENS and the scratch accumulator are stored in 2x2x2 bricks = one 32-byte sector per brick, the cell
//     behind the next face is in the sector just touched with probability 1/2 for any direction (x-fastest order:
//     7/8 for steps along x, 0 otherwise = 0.29 on average).  The index increment per axis is one of two values
//     picked by the parity bit of the axis, which is a bit of the bricked index itself.
// Same Philox packet streams, same draw order and same physics as the fast kernel.
// =================================================================================================================
__host__ __device__ __forceinline__ int brick_index(int ix, int iy, int iz, int hx, int hy) {
    return ((((iz >> 1) * hy + (iy >> 1)) * hx + (ix >> 1)) << 3) | ((iz & 1) << 2) | ((iy & 1) << 1) | (ix & 1);
}
// Domain-major brick order: the grid is cut into dsplit[0] x dsplit[1] x dsplit[2] boxes of dsize[] cells (all of one size,
// even edges), each box is one contiguous run of 2x2x2 bricks -- a launch that works on one box touches one contiguous
// piece of DENS and of the accumulator (L2 sets, DRAM pages and the 256 MB reach of the TLB all see a small array).
// One box = the plain brick order.
__device__ __forceinline__ int layout_index(const SimArgs &A, int ix, int iy, int iz) {
    const int dx = ix / A.dsize[0], dy = iy / A.dsize[1], dz = iz / A.dsize[2];
    const int d = (dz * A.dsplit[1] + dy) * A.dsplit[0] + dx;
    return d * (A.dsize[0] * A.dsize[1] * A.dsize[2]) +
           brick_index(ix - dx * A.dsize[0], iy - dy * A.dsize[1], iz - dz * A.dsize[2], A.dsize[0] >> 1, A.dsize[1] >> 1);
}

// cell-steps per pass of the outer loop (refill and scattering checks once per pass): measured on the bench step, lean
// kernel (point-source launch) 64.1 ms with 1, 61.0 with 2, 66.5 with 4; look-ahead kernel (background) 58.0 / 57.5 / 56.7
#ifndef SOC_LEAN_REPS
#define SOC_LEAN_REPS 2
#endif
#ifndef SOC_AHEAD_REPS
#define SOC_AHEAD_REPS 3
#endif
template <bool BRICK>
struct LeanPk {
    float tx, ty, tz, rdx, rdy, rdz;     // distance to the next face per axis; 1/|d| (the direction itself is not kept:
                                         // |d_i| = 1/rd_i, sign from upm)
    int cx, cy, cz;                  // cell crossings left along each axis before the packet leaves the grid
    int ind;                         // index of the current cell in the layout of the density array
    int upm;                         // bit a set: the packet moves towards +axis a
    float rho, photons, free_path, tau;
    unsigned sn;                     // cell-steps taken (low 24 bits) and scatterings (high 8 bits) of the packet
    unsigned u;                      // work unit of this rank; the Philox stream id is u*world + rank
};
#define LEAN_STEPS(sn) ((sn) & 0xffffffu)
#define LEAN_SCAT(sn)  ((sn) >> 24)

// direction-dependent part of the state from the fractional position (fx,fy,fz) inside cell (ix,iy,iz)
// cells lo..hi (inclusive) per axis: the box the crossing counters refer to -- the grid, or the domain of the launch
struct Box { int lox, loy, loz, hix, hiy, hiz; };
template <bool DOM>
__device__ __forceinline__ Box launch_box(const SimArgs &A) {
    Box b;
    b.lox = DOM ? A.dom_lo[0] : 0; b.loy = DOM ? A.dom_lo[1] : 0; b.loz = DOM ? A.dom_lo[2] : 0;
    b.hix = DOM ? A.dom_hi[0] : A.G.nx - 1; b.hiy = DOM ? A.dom_hi[1] : A.G.ny - 1; b.hiz = DOM ? A.dom_hi[2] : A.G.nz - 1;
    return b;
}
// the box that holds cell (ix,iy,iz) and the first position of its bricks (domain-major layout, A.dsplit / A.dsize)
__device__ __forceinline__ void roam_box(const SimArgs &A, int ix, int iy, int iz, Box &b, int &dbase) {
    const int dx = ix / A.dsize[0], dy = iy / A.dsize[1], dz = iz / A.dsize[2];
    b.lox = dx * A.dsize[0]; b.loy = dy * A.dsize[1]; b.loz = dz * A.dsize[2];
    b.hix = b.lox + A.dsize[0] - 1; b.hiy = b.loy + A.dsize[1] - 1; b.hiz = b.loz + A.dsize[2] - 1;
    dbase = ((dz * A.dsplit[1] + dy) * A.dsplit[0] + dx) * (A.dsize[0] * A.dsize[1] * A.dsize[2]);
}
template <bool BRICK>
__device__ __forceinline__ void lean_set_direction(const Box &b, LeanPk<BRICK> &f, const vec3 &d, int ix, int iy, int iz,
                                                   float fx, float fy, float fz) {
    f.rdx = rcp_approx(fabsf(d.x)); f.rdy = rcp_approx(fabsf(d.y)); f.rdz = rcp_approx(fabsf(d.z));     // |d_i| >= DEPS/sqrt(3)
    f.tx = face_distance(fx, d.x, f.rdx); f.ty = face_distance(fy, d.y, f.rdy); f.tz = face_distance(fz, d.z, f.rdz);
    const bool ux = d.x > 0.0f, uy = d.y > 0.0f, uz = d.z > 0.0f;
    f.cx = ux ? b.hix - ix : ix - b.lox; f.cy = uy ? b.hiy - iy : iy - b.loy; f.cz = uz ? b.hiz - iz : iz - b.loz;
    f.upm = (ux ? 1 : 0) | (uy ? 2 : 0) | (uz ? 4 : 0);
}

// Order in which the work units of a launch are handed out (the packet a unit stands for keeps its Philox stream and its
// emission, so the results do not depend on it -- only which packets are in flight together does).  Background launches:
// the surface elements of a face are visited in 16 x 16 tiles instead of row by row, so that the ~1e5 packets in flight
// start from a compact patch of the surface rather than a strip across it and share more L1 / L2 sectors.
__device__ __forceinline__ unsigned long long unit_order(const SimArgs &A, unsigned long long u) {
    if (A.scramble > 1) return (u * A.scramble) % (unsigned long long)A.nlocal;          // experiment: scattered order
    if (A.scramble == 0 || A.kind != SIM_BG || A.world != 1) return u;
    const GridDesc &G = A.G;
    const unsigned batch = (unsigned)A.batch;
    const unsigned long long id = u / batch;
    const unsigned iii = (unsigned)(u - id * batch);
    const unsigned long long round = id / (unsigned)G.area;
    int e = (int)(id - round * (unsigned)G.area);
    const int nyz = G.ny * G.nz, nxz = G.nx * G.nz;
    int off, da;                                    // first element of the face, length of its rows
    if (e < 2 * nyz)            { off = (e < nyz) ? 0 : nyz; da = G.ny; }
    else if (e < 2 * nyz + 2 * nxz) { off = (e < 2 * nyz + nxz) ? 2 * nyz : 2 * nyz + nxz; da = G.nx; }
    else                        { off = (e < 2 * nyz + 2 * nxz + G.nx * G.ny) ? 2 * nyz + 2 * nxz : 2 * nyz + 2 * nxz + G.nx * G.ny; da = G.nx; }
    const int p = e - off, tpr = da >> 4;
    const int t = p >> 8, w = p & 255;
    const int a = ((t % tpr) << 4) | (w & 15), b = ((t / tpr) << 4) | (w >> 4);
    e = off + b * da + a;
    return ((round * (unsigned)G.area + (unsigned)e) * batch) + iii;
}

// work counters: 32-bit native shared-memory adds, spilled to the 64-bit global counter before they can wrap
__device__ __forceinline__ void count_add(unsigned *s32, unsigned long long *g64, unsigned v) {
    const unsigned old = atomicAdd(s32, v);
    if (old < 0x80000000u && old + v >= 0x80000000u) { atomicSub(s32, 0x80000000u); atomicAdd(g64, 0x80000000ull); }
}

template <int DEP, bool BRICK, bool PEND, bool DOM>
__global__ void __launch_bounds__(256, 4) sim_lean_kernel(const __grid_constant__ SimArgs A) {
    __shared__ float smem[DEP == DEP_TILE ? SOC_TILE_CELLS : 1];
    __shared__ float s_pend[PEND ? 4 * 256 : 1];
    int pend_h = -1;
    __shared__ unsigned s_cnt[4];
    __shared__ __align__(16) QPk s_stage[DOM ? 8 * Q_STAGE_N : 1];     // parked packets of each warp (q_stage_push)
    QPk *const stage = s_stage + (DOM ? (threadIdx.x >> 5) * Q_STAGE_N : 0);
    int nstage = 0;
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0u;
    float *tile = nullptr;
    if (DEP == DEP_TILE) tile = tile_begin(A, smem); else __syncthreads();
    const GridDesc &G = A.G;
    const float *__restrict__ dens = BRICK ? A.dens_brick : G.dens;
    const int lane = threadIdx.x & 31;
    const float kabs = A.kabs, ksca = A.ksca;
    const Box box = launch_box<DOM>(A);
    LeanPk<BRICK> f; f.ind = 0; f.u = 0; f.rho = 0.0f; f.sn = 0;
    bool alive = false, wsc = false;
    int tskip = 0;                   // DEP_TILE: the packet cannot be inside the tile during its next tskip steps
    bool more = true;                                        // warp-uniform: the work counter has not run out yet
    for (;;) {
        unsigned live = __ballot_sync(FULL, alive);
        if (more ? (32 - __popc(live) >= A.refill) : (live == 0u)) {
            if (!more) break;
            const unsigned nm = ~live;                       // lanes that take a new packet
            const int leader = __ffs(nm) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(A.work, (unsigned long long)__popc(nm));
            base = __shfl_sync(FULL, base, leader);
            if (!alive) {
                const unsigned long long u = base + __popc(nm & ((1u << lane) - 1u));
                if (DOM && u < (unsigned long long)A.nlocal) {        // a packet parked at the border of this domain
                    const QPk s = q_load(A.q_in + u);
                    alive = true; wsc = false; tskip = 0;
                    f.tx = s.tx; f.ty = s.ty; f.tz = s.tz; f.rdx = s.rdx; f.rdy = s.rdy; f.rdz = s.rdz;
                    f.photons = s.photons; f.free_path = s.free_path; f.tau = s.tau; f.sn = s.sn; f.u = s.u; f.upm = (int)s.upm;
                    f.cx = (s.upm & 1u) ? box.hix - s.ix : s.ix - box.lox; f.cy = (s.upm & 2u) ? box.hiy - s.iy : s.iy - box.loy;
                    f.cz = (s.upm & 4u) ? box.hiz - s.iz : s.iz - box.loz;
                    f.ind = A.dom_base + brick_index(s.ix - box.lox, s.iy - box.loy, s.iz - box.loz, A.dsize[0] >> 1, A.dsize[1] >> 1);
                    f.rho = __ldg(dens + f.ind);
                }
                if (!DOM && u < (unsigned long long)A.nlocal) {
                    const unsigned long long us = unit_order(A, u);
                    const unsigned long long q = us * A.world + A.rank;
                    RngPhilox rng; rng.seed(A.phx, q);
                    Packet pk; pk.ind = -1; pk.level = 0; pk.eidx = -1; pk.rho = 0.0f;
                    const int id = (int)(q / (unsigned)A.batch), III = (int)(q % (unsigned)A.batch);
                    if (A.kind == SIM_PS)      emit_ps<SimArgs, RngPhilox, false>(A, rng, III, pk);
                    else if (A.kind == SIM_BG) emit_bg<SimArgs, RngPhilox, false>(A, rng, id, pk);
                    else                       emit_hp<SimArgs, RngPhilox, false>(A, rng, pk);
                    start_packet(A, rng, pk, A.kind != SIM_HP);
                    if (pk.ind >= 0) {
                        alive = true; wsc = false; tskip = 0;
                        const int ix = clampi((int)floorf(pk.pos.x), 0, G.nx - 1), iy = clampi((int)floorf(pk.pos.y), 0, G.ny - 1),
                                  iz = clampi((int)floorf(pk.pos.z), 0, G.nz - 1);
                        lean_set_direction<BRICK>(box, f, pk.dir, ix, iy, iz, pk.pos.x - (float)ix, pk.pos.y - (float)iy, pk.pos.z - (float)iz);
                        f.ind = BRICK ? brick_index(ix, iy, iz, G.nx >> 1, G.ny >> 1) : (iz * G.ny + iy) * G.nx + ix;
                        f.rho = pk.rho; f.photons = pk.photons; f.free_path = pk.free_path; f.tau = 0.0f;
                        f.sn = 0; f.u = (unsigned)us;
                    }
                }
            }
            more = base + (unsigned long long)__popc(nm) < (unsigned long long)A.nlocal;
            if (!DOM && lane == leader) {                          // packets started by this warp (domain mode: counted at emission)
                const unsigned long long left = base < (unsigned long long)A.nlocal ? (unsigned long long)A.nlocal - base : 0ull;
                count_add(&s_cnt[0], A.counters + 0, (unsigned)min((unsigned long long)__popc(nm), left));
            }
            live = __ballot_sync(FULL, alive);
            if (live == 0u && !more) break;
        }
        // ---- scatterings, batched over the lanes of the warp (see sim_fast_kernel) ------------------------------
        {
            const unsigned sm = __ballot_sync(FULL, wsc);
            if (sm != 0u && (__popc(sm) >= A.sc_batch || sm == live)) {
                if (wsc) {
                    const bool ux = (f.upm & 1) != 0, uy = (f.upm & 2) != 0, uz = (f.upm & 4) != 0;
                    const float adx = rcp_approx(f.rdx), ady = rcp_approx(f.rdy), adz = rcp_approx(f.rdz);
                    const float ax = f.tx * adx, ay = f.ty * ady, az = f.tz * adz;
                    const float fx = ux ? 1.0f - ax : ax, fy = uy ? 1.0f - ay : ay, fz = uz ? 1.0f - az : az;
                    const int ix = ux ? box.hix - f.cx : box.lox + f.cx, iy = uy ? box.hiy - f.cy : box.loy + f.cy, iz = uz ? box.hiz - f.cz : box.loz + f.cz;
                    RngBlock rb(A.phx, (unsigned long long)f.u * A.world + A.rank, 0x10000u + LEAN_SCAT(f.sn));
                    f.free_path = free_path_fast(A, rb, f.photons);
                    const float ct = __ldg(A.csc + clampi((int)(rb.uniform() * A.bins), 0, A.bins - 1));
                    vec3 nd = { ux ? adx : -adx, uy ? ady : -ady, uz ? adz : -adz };
                    scatter_rotate(nd, ct, SOC_TWOPI * rb.uniform());
                    lean_set_direction<BRICK>(box, f, nd, ix, iy, iz, fx, fy, fz);
                    f.tau = 0.0f;
                    wsc = false;
                }
            }
        }
        #pragma unroll
        for (int rep = 0; rep < SOC_LEAN_REPS; rep++) {      // cell-steps per refill / scattering check
        // ---- one cell-step ---------------------------------------------------------------------------------------
        const bool run = alive && !wsc;
        float delta = 0.0f, tmin = 0.0f, rho_n = 0.0f;
        int nind = 0;
        bool px = false, py = false, inb = false, sc = false;
        bool parked = false;             // DOM: the packet went to the queue of the next domain
        const int oind = f.ind;
        if (run) {
            tmin = fminf(f.tx, fminf(f.ty, f.tz));
            px = f.tx == tmin; py = !px && (f.ty == tmin);
            // index increment: x-fastest order +-(1, nx, nx*ny); bricks: +-(1,2,4) inside the brick (the parity bit of the
            // axis, a bit of the index itself, tells on which side of its brick the cell lies), else to the next brick
            const int abit = px ? 1 : (py ? 2 : 4);
            // (written with the y/z stride as a value of its own: nvcc 12.9 folds `px ? 7 : (py ? by : bz)` to
            //  `px ? 7 : bz` once the kernel also holds pz = !px && !py -- seen in the PTX, so keep this form)
            const int far_yz = BRICK ? (py ? A.brick_by : A.brick_bz) : (py ? G.nx : A.slab_xy);
            const int far = px ? (BRICK ? 7 : 1) : far_yz;
            const int mag = (BRICK && ((f.ind ^ f.upm) & abit) != 0) ? abit : far;
            const int step = (f.upm & abit) ? mag : -mag;
            const int crem = px ? f.cx : (py ? f.cy : f.cz);
            nind = f.ind + step;
            inb = crem > 0;
            if (inb) rho_n = __ldg(dens + nind);
            const float krho = ksca * f.rho;
            const float tend = fmaf(tmin, krho, f.tau);
            sc = f.free_path < tend;
            const float tsc = (f.free_path - f.tau) * rcp_approx(krho);
            if (sc) { tmin = fminf(tmin, tsc); f.sn += 1u << 24; } else f.tau = tend;
            const float x = tmin * f.rho * kabs;
            const float e = exp2f_approx(-1.4426950408889634f * x);
            const float ser = x * fmaf(x, fmaf(x, 0.16666667f, -0.5f), 1.0f);
            const float dfrac = (x < 0.01f) ? ser : (1.0f - e);
            delta = f.photons * dfrac;
            f.photons -= delta;
            f.sn++;
        }
        // ---- deposit ----------------------------------------------------------------------------------------------
        bool d = run;
        if (DEP != DEP_RED) {
            if (__any_sync(FULL, d && LEAN_STEPS(f.sn) < (unsigned)A.agg_steps)) {
                const unsigned act = __ballot_sync(FULL, d);
                if (d) {
                    const unsigned peers = __match_any_sync(act, oind);
                    if (peers != (1u << lane)) {
                        delta = reduce_peers(peers, delta, lane);
                        if (lane != __ffs(peers) - 1) d = false;
                    }
                }
            }
        }
        if (d) {
            bool in_tile = false;
            if (DEP == DEP_TILE && tskip == 0) {
                // coordinates at the time of the deposit: the crossing counters have not been updated yet.  A cell at
                // Chebyshev distance D from the tile cannot be followed by a tile cell within the next D - 1 steps (one
                // step moves one cell along one axis), so the test is skipped that long.
                const int ix = (f.upm & 1) ? box.hix - f.cx : box.lox + f.cx, iy = (f.upm & 2) ? box.hiy - f.cy : box.loy + f.cy,
                          iz = (f.upm & 4) ? box.hiz - f.cz : box.loz + f.cz;
                const int ux = ix - A.tile_x0, uy = iy - A.tile_y0, uz = iz - A.tile_z0;
                const int tfar = max(max(max(-ux, ux - (SOC_TILE_N - 1)), max(-uy, uy - (SOC_TILE_N - 1))), max(-uz, uz - (SOC_TILE_N - 1)));
                if (tfar <= 0) {
                    atomicAdd(&tile[(uz * SOC_TILE_N + uy) * SOC_TILE_N + ux], delta);
                    in_tile = true;
                } else tskip = tfar;                             // decremented below: tfar - 1 steps without the test
            }
            if (!in_tile) {
                if (PEND) {
                    // deposits of consecutive steps that fall into the same aligned group of four cells (half a brick)
                    // are summed in a per-lane shared-memory slot and leave the SM as one red.global.add.v4.f32
                    const int h = oind >> 2, k = oind & 3;
                    float *slot = s_pend + threadIdx.x;
                    // branch-free: read the slot, send it off if the group changes, write back old-or-zero plus delta
                    const bool same = h == pend_h;
                    float v0 = slot[0], v1 = slot[256], v2 = slot[512], v3 = slot[768];
                    if (!same && pend_h >= 0) atomicAdd(reinterpret_cast<float4 *>(A.acc) + pend_h, make_float4(v0, v1, v2, v3));
                    v0 = same ? v0 : 0.0f; v1 = same ? v1 : 0.0f; v2 = same ? v2 : 0.0f; v3 = same ? v3 : 0.0f;
                    slot[0] = v0 + ((k == 0) ? delta : 0.0f); slot[256] = v1 + ((k == 1) ? delta : 0.0f);
                    slot[512] = v2 + ((k == 2) ? delta : 0.0f); slot[768] = v3 + ((k == 3) ? delta : 0.0f);
                    pend_h = h;
                } else red_add(&A.acc[oind], delta);
            }
        }
        if (run) {
            f.tx -= tmin; f.ty -= tmin; f.tz -= tmin;
            if (DEP == DEP_TILE && tskip > 0) tskip--;
            if (sc) {
                wsc = true;
                if (LEAN_SCAT(f.sn) > 20u) { alive = false; wsc = false; }
            } else {
                const int abit_ = px ? 1 : (py ? 2 : 4);
                if (A.mirror != 0 && !inb && (A.mirror & (px ? 3 : (py ? 12 : 48)) & ((f.upm & abit_) ? 42 : 21))) {
                    // reflecting border: same cell, the packet turns around on this axis
                    f.upm ^= abit_;
                    if (px)      { f.cx = G.nx - 1; f.tx = f.rdx; }
                    else if (py) { f.cy = G.ny - 1; f.ty = f.rdy; }
                    else         { f.cz = G.nz - 1; f.tz = f.rdz; }
                } else {                                         // branch-free: the three axes are selects, not code paths
                    const bool pz = !px && !py;
                    f.tx = px ? f.rdx : f.tx; f.ty = py ? f.rdy : f.ty; f.tz = pz ? f.rdz : f.tz;
                    f.cx -= px; f.cy -= py; f.cz -= pz;
                    f.ind = nind; f.rho = rho_n;
                    alive = inb;
                    if (DOM && !inb) {
                        // left the domain: through a face of the grid the packet is gone, through an interior face it is parked
                        // for the domain it enters (complete DDA state; the counter of the crossed axis stands at -1 = one
                        // cell beyond the border)
                        const int face = (px ? 0 : (py ? 2 : 4)) + ((f.upm & abit_) ? 1 : 0);
                        if (!((A.dom_faces >> face) & 1)) parked = true;       // stored at the end of the step, the warp together
                    }
                }
            }
            bool stuck = false;
            if (LEAN_STEPS(f.sn) > (unsigned)A.max_steps && !parked) { alive = false; wsc = false; stuck = true; }
            if (!alive && !parked) {                            // packet finished: once per packet
                count_add(&s_cnt[1], A.counters + 1, LEAN_STEPS(f.sn));
                count_add(&s_cnt[2], A.counters + 2, min(LEAN_SCAT(f.sn), 20u));
                if (stuck) count_add(&s_cnt[3], A.counters + 3, 1u);
            }
        }
        if (DOM && __any_sync(FULL, parked)) {                  // one queue slot request per target domain and warp
            QPk o;
            o.tx = f.tx; o.ty = f.ty; o.tz = f.tz; o.rdx = f.rdx; o.rdy = f.rdy; o.rdz = f.rdz;
            o.photons = f.photons; o.free_path = f.free_path; o.tau = f.tau;
            o.ix = (f.upm & 1) ? box.hix - f.cx : box.lox + f.cx; o.iy = (f.upm & 2) ? box.hiy - f.cy : box.loy + f.cy;
            o.iz = (f.upm & 4) ? box.hiz - f.cz : box.loz + f.cz;
            o.upm = (unsigned)f.upm; o.sn = f.sn; o.u = f.u; o.pad = 0u;
            q_stage_push(A, stage, nstage, o, parked);
        }
        }
    }
    if (DOM) q_stage_flush(A, stage, nstage);
    if (PEND && pend_h >= 0) {
        const float *slot = s_pend + threadIdx.x;
        atomicAdd(reinterpret_cast<float4 *>(A.acc) + pend_h, make_float4(slot[0], slot[256], slot[512], slot[768]));
    }
    if (DEP == DEP_TILE) {
        // un-brick the tile flush: tile_end() adds into A.acc with x-fastest indices
        __syncthreads();
        for (int i = threadIdx.x; i < SOC_TILE_CELLS; i += blockDim.x) {
            const float v = tile[i];
            if (v != 0.0f) {
                const int ux = i % SOC_TILE_N, uy = (i / SOC_TILE_N) % SOC_TILE_N, uz = i / (SOC_TILE_N * SOC_TILE_N);
                const int ix = A.tile_x0 + ux, iy = A.tile_y0 + uy, iz = A.tile_z0 + uz;
                if (ix < G.nx && iy < G.ny && iz < G.nz)
                    red_add(&A.acc[BRICK ? layout_index(A, ix, iy, iz) : (iz * G.ny + iy) * G.nx + ix], v);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 4 && s_cnt[threadIdx.x]) atomicAdd(A.counters + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}

// =================================================================================================================
// Look-ahead kernel: the lean kernel with the geometry running one cell ahead of the physics.
// ncu on the lean kernel: half of all stall samples sit on the one instruction that receives the density of the
// next cell -- the gather is issued and consumed within the same iteration (~45 instructions apart), so every
// iteration pays an L2 (or DRAM) round trip that only the other 7 warps of the scheduler can cover.  The DDA does
// not depend on the density, so here
//   * the geometry state (face distances, crossing counters, index) describes the entry point of cell A, the cell
//     AFTER the one the physics works on; every iteration advances it by one cell and requests the density of the
//     cell behind A with cp.async (LDGSTS: global -> a two-slot per-lane ring in shared memory, no register
//     scoreboard, no register move that would force the wait);
//   * the physics of cell i uses (rho, seg, ind) in registers; at the end of the iteration the density of A --
//     requested one full iteration earlier -- is picked up from the ring after cp.async.wait_group 1;
//   * a packet that scatters in cell i steps its geometry back by the unused part of the segment (one crossing to
//     undo, the axis is remembered) and re-primes: one geometry-only iteration per emission / scattering;
//   * the axis update is branch-free (the lean kernel ran three divergent paths, 35 issue slots per iteration).
// Same packets, same Philox draws and same physics as the lean kernel; results differ by float rounding of the
// stepped-back scattering positions only.  Reflecting borders (MIRROR) stay with the lean kernel.
// =================================================================================================================
__device__ __forceinline__ void cp_async_f32(unsigned saddr, const float *g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ float lds_f32(unsigned saddr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory"); return v; }
__device__ __forceinline__ void cp_async_f32x2(unsigned saddr, const float2 *g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ float2 lds_f32x2(unsigned saddr) {
    float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr) : "memory"); return v;
}

// per-lane state word of the look-ahead kernel
#define AH_UPM     0x7u        // bit a: the packet moves towards +axis a
#define AH_AXA     0x70u       // axis bit (<<4) of the crossing from the physics cell into A
#define AH_ALIVE   0x100u
#define AH_WSC     0x200u      // waits at a scattering point
#define AH_SLOT    0x400u      // ring slot the next density request writes (byte offset 1024 = one row of the ring)
#define AH_PRIMED  0x800u      // (rho, seg, ind) of the physics cell are valid
#define AH_AIN     0x1000u     // A lies inside the grid

// KAPPA: per-cell opacities (WITH_ABU).  The kernel then reads ONE float2 per cell, (kabs*n, ksca*n) in the layout of the
// density array (kappa_kernel builds it from DENS and OPT before the launch), instead of the density and the two
// opacities: 8 B gathered per step instead of 12, and the ring slots are 8 bytes wide.
#define AH_SLOT_SHIFT(KAPPA) ((KAPPA) ? 1 : 0)
// ROAM (with DOM): the clean-up pass of the domain mode.  The launch takes the packets of ALL queues (q_part) and nothing is
// parked: a packet that leaves its box through an interior face is re-homed on the spot -- box bounds, crossing counters and brick
// index of the box it enters, exactly what parking and picking it up again would have produced -- so the last few packets of a
// launch (ping-pong between boxes after many scatterings) finish on the bricked layout instead of the general kernel's x-fastest one.
template <int DEP, bool BRICK, int CTAS, bool KAPPA, bool DOM, bool ROAM = false>
__global__ void __launch_bounds__(256, CTAS) sim_ahead_kernel(const __grid_constant__ SimArgs A) {
    __shared__ float smem[DEP == DEP_TILE ? SOC_TILE_CELLS : 1];
    __shared__ __align__(8) float s_ring[(KAPPA ? 4 : 2) * 256];   // slot s of lane t: s_ring[s * 256 + t] (float or float2 units)
    __shared__ unsigned s_cnt[4];
    __shared__ __align__(16) QPk s_stage[DOM ? 8 * Q_STAGE_N : 1];     // parked packets of each warp (q_stage_push)
    QPk *const stage = s_stage + (DOM ? (threadIdx.x >> 5) * Q_STAGE_N : 0);
    int nstage = 0;
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0u;
    float *tile = nullptr;
    if (DEP == DEP_TILE) tile = tile_begin(A, smem); else __syncthreads();
    const GridDesc &G = A.G;
    const float *__restrict__ dens = BRICK ? A.dens_brick : G.dens;
    const float2 *__restrict__ kappa = A.kappa;
    const int lane = threadIdx.x & 31;
    const float kabs = A.kabs, ksca = A.ksca;
    const unsigned ring = (unsigned)__cvta_generic_to_shared(&s_ring[KAPPA ? 2 * threadIdx.x : threadIdx.x]);
    Box box = launch_box<DOM>(A);    // ROAM: the box of this lane's packet
    int dbase = A.dom_base;          // first position of that box in the bricked arrays
    float ka = 0.0f;                 // KAPPA: kabs*n of the physics cell (f.rho holds ksca*n)
    LeanPk<BRICK> f; f.ind = 0; f.u = 0; f.rho = 0.0f; f.sn = 0; f.upm = 0;        // f.ind = index of cell A
    int ind = 0;                     // cell the physics works on
    float seg = 0.0f;                // path length through it (at a scattering: the part not used)
    unsigned st = 0u;                // AH_* bits
    int tskip = 0;                   // DEP_TILE: the packet cannot be inside the tile during its next tskip steps
    bool more = true;
    for (;;) {
        unsigned live = __ballot_sync(FULL, (st & AH_ALIVE) != 0u);
        if (more ? (32 - __popc(live) >= A.refill) : (live == 0u)) {
            if (!more) break;
            const unsigned nm = ~live;
            const int leader = __ffs(nm) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(A.work, (unsigned long long)__popc(nm));
            base = __shfl_sync(FULL, base, leader);
            if (!(st & AH_ALIVE)) {
                const unsigned long long u = base + __popc(nm & ((1u << lane) - 1u));
                if (DOM && u < (unsigned long long)A.nlocal) {        // a packet parked at the border of this domain: not primed
                    const QPk *src = A.q_in + u;
                    if (ROAM) {                                       // the queues back to back in unit space
                        int part = 0;
                        while (part + 1 < A.q_nparts && (long long)u >= A.q_part[part + 1]) part++;
                        src = A.q_base + (size_t)part * (size_t)A.q_cap + (u - (unsigned long long)A.q_part[part]);
                    }
                    const QPk s = q_load(src);
                    if (ROAM) roam_box(A, s.ix, s.iy, s.iz, box, dbase);
                    f.tx = s.tx; f.ty = s.ty; f.tz = s.tz; f.rdx = s.rdx; f.rdy = s.rdy; f.rdz = s.rdz;
                    f.photons = s.photons; f.free_path = s.free_path; f.tau = s.tau; f.sn = s.sn; f.u = s.u; f.upm = (int)s.upm;
                    f.cx = (s.upm & 1u) ? box.hix - s.ix : s.ix - box.lox; f.cy = (s.upm & 2u) ? box.hiy - s.iy : s.iy - box.loy;
                    f.cz = (s.upm & 4u) ? box.hiz - s.iz : s.iz - box.loz;
                    f.ind = dbase + brick_index(s.ix - box.lox, s.iy - box.loy, s.iz - box.loz, A.dsize[0] >> 1, A.dsize[1] >> 1);
                    ind = f.ind;
                    st = (st & AH_SLOT) | AH_ALIVE | s.upm;
                    tskip = 0;
                    if (KAPPA) { const float2 k2 = __ldg(kappa + f.ind); ka = k2.x; f.rho = k2.y; }
                    else f.rho = __ldg(dens + f.ind);
                }
                if (!DOM && u < (unsigned long long)A.nlocal) {
                    const unsigned long long us = unit_order(A, u);
                    const unsigned long long q = us * A.world + A.rank;
                    RngPhilox rng; rng.seed(A.phx, q);
                    Packet pk; pk.ind = -1; pk.level = 0; pk.eidx = -1; pk.rho = 0.0f;
                    const int id = (int)(q / (unsigned)A.batch), III = (int)(q % (unsigned)A.batch);
                    if (A.kind == SIM_PS)      emit_ps<SimArgs, RngPhilox, false>(A, rng, III, pk);
                    else if (A.kind == SIM_BG) emit_bg<SimArgs, RngPhilox, false>(A, rng, id, pk);
                    else                       emit_hp<SimArgs, RngPhilox, false>(A, rng, pk);
                    start_packet(A, rng, pk, A.kind != SIM_HP);
                    if (pk.ind >= 0) {
                        const int ix = clampi((int)floorf(pk.pos.x), 0, G.nx - 1), iy = clampi((int)floorf(pk.pos.y), 0, G.ny - 1),
                                  iz = clampi((int)floorf(pk.pos.z), 0, G.nz - 1);
                        lean_set_direction<BRICK>(box, f, pk.dir, ix, iy, iz, pk.pos.x - (float)ix, pk.pos.y - (float)iy, pk.pos.z - (float)iz);
                        st = (st & AH_SLOT) | AH_ALIVE | (unsigned)f.upm;
                        tskip = 0;
                        f.ind = BRICK ? brick_index(ix, iy, iz, G.nx >> 1, G.ny >> 1) : (iz * G.ny + iy) * G.nx + ix;
                        ind = f.ind;
                        f.rho = pk.rho; f.photons = pk.photons; f.free_path = pk.free_path; f.tau = 0.0f;
                        if (KAPPA) { const float2 k2 = __ldg(kappa + f.ind); ka = k2.x; f.rho = k2.y; }
                        f.sn = 0; f.u = (unsigned)us;
                    }
                }
            }
            more = base + (unsigned long long)__popc(nm) < (unsigned long long)A.nlocal;
            if (!DOM && lane == leader) {
                const unsigned long long left = base < (unsigned long long)A.nlocal ? (unsigned long long)A.nlocal - base : 0ull;
                count_add(&s_cnt[0], A.counters + 0, (unsigned)min((unsigned long long)__popc(nm), left));
            }
            live = __ballot_sync(FULL, (st & AH_ALIVE) != 0u);
            if (live == 0u && !more) break;
        }
        // ---- scatterings, batched over the lanes of the warp ------------------------------------------------------
        {
            const unsigned sm = __ballot_sync(FULL, (st & AH_WSC) != 0u);
            if (sm != 0u && (__popc(sm) >= A.sc_batch || sm == live)) {
                if (st & AH_WSC) {
                    // the geometry stands at the entry of A: step back by `seg` (the unused part of the segment) into
                    // the physics cell -- every face is that much further away, the face just crossed is `seg` ahead
                    asm volatile("cp.async.wait_all;" ::: "memory");      // the abandoned request of A must not land later
                    const bool ax = (st & 0x10u) != 0u, ay = (st & 0x20u) != 0u, az = (st & 0x40u) != 0u;
                    f.tx = ax ? seg : f.tx + seg; f.ty = ay ? seg : f.ty + seg; f.tz = az ? seg : f.tz + seg;
                    f.cx += ax; f.cy += ay; f.cz += az;
                    const bool ux = (st & 1u) != 0u, uy = (st & 2u) != 0u, uz = (st & 4u) != 0u;
                    const float adx = rcp_approx(f.rdx), ady = rcp_approx(f.rdy), adz = rcp_approx(f.rdz);
                    const float px_ = f.tx * adx, py_ = f.ty * ady, pz_ = f.tz * adz;
                    const float fx = ux ? 1.0f - px_ : px_, fy = uy ? 1.0f - py_ : py_, fz = uz ? 1.0f - pz_ : pz_;
                    const int ix = ux ? box.hix - f.cx : box.lox + f.cx, iy = uy ? box.hiy - f.cy : box.loy + f.cy, iz = uz ? box.hiz - f.cz : box.loz + f.cz;
                    RngBlock rb(A.phx, (unsigned long long)f.u * A.world + A.rank, 0x10000u + LEAN_SCAT(f.sn));
                    f.free_path = free_path_fast(A, rb, f.photons);
                    const float ct = __ldg(A.csc + clampi((int)(rb.uniform() * A.bins), 0, A.bins - 1));
                    vec3 nd = { ux ? adx : -adx, uy ? ady : -ady, uz ? adz : -adz };
                    scatter_rotate(nd, ct, SOC_TWOPI * rb.uniform());
                    lean_set_direction<BRICK>(box, f, nd, ix, iy, iz, fx, fy, fz);
                    st = (st & AH_SLOT) | AH_ALIVE | (unsigned)f.upm;          // not primed, not waiting
                    f.ind = ind;
                    f.tau = 0.0f;
                }
            }
        }
        #pragma unroll
        for (int rep = 0; rep < SOC_AHEAD_REPS; rep++) {      // cell-steps per refill / scattering check
        // ---- one iteration: physics of cell `ind`, geometry from A to the cell behind it --------------------------
        const bool run = (st & (AH_ALIVE | AH_WSC)) == AH_ALIVE;
        const bool phys = (st & (AH_ALIVE | AH_WSC | AH_PRIMED)) == (AH_ALIVE | AH_PRIMED);
        float delta = 0.0f, len = seg;
        bool sc = false;
        bool parked = false;             // DOM: the packet went to the queue of the next domain
        if (phys) {
            const float krho = KAPPA ? f.rho : ksca * f.rho;
            const float tend = fmaf(seg, krho, f.tau);
            sc = f.free_path < tend;
            if (__builtin_expect(sc, 0)) {           // rare: the length to the scattering point is worked out only here
                const float tsc = (f.free_path - f.tau) * rcp_approx(krho);
                len = fminf(seg, tsc); f.sn += 1u << 24;
            } else f.tau = tend;
            const float x = KAPPA ? len * ka : len * f.rho * kabs;
            const float e = exp2f_approx(-1.4426950408889634f * x);
            const float ser = x * fmaf(x, fmaf(x, 0.16666667f, -0.5f), 1.0f);
            const float dfrac = (x < 0.01f) ? ser : (1.0f - e);
            delta = f.photons * dfrac;
            f.photons -= delta;
            f.sn++;
        }
        // ---- deposit ----------------------------------------------------------------------------------------------
        bool d = phys;
        if (DEP != DEP_RED) {
            if (__any_sync(FULL, d && LEAN_STEPS(f.sn) < (unsigned)A.agg_steps)) {
                const unsigned act = __ballot_sync(FULL, d);
                if (d) {
                    const unsigned peers = __match_any_sync(act, ind);
                    if (peers != (1u << lane)) {
                        delta = reduce_peers(peers, delta, lane);
                        if (lane != __ffs(peers) - 1) d = false;
                    }
                }
            }
        }
        if (d) {
            bool in_tile = false;
            if (DEP == DEP_TILE && tskip == 0) {
                // coordinates of the physics cell: the counters belong to A, one crossing (axis AH_AXA) further on.  A cell
                // at Chebyshev distance D from the tile is not followed by a tile cell within D - 1 steps: test skipped
                const int kx = f.cx + ((st >> 4) & 1u), ky = f.cy + ((st >> 5) & 1u), kz = f.cz + ((st >> 6) & 1u);
                const int ix = (st & 1u) ? box.hix - kx : box.lox + kx, iy = (st & 2u) ? box.hiy - ky : box.loy + ky,
                          iz = (st & 4u) ? box.hiz - kz : box.loz + kz;
                const int ux = ix - A.tile_x0, uy = iy - A.tile_y0, uz = iz - A.tile_z0;
                const int tfar = max(max(max(-ux, ux - (SOC_TILE_N - 1)), max(-uy, uy - (SOC_TILE_N - 1))), max(-uz, uz - (SOC_TILE_N - 1)));
                if (tfar <= 0) {
                    atomicAdd(&tile[(uz * SOC_TILE_N + uy) * SOC_TILE_N + ux], delta);
                    in_tile = true;
                } else tskip = tfar;
            }
            if (!in_tile) red_add(&A.acc[ind], delta);
        }
        if (DEP == DEP_TILE && phys && tskip > 0) tskip--;
        if (run) {
            if (sc) {
                st |= AH_WSC;
                seg -= len;                                      // distance back from the entry of A to the scattering point
                if (LEAN_SCAT(f.sn) > 20u) st &= ~(AH_ALIVE | AH_WSC);
            } else if ((st & (AH_PRIMED | AH_AIN)) == AH_PRIMED) {
                st &= ~AH_ALIVE;                                 // the physics cell was the last one inside the grid (DOM: the domain)
                if (DOM) {
                    // the geometry stands at the entry of A, one cell beyond the border (the counter of the crossed axis is -1):
                    // through an interior face the packet is parked for the domain that holds A, not primed
                    const unsigned axb = (st >> 4) & 7u;
                    const int face = ((axb & 1u) ? 0 : ((axb & 2u) ? 2 : 4)) + ((st & axb) ? 1 : 0);
                    if (ROAM) {                                      // is the face crossed a face of the grid?
                        const bool upw = (st & axb) != 0u;
                        const int lo = (axb & 1u) ? box.lox : ((axb & 2u) ? box.loy : box.loz), hi = (axb & 1u) ? box.hix : ((axb & 2u) ? box.hiy : box.hiz);
                        const int dimv = (axb & 1u) ? G.nx : ((axb & 2u) ? G.ny : G.nz);
                        if (!(upw ? hi == dimv - 1 : lo == 0)) parked = true;  // re-homed below
                    } else if (!((A.dom_faces >> face) & 1)) parked = true;    // stored at the end of the step, the warp together
                }
            } else {
                // geometry: from the entry of A (unprimed: from the packet's position in its cell) to the next face
                const float tmin = fminf(f.tx, fminf(f.ty, f.tz));
                const bool px = f.tx == tmin, py = !px && (f.ty == tmin), pz = !px && !py;
                const unsigned abit = px ? 1u : (py ? 2u : 4u);
                const int far_yz = BRICK ? (py ? A.brick_by : A.brick_bz) : (py ? G.nx : A.slab_xy);
                const int far = px ? (BRICK ? 7 : 1) : far_yz;
                const int mag = (BRICK && (((unsigned)f.ind ^ st) & abit) != 0u) ? (int)abit : far;
                const int nind = f.ind + ((st & abit) ? mag : -mag);
                const int crem = px ? f.cx : (py ? f.cy : f.cz);
                const bool inb = crem > 0;
                const unsigned wslot = ring + ((st & AH_SLOT) << AH_SLOT_SHIFT(KAPPA));
                if (inb) { if (KAPPA) cp_async_f32x2(wslot, kappa + nind); else cp_async_f32(wslot, dens + nind); }
                cp_async_commit();
                f.tx = px ? f.rdx : f.tx - tmin; f.ty = py ? f.rdy : f.ty - tmin; f.tz = pz ? f.rdz : f.tz - tmin;
                f.cx -= px; f.cy -= py; f.cz -= pz;
                if (st & AH_PRIMED) {                            // density of A, requested one iteration ago
                    cp_async_wait1();
                    const unsigned rslot = ring + (((st & AH_SLOT) ^ AH_SLOT) << AH_SLOT_SHIFT(KAPPA));
                    if (KAPPA) { const float2 k2 = lds_f32x2(rslot); ka = k2.x; f.rho = k2.y; }
                    else f.rho = lds_f32(rslot);
                }
                ind = f.ind; f.ind = nind; seg = tmin;
                st = ((st & ~(AH_AXA | AH_AIN)) ^ AH_SLOT) | AH_PRIMED | (abit << 4) | (inb ? AH_AIN : 0u);
            }
            bool stuck = false;
            // guard against packets that never leave: looked at once per pass of the outer loop
            if (rep == 0 && LEAN_STEPS(f.sn) > (unsigned)A.max_steps && !parked) { st &= ~(AH_ALIVE | AH_WSC); stuck = true; }
            if (!(st & AH_ALIVE) && !parked) {                  // packet finished: once per packet
                count_add(&s_cnt[1], A.counters + 1, LEAN_STEPS(f.sn));
                count_add(&s_cnt[2], A.counters + 2, min(LEAN_SCAT(f.sn), 20u));
                if (stuck) count_add(&s_cnt[3], A.counters + 3, 1u);
            }
        }
        if (ROAM) {
            if (parked) {
                // the geometry stands at the entry of A, one cell into the next box: that box becomes the lane's box
                const int ix = (st & 1u) ? box.hix - f.cx : box.lox + f.cx, iy = (st & 2u) ? box.hiy - f.cy : box.loy + f.cy,
                          iz = (st & 4u) ? box.hiz - f.cz : box.loz + f.cz;
                roam_box(A, ix, iy, iz, box, dbase);
                f.cx = (st & 1u) ? box.hix - ix : ix - box.lox; f.cy = (st & 2u) ? box.hiy - iy : iy - box.loy;
                f.cz = (st & 4u) ? box.hiz - iz : iz - box.loz;
                f.ind = dbase + brick_index(ix - box.lox, iy - box.loy, iz - box.loz, A.dsize[0] >> 1, A.dsize[1] >> 1);
                ind = f.ind;
                st = (st & (AH_SLOT | AH_UPM)) | AH_ALIVE;           // as picked up from a queue: not primed
                if (KAPPA) { const float2 k2 = __ldg(kappa + f.ind); ka = k2.x; f.rho = k2.y; }
                else f.rho = __ldg(dens + f.ind);
            }
        } else if (DOM && __any_sync(FULL, parked)) {           // one queue slot request per target domain and warp
            QPk o;
            o.tx = f.tx; o.ty = f.ty; o.tz = f.tz; o.rdx = f.rdx; o.rdy = f.rdy; o.rdz = f.rdz;
            o.photons = f.photons; o.free_path = f.free_path; o.tau = f.tau;
            o.ix = (st & 1u) ? box.hix - f.cx : box.lox + f.cx; o.iy = (st & 2u) ? box.hiy - f.cy : box.loy + f.cy;
            o.iz = (st & 4u) ? box.hiz - f.cz : box.loz + f.cz;
            o.upm = st & AH_UPM; o.sn = f.sn; o.u = f.u; o.pad = 0u;
            q_stage_push(A, stage, nstage, o, parked);
        }
        }
    }
    if (DOM && !ROAM) q_stage_flush(A, stage, nstage);
    if (DEP == DEP_TILE) {
        __syncthreads();
        for (int i = threadIdx.x; i < SOC_TILE_CELLS; i += blockDim.x) {
            const float v = tile[i];
            if (v != 0.0f) {
                const int ux = i % SOC_TILE_N, uy = (i / SOC_TILE_N) % SOC_TILE_N, uz = i / (SOC_TILE_N * SOC_TILE_N);
                const int ix = A.tile_x0 + ux, iy = A.tile_y0 + uy, iz = A.tile_z0 + uz;
                if (ix < G.nx && iy < G.ny && iz < G.nz)
                    red_add(&A.acc[BRICK ? layout_index(A, ix, iy, iz) : (iz * G.ny + iy) * G.nx + ix], v);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 4 && s_cnt[threadIdx.x]) atomicAdd(A.counters + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}

// =================================================================================================================
// Walk kernel: the production path on octree clouds (LEVELS > 1).  Same persistent-warp / refill / Philox /
// scratch-accumulator design as the fast kernel; the stepping is the incremental octree walk of walk.cuh, taken
// one hop (climb / cross / descend) per loop iteration so that lanes with long climbs do not idle the warp.
// =================================================================================================================
template <int DEP, bool GENERAL>
__global__ void __launch_bounds__(256, 4) sim_walk_kernel(const __grid_constant__ SimArgs A) {
    const GridDesc &G = A.G;
    Counters cnt = { 0, 0, 0, 0 };
    const int lane = threadIdx.x & 31;
    const long long nlocal = (A.nunits - A.rank + A.world - 1) / A.world;
    const bool cl = GENERAL && A.kind == SIM_CL;
    const bool abu = GENERAL && A.with_abu;
    Walker w; w.ind = -1; w.level = 0;
    int phase = WALK_LEAF, ax = 0;
    bool in_roi_now = false;         // WITH_ROI_SAVE: the packet is inside the region of interest
    float photons = 0.0f, free_path = 0.0f, tau = 0.0f;
    int scat = 0, nstep = 0, eidx = -1;
    unsigned long long rid = 0;
    bool alive = false, more = true;
    int icell = 0, iray = 0, nray = 0;
    float pwei = 1.0f;
    const int refill = A.refill;
    for (;;) {
        unsigned idle = __ballot_sync(FULL, !alive);
        if (idle == FULL || (__popc(idle) >= refill && __any_sync(FULL, more))) {
            bool need = !alive && more && iray >= nray;
            unsigned nm = __ballot_sync(FULL, need);
            unsigned long long q = 0;
            bool got = false;
            if (nm) {
                int leader = __ffs(nm) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(A.work, (unsigned long long)__popc(nm));
                base = __shfl_sync(FULL, base, leader);
                if (need) {
                    long long u = (long long)base + __popc(nm & ((1u << lane) - 1u));
                    if (u >= nlocal) more = false;
                    else { q = (unsigned long long)u * A.world + A.rank; got = true; }
                }
            }
            if (got && cl) { icell = (int)q; iray = 0; nray = cl_rays(A, icell, pwei); got = false; }
            if (cl && !alive && iray < nray) {
                q = (unsigned long long)(unsigned)icell | ((unsigned long long)(unsigned)iray << 32);
                iray++; got = true;
            }
            if (got) {
                RngPhilox rng; rng.seed(A.phx, q);
                Packet pk; pk.ind = -1; pk.level = 0; pk.eidx = -1; pk.rho = 0.0f;
                bool emitted = true;
                if (GENERAL && cl) emit_cl(A, rng, icell, pwei, pk);
                else emitted = emit_source<RngPhilox, true>(A, rng, (int)(q / (unsigned)A.batch), (int)(q % (unsigned)A.batch), pk);
                if (emitted) {
                    start_packet(A, rng, pk, A.kind != SIM_HP);
                    cnt.packets++;
                    alive = pk.ind >= 0;
                }
                if (alive) {
                    walker_init<true>(G, w, pk.pos, pk.dir, pk.level, pk.ind, pk.rho);
                    photons = pk.photons; free_path = pk.free_path; tau = 0.0f; scat = 0; nstep = 0; eidx = pk.eidx; rid = q;
                    phase = WALK_LEAF;
                    if (GENERAL && (A.roi.flags & 2)) in_roi_now = in_roi_xyz(A.roi, w.ix, w.iy, w.iz);
                }
            }
            if (!__any_sync(FULL, alive || more || iray < nray)) break;
        }
        // scatterings, several lanes at a time (see sim_fast_kernel)
        {
            const unsigned sm = __ballot_sync(FULL, alive && phase == WALK_SCATTER);
            if (sm && (__popc(sm) >= A.sc_batch || !__any_sync(FULL, alive && phase != WALK_SCATTER))) {
                if (alive && phase == WALK_SCATTER) {
                    float fx, fy, fz;
                    walker_fraction(w, fx, fy, fz);
                    RngBlock rb(A.phx, rid, 0x10000u + (unsigned)scat);
                    free_path = free_path_fast(A, rb, photons);
                    const float u_ct = rb.uniform(), u_phi = rb.uniform();
                    const float *csc = A.csc;
                    if (GENERAL && A.with_msf) {
                        const int oc = G.off[w.level] + w.ind;
                        csc += A.bins * msf_pick(A.abu, A.scav, A.ndust, __ldg(A.opt + 2 * (size_t)oc + 1), oc, rb.uniform());
                    }
                    float ct = __ldg(csc + clampi((int)(u_ct * A.bins), 0, A.bins - 1));
                    vec3 nd = w.d;
                    scatter_rotate(nd, ct, SOC_TWOPI * u_phi);
                    walker_set_direction(w, nd, fx, fy, fz);
                    tau = 0.0f;
                    phase = WALK_LEAF;
                }
            }
        }
        bool d = false, sc = false;
        float delta = 0.0f, ds = 0.0f, tmin = 0.0f;
        int oind = 0;
        const bool was_alive = alive;
        const bool ready = alive && phase == WALK_LEAF;
        if (ready) {
            // physics of the current leaf
            oind = G.off[w.level] + w.ind;
            tmin = fminf(w.tx, fminf(w.ty, w.tz));
            ax = (w.tx <= w.ty && w.tx <= w.tz) ? 0 : ((w.ty <= w.tz) ? 1 : 2);
            ds = fmaxf(tmin, 0.0f);
            float kabs = A.kabs, ksca = A.ksca;
            if (abu) { float2 o = __ldg(reinterpret_cast<const float2 *>(A.opt) + oind); kabs = o.x; ksca = o.y; }
            const float dtau = ds * w.rho * ksca;
            sc = free_path < tau + dtau;
            d = true;
            if (sc) {
                scat++;
                if (cl && scat > 20) { d = false; alive = false; }
                ds = fminf(ds, (free_path - tau) / (ksca * w.rho));
            } else tau += dtau;
            const float tauA = ds * w.rho * kabs;
            const float e = expf(-tauA);
            delta = (tauA > SOC_TAULIM) ? (photons * (1.0f - e)) : (photons * tauA * (1.0f - 0.5f * tauA));
            if (d) { photons *= e; nstep++; }
        }
        if (GENERAL && (A.save_int2 || A.with_ali)) {
            if (d) {
                if (A.with_ali && oind == eidx) {            // kernel_ASOC.c:1486-1499: XAB instead of TABS, INT as ever
                    red_add(&A.xab[oind], delta * A.tw);
                    if (A.use_int) red_add(&A.inten[oind], delta);
                } else red_add(&A.acc[oind], delta);
                if (A.save_int2) {
                    red_add(&A.intx[oind], delta * w.d.x); red_add(&A.inty[oind], delta * w.d.y); red_add(&A.intz[oind], delta * w.d.z);
                }
            }
        } else if (DEP == DEP_RED) {
            if (d) red_add(&A.acc[oind], delta);
        } else {
            if (__any_sync(FULL, d && nstep < A.agg_steps)) {
                unsigned act = __ballot_sync(FULL, d);
                if (d) {
                    unsigned peers = __match_any_sync(act, oind);
                    if (peers != (1u << lane)) {
                        delta = reduce_peers(peers, delta, lane);
                        if (lane != __ffs(peers) - 1) d = false;
                    }
                }
            }
            if (d) red_add(&A.acc[oind], delta);
        }
        if (ready && alive) {
            if (sc) {
                w.tx -= ds; w.ty -= ds; w.tz -= ds;
                phase = WALK_SCATTER;
                if (!cl && scat > 20) { alive = false; phase = WALK_LEAF; }
            } else {
                w.tx -= tmin; w.ty -= tmin; w.tz -= tmin;
                phase = WALK_CROSS;
            }
            if (nstep > A.max_steps) { alive = false; phase = WALK_LEAF; cnt.stuck++; }
        }
        // navigation, one hop of each kind per iteration: climb, try to cross, descend
        #pragma unroll 1
        for (int hop = 0; hop < A.nav_hops; hop++) {
            if (alive && phase == WALK_CLIMB) { nav_climb(G, w, ax); phase = WALK_CROSS; }
            if (alive && phase == WALK_CROSS) {
                const bool at_root = GENERAL && (A.roi.flags & 2) && w.level == 0;
                phase = nav_cross(G, w, ax, A.mirror);
                if (w.ind < 0) alive = false;
                else if (at_root && phase != WALK_CLIMB) {           // a new root cell: WITH_ROI_SAVE, kernel_ASOC.c:615-643
                    const bool r = in_roi_xyz(A.roi, w.ix, w.iy, w.iz);
                    if (r && !in_roi_now) {
                        const float ax_ = w.tx * fabsf(w.d.x), ay_ = w.ty * fabsf(w.d.y), az_ = w.tz * fabsf(w.d.z);
                        vec3 rp = { (float)w.ix + ((w.d.x > 0.0f) ? 1.0f - ax_ : ax_), (float)w.iy + ((w.d.y > 0.0f) ? 1.0f - ay_ : ay_),
                                    (float)w.iz + ((w.d.z > 0.0f) ? 1.0f - az_ : az_) };
                        roi_save_add(A.roi, rp, w.d, photons);
                    }
                    in_roi_now = r;
                }
            }
            if (alive && phase == WALK_DESCEND) phase = nav_descend(G, w, ax);
        }
        if (was_alive && !alive) { cnt.steps += nstep; cnt.scat += min(scat, 20); }
    }
    flush_counters(A, cnt);
}

#ifndef SOC_LINK_REPS
#define SOC_LINK_REPS 2
#endif
template <int DEP, bool GENERAL>
__global__ void __launch_bounds__(256, 4) sim_link_kernel(const __grid_constant__ SimArgs A) {
    const GridDesc &G = A.G;
    // work counters: shared memory, touched once per packet (8 registers of 64-bit counters were spilled in the loop)
    __shared__ unsigned s_cnt[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long nlocal = (A.nunits - A.rank + A.world - 1) / A.world;
    const bool cl = GENERAL && A.kind == SIM_CL;
    const bool abu = GENERAL && A.with_abu;
    LWalker w; w.cell = -1; w.level = 0;
    const int *__restrict__ nbr = A.nbr;
    int phase = WALK_LEAF, ax = 0;
    bool in_roi_now = false;         // WITH_ROI_SAVE: the packet is inside the region of interest
    float photons = 0.0f, free_path = 0.0f, tau = 0.0f;
    int scat = 0, nstep = 0, eidx = -1;
    unsigned long long rid = 0;
    bool alive = false, more = true;
    int icell = 0, iray = 0, nray = 0;
    float pwei = 1.0f;
    const int refill = A.refill;
    for (;;) {
        unsigned idle = __ballot_sync(FULL, !alive);
        if (idle == FULL || (__popc(idle) >= refill && __any_sync(FULL, more))) {
            bool need = !alive && more && iray >= nray;
            unsigned nm = __ballot_sync(FULL, need);
            unsigned long long q = 0;
            bool got = false;
            if (nm) {
                int leader = __ffs(nm) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(A.work, (unsigned long long)__popc(nm));
                base = __shfl_sync(FULL, base, leader);
                if (need) {
                    long long u = (long long)base + __popc(nm & ((1u << lane) - 1u));
                    if (u >= nlocal) more = false;
                    else { q = (unsigned long long)u * A.world + A.rank; got = true; }
                }
            }
            if (got && cl) { icell = (int)q; iray = 0; nray = cl_rays(A, icell, pwei); got = false; }
            if (cl && !alive && iray < nray) {
                q = (unsigned long long)(unsigned)icell | ((unsigned long long)(unsigned)iray << 32);
                iray++; got = true;
            }
            if (got) {
                RngPhilox rng; rng.seed(A.phx, q);
                Packet pk; pk.ind = -1; pk.level = 0; pk.eidx = -1; pk.rho = 0.0f;
                bool emitted = true;
                if (GENERAL && cl) emit_cl(A, rng, icell, pwei, pk);
                else emitted = emit_source<RngPhilox, true>(A, rng, (int)(q / (unsigned)A.batch), (int)(q % (unsigned)A.batch), pk);
                if (emitted) {
                    start_packet(A, rng, pk, A.kind != SIM_HP);
                    count_add(&s_cnt[0], A.counters + 0, 1u);
                    alive = pk.ind >= 0;
                }
                if (alive) {
                    lw_init(G, w, pk.pos, pk.dir, pk.level, pk.ind, pk.rho);
                    photons = pk.photons; free_path = pk.free_path; tau = 0.0f; scat = 0; nstep = 0; eidx = pk.eidx; rid = q;
                    phase = WALK_LEAF;
                    if (GENERAL && (A.roi.flags & 2)) in_roi_now = in_roi_xyz(A.roi, w.cx >> w.level, w.cy >> w.level, w.cz >> w.level);
                }
            }
            if (!__any_sync(FULL, alive || more || iray < nray)) break;
        }
        // scatterings, several lanes at a time (see sim_fast_kernel)
        {
            const unsigned sm = __ballot_sync(FULL, alive && phase == WALK_SCATTER);
            if (sm && (__popc(sm) >= A.sc_batch || !__any_sync(FULL, alive && phase != WALK_SCATTER))) {
                if (alive && phase == WALK_SCATTER) {
                    float fx, fy, fz;
                    lw_fraction(w, fx, fy, fz);
                    RngBlock rb(A.phx, rid, 0x10000u + (unsigned)scat);
                    free_path = free_path_fast(A, rb, photons);
                    const float u_ct = rb.uniform(), u_phi = rb.uniform();
                    const float *csc = A.csc;
                    if (GENERAL && A.with_msf) {
                        const int oc = w.cell;
                        csc += A.bins * msf_pick(A.abu, A.scav, A.ndust, __ldg(A.opt + 2 * (size_t)oc + 1), oc, rb.uniform());
                    }
                    float ct = __ldg(csc + clampi((int)(u_ct * A.bins), 0, A.bins - 1));
                    vec3 nd = lw_dir(w);
                    scatter_rotate(nd, ct, SOC_TWOPI * u_phi);
                    lw_set_direction(w, nd, fx, fy, fz);
                    tau = 0.0f;
                    phase = WALK_LEAF;
                }
            }
        }
        #pragma unroll
        for (int rep = 0; rep < SOC_LINK_REPS; rep++) {          // cell-steps / navigation rounds per refill and scattering check
        bool d = false, sc = false;
        float delta = 0.0f, ds = 0.0f, tmin = 0.0f;
        int oind = 0;
        const bool was_alive = alive;
        const bool ready = alive && phase == WALK_LEAF;
        int2 e2 = make_int2(0, 0);
        if (ready) {
            // physics of the current leaf
            oind = w.cell;
            tmin = fminf(w.tx, fminf(w.ty, w.tz));
            ax = (w.tx <= w.ty && w.tx <= w.tz) ? 0 : ((w.ty <= w.tz) ? 1 : 2);
            // the exit face is known: the table entry behind it is requested now and used after the physics (it is not needed
            // when the packet scatters in this cell: rare)
            e2 = lw_entry(nbr, w, ax);
            ds = fmaxf(tmin, 0.0f);
            float kabs = A.kabs, ksca = A.ksca;
            if (abu) { float2 o = __ldg(reinterpret_cast<const float2 *>(A.opt) + oind); kabs = o.x; ksca = o.y; }
            const float dtau = ds * w.rho * ksca;
            sc = free_path < tau + dtau;
            d = true;
            if (sc) {
                scat++;
                if (cl && scat > 20) { d = false; alive = false; }
                ds = fminf(ds, (free_path - tau) * rcp_approx(ksca * w.rho));
            } else tau += dtau;
            const float tauA = ds * w.rho * kabs;
            const float e = expf(-tauA);
            delta = (tauA > SOC_TAULIM) ? (photons * (1.0f - e)) : (photons * tauA * (1.0f - 0.5f * tauA));
            if (d) { photons *= e; nstep++; }
        }
        if (GENERAL && (A.save_int2 || A.with_ali)) {
            if (d) {
                if (A.with_ali && oind == eidx) {            // kernel_ASOC.c:1486-1499: XAB instead of TABS, INT as ever
                    red_add(&A.xab[oind], delta * A.tw);
                    if (A.use_int) red_add(&A.inten[oind], delta);
                } else red_add(&A.acc[oind], delta);
                if (A.save_int2) {
                    const vec3 wd = lw_dir(w);
                    red_add(&A.intx[oind], delta * wd.x); red_add(&A.inty[oind], delta * wd.y); red_add(&A.intz[oind], delta * wd.z);
                }
            }
        } else if (DEP == DEP_RED) {
            if (d) red_add(&A.acc[oind], delta);
        } else {
            if (__any_sync(FULL, d && nstep < A.agg_steps)) {
                unsigned act = __ballot_sync(FULL, d);
                if (d) {
                    unsigned peers = __match_any_sync(act, oind);
                    if (peers != (1u << lane)) {
                        delta = reduce_peers(peers, delta, lane);
                        if (lane != __ffs(peers) - 1) d = false;
                    }
                }
            }
            if (d) red_add(&A.acc[oind], delta);
        }
        if (ready && alive) {
            if (sc) {
                w.tx -= ds; w.ty -= ds; w.tz -= ds;
                phase = WALK_SCATTER;
                if (!cl && scat > 20) { alive = false; phase = WALK_LEAF; }
            } else {
                w.tx -= tmin; w.ty -= tmin; w.tz -= tmin;
                phase = WALK_CROSS;
            }
            if (nstep > A.max_steps) { alive = false; phase = WALK_LEAF; count_add(&s_cnt[3], A.counters + 3, 1u); }
        }
        // navigation: one table look-up per face crossing, then one descent per iteration while the cell entered is refined
        if (alive && phase == WALK_DESCEND) phase = lw_descend(G, w, ax) ? WALK_LEAF : WALK_DESCEND;
        if (alive && phase == WALK_CROSS) {
            const int r0x = w.cx >> w.level, r0y = w.cy >> w.level, r0z = w.cz >> w.level;
            phase = lw_cross(G, e2, w, ax, A.mirror) ? WALK_LEAF : WALK_DESCEND;
            if (w.cell < 0) alive = false;
            else if (phase == WALK_DESCEND) phase = lw_descend(G, w, ax) ? WALK_LEAF : WALK_DESCEND;    // one level right away
            if (alive && GENERAL && (A.roi.flags & 2)) {                 // WITH_ROI_SAVE: a new root cell? kernel_ASOC.c:615-643
                const int rx = w.cx >> w.level, ry = w.cy >> w.level, rz = w.cz >> w.level;
                if (rx != r0x || ry != r0y || rz != r0z) {
                    const bool r = in_roi_xyz(A.roi, rx, ry, rz);
                    if (r && !in_roi_now) {
                        // entry point in root coordinates: cell origin + fractional position, scaled by the cell size
                        float fx, fy, fz;
                        lw_fraction(w, fx, fy, fz);
                        const float sz = lw_size(w.level);
                        vec3 rp = { ((float)w.cx + fx) * sz, ((float)w.cy + fy) * sz, ((float)w.cz + fz) * sz };
                        roi_save_add(A.roi, rp, lw_dir(w), photons);
                    }
                    in_roi_now = r;
                }
            }
        }
        if (was_alive && !alive) { count_add(&s_cnt[1], A.counters + 1, (unsigned)nstep); count_add(&s_cnt[2], A.counters + 2, (unsigned)min(scat, 20)); }
        }
    }
    __syncthreads();
    if (threadIdx.x < 4 && s_cnt[threadIdx.x]) atomicAdd(A.counters + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}

// =================================================================================================================
// Domain mode, emission pass: one thread per work unit of the chunk; the packet is emitted exactly as the lean / look-ahead
// kernels do it in their refill block and parked in the queue of the domain its first cell belongs to.
// =================================================================================================================
// =================================================================================================================
// Tile pass of the two-pass point-source launch (sim_launch_two_pass, api.cu): emission, the steps of every packet inside the
// SOC_TILE_N^3 cells around the source, then the packet is parked -- complete stepping state, QPk -- at the border of the tile
// for the plain-add look-ahead kernel.  The lean kernel with the tile as its box (DOM = 2) spends 293 warp instructions per
// packet on this (ncu: 16 lanes per instruction, emission with 10; a third of the stall samples wait for instruction fetch):
// here the density of the tile and its accumulator sit in shared memory under tile-local x-fastest indices (no tile test, no
// brick arithmetic), lanes are refilled when half the warp is idle, and lanes are combined (__match_any_sync) only during a
// packet's first steps, where the lanes of a warp share cells.  Arithmetic and order of operations are those of
// sim_lean_kernel, so the paths are the same.
// =================================================================================================================
// KAPPA: per-cell opacities, (kabs*n, ksca*n) per cell as in sim_ahead_kernel
template <bool KAPPA>
__global__ void __launch_bounds__(256, 3) sim_tile_pass_kernel(const __grid_constant__ SimArgs A) {
    extern __shared__ __align__(8) float tp_smem[];
    float *const s_acc = tp_smem;                                             // accumulator of the tile, x fastest
    float *const s_dens = tp_smem + SOC_TILE_CELLS;                           // its density (KAPPA: float2 per cell)
    __shared__ unsigned s_cnt[4];
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0u;
    const GridDesc &G = A.G;
    for (int i = threadIdx.x; i < SOC_TILE_CELLS; i += blockDim.x) {
        const int ux = i % SOC_TILE_N, uy = (i / SOC_TILE_N) % SOC_TILE_N, uz = i / (SOC_TILE_N * SOC_TILE_N);
        const int gi = layout_index(A, A.tile_x0 + ux, A.tile_y0 + uy, A.tile_z0 + uz);
        if (KAPPA) reinterpret_cast<float2 *>(s_dens)[i] = __ldg(A.kappa + gi);
        else s_dens[i] = __ldg(A.dens_brick + gi);
        s_acc[i] = 0.0f;
    }
    __syncthreads();
    float ka = 0.0f;                 // KAPPA: kabs*n of the cell (f.rho holds ksca*n)
    const float kabs = A.kabs, ksca = A.ksca;
    const Box box = launch_box<true>(A);                                      // the tile
    LeanPk<true> f; f.rho = 0.0f; f.sn = 0; f.u = 0; f.upm = 0;
    int ti = 0;                                                               // cell inside the tile
    bool alive = false, wsc = false, more = true;
    bool pend = false;               // parked at the border of the tile: stored at the next refill, the lanes of the warp together
    const int refill = A.refill;
    const unsigned agg = (unsigned)A.agg_steps;
    for (;;) {
        unsigned live = __ballot_sync(FULL, alive);
        if (more ? (32 - __popc(live) >= refill) : (live == 0u)) {
            // one round trip for both requests: queue slots for the packets parked since the last refill and new work units
            const unsigned pm = __ballot_sync(FULL, pend);
            const unsigned nm = ~live;
            const int leader = __ffs(nm) - 1;
            unsigned long long base = 0;
            if (more && lane == leader) base = atomicAdd(A.work, (unsigned long long)__popc(nm));
            if (pm != 0u) {
                QPk o;
                o.tx = f.tx; o.ty = f.ty; o.tz = f.tz; o.rdx = f.rdx; o.rdy = f.rdy; o.rdz = f.rdz;
                o.photons = f.photons; o.free_path = f.free_path; o.tau = f.tau;
                o.ix = (f.upm & 1) ? box.hix - f.cx : box.lox + f.cx; o.iy = (f.upm & 2) ? box.hiy - f.cy : box.loy + f.cy;
                o.iz = (f.upm & 4) ? box.hiz - f.cz : box.loz + f.cz;
                o.upm = (unsigned)f.upm; o.sn = f.sn; o.u = f.u; o.pad = 0u;
                if (!pend) { o.ix = o.iy = o.iz = 0; }
                q_push_warp(A, o, pend);                                      // the queue of the domain that holds the cell (one atomic per domain)
                pend = false;
            }
            if (!more) break;
            base = __shfl_sync(FULL, base, leader);
            if (!alive) {
                const unsigned long long u = base + __popc(nm & ((1u << lane) - 1u));
                if (u < (unsigned long long)A.nlocal) {
                    const unsigned long long us = u + (unsigned long long)A.unit0;
                    const unsigned long long q = us * A.world + A.rank;
                    RngPhilox rng; rng.seed(A.phx, q);
                    Packet pk; pk.ind = -1; pk.level = 0; pk.eidx = -1; pk.rho = 0.0f;
                    emit_ps<SimArgs, RngPhilox, false>(A, rng, (int)(q % (unsigned)A.batch), pk);
                    start_packet(A, rng, pk, true);
                    if (pk.ind >= 0) {
                        alive = true; wsc = false;
                        const int ix = clampi((int)floorf(pk.pos.x), 0, G.nx - 1), iy = clampi((int)floorf(pk.pos.y), 0, G.ny - 1),
                                  iz = clampi((int)floorf(pk.pos.z), 0, G.nz - 1);
                        lean_set_direction<true>(box, f, pk.dir, ix, iy, iz, pk.pos.x - (float)ix, pk.pos.y - (float)iy, pk.pos.z - (float)iz);
                        ti = ((iz - box.loz) * SOC_TILE_N + (iy - box.loy)) * SOC_TILE_N + (ix - box.lox);
                        f.rho = pk.rho; f.photons = pk.photons; f.free_path = pk.free_path; f.tau = 0.0f;
                        if (KAPPA) { const float2 k2 = reinterpret_cast<const float2 *>(s_dens)[ti]; ka = k2.x; f.rho = k2.y; }
                        f.sn = 0; f.u = (unsigned)us;
                    }
                }
            }
            more = base + (unsigned long long)__popc(nm) < (unsigned long long)A.nlocal;
            if (lane == leader) {
                const unsigned long long left = base < (unsigned long long)A.nlocal ? (unsigned long long)A.nlocal - base : 0ull;
                count_add(&s_cnt[0], A.counters + 0, (unsigned)min((unsigned long long)__popc(nm), left));
            }
            live = __ballot_sync(FULL, alive);
            if (live == 0u && !more) break;
        }
        #pragma unroll 1
        for (int rep = 0; rep < 4; rep++) {
            // scatterings, batched: a lane at a scattering point waits until three are (or nothing else can run)
            const unsigned smask = __ballot_sync(FULL, wsc);
            const bool scatter_now = smask != 0u && (__popc(smask) >= 3 || smask == __ballot_sync(FULL, alive));
            if (wsc && scatter_now) {                                         // (as in sim_lean_kernel)
                const bool ux = (f.upm & 1) != 0, uy = (f.upm & 2) != 0, uz = (f.upm & 4) != 0;
                const float adx = rcp_approx(f.rdx), ady = rcp_approx(f.rdy), adz = rcp_approx(f.rdz);
                const float ax = f.tx * adx, ay = f.ty * ady, az = f.tz * adz;
                const float fx = ux ? 1.0f - ax : ax, fy = uy ? 1.0f - ay : ay, fz = uz ? 1.0f - az : az;
                const int ix = ux ? box.hix - f.cx : box.lox + f.cx, iy = uy ? box.hiy - f.cy : box.loy + f.cy, iz = uz ? box.hiz - f.cz : box.loz + f.cz;
                RngBlock rb(A.phx, (unsigned long long)f.u * A.world + A.rank, 0x10000u + LEAN_SCAT(f.sn));
                f.free_path = free_path_fast(A, rb, f.photons);
                const float ct = __ldg(A.csc + clampi((int)(rb.uniform() * A.bins), 0, A.bins - 1));
                vec3 nd = { ux ? adx : -adx, uy ? ady : -ady, uz ? adz : -adz };
                scatter_rotate(nd, ct, SOC_TWOPI * rb.uniform());
                lean_set_direction<true>(box, f, nd, ix, iy, iz, fx, fy, fz);
                f.tau = 0.0f;
                wsc = false;
            }
            // ---- one cell-step ---------------------------------------------------------------------------------------
            float delta = 0.0f, tmin = 0.0f;
            bool px = false, py = false, inb = false, sc = false, parked = false;
            const int oti = ti;
            const bool run = alive && !wsc;
            bool d = run;
            if (run) {
                tmin = fminf(f.tx, fminf(f.ty, f.tz));
                px = f.tx == tmin; py = !px && (f.ty == tmin);
                const int crem = px ? f.cx : (py ? f.cy : f.cz);
                inb = crem > 0;
                const float krho = KAPPA ? f.rho : ksca * f.rho;
                const float tend = fmaf(tmin, krho, f.tau);
                sc = f.free_path < tend;
                const float tsc = (f.free_path - f.tau) * rcp_approx(krho);
                if (sc) { tmin = fminf(tmin, tsc); f.sn += 1u << 24; } else f.tau = tend;
                const float x = KAPPA ? tmin * ka : tmin * f.rho * kabs;
                const float e = exp2f_approx(-1.4426950408889634f * x);
                const float ser = x * fmaf(x, fmaf(x, 0.16666667f, -0.5f), 1.0f);
                const float dfrac = (x < 0.01f) ? ser : (1.0f - e);
                delta = f.photons * dfrac;
                f.photons -= delta;
                f.sn++;
            }
            // ---- deposit ----------------------------------------------------------------------------------------------
            // the first step of every packet lies in the cell of the source: the new lanes of the warp are summed with a
            // butterfly and leave as one add
            {
                const bool first = d && f.sn == 1u;
                const unsigned fm = __ballot_sync(FULL, first);
                if (fm != 0u) {
                    float v = first ? delta : 0.0f;
                    #pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
                    if (lane == __ffs(fm) - 1) atomicAdd(&s_acc[oti], v);
                    if (first) d = false;
                }
            }
            // young packets of a warp still share cells: combined first
            if (__any_sync(FULL, d && LEAN_STEPS(f.sn) <= agg)) {
                const unsigned act = __ballot_sync(FULL, d);
                if (d) {
                    const unsigned peers = __match_any_sync(act, oti);
                    if (peers != (1u << lane)) {
                        delta = reduce_peers(peers, delta, lane);
                        if (lane != __ffs(peers) - 1) d = false;
                    }
                }
            }
            if (d) atomicAdd(&s_acc[oti], delta);
            if (run) {
                f.tx -= tmin; f.ty -= tmin; f.tz -= tmin;
                if (sc) {
                    wsc = true;
                    if (LEAN_SCAT(f.sn) > 20u) { alive = false; wsc = false; }
                } else {
                    const bool pz = !px && !py;
                    const int abit = px ? 1 : (py ? 2 : 4);
                    const int stride = px ? 1 : (py ? SOC_TILE_N : SOC_TILE_N * SOC_TILE_N);
                    f.tx = px ? f.rdx : f.tx; f.ty = py ? f.rdy : f.ty; f.tz = pz ? f.rdz : f.tz;
                    f.cx -= px; f.cy -= py; f.cz -= pz;
                    ti += (f.upm & abit) ? stride : -stride;
                    alive = inb;
                    if (inb) {
                        if (KAPPA) { const float2 k2 = reinterpret_cast<const float2 *>(s_dens)[ti]; ka = k2.x; f.rho = k2.y; }
                        else f.rho = s_dens[ti];
                    }
                    else {
                        // left the tile: through a face of the grid the packet is gone, else it is parked for the second pass
                        // (the counter of the crossed axis stands at -1 = one cell beyond the border)
                        const int face = (px ? 0 : (py ? 2 : 4)) + ((f.upm & abit) ? 1 : 0);
                        if (!((A.dom_faces >> face) & 1)) { parked = true; pend = true; }
                    }
                }
                bool stuck = false;
                if (LEAN_STEPS(f.sn) > (unsigned)A.max_steps && !parked) { alive = false; wsc = false; stuck = true; }
                if (!alive && !parked) {                            // packet finished: once per packet
                    count_add(&s_cnt[1], A.counters + 1, LEAN_STEPS(f.sn));
                    count_add(&s_cnt[2], A.counters + 2, min(LEAN_SCAT(f.sn), 20u));
                    if (stuck) count_add(&s_cnt[3], A.counters + 3, 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SOC_TILE_CELLS; i += blockDim.x) {          // the tile's accumulator -> ACC (brick order)
        const float v = s_acc[i];
        if (v != 0.0f) {
            const int ux = i % SOC_TILE_N, uy = (i / SOC_TILE_N) % SOC_TILE_N, uz = i / (SOC_TILE_N * SOC_TILE_N);
            red_add(&A.acc[layout_index(A, A.tile_x0 + ux, A.tile_y0 + uy, A.tile_z0 + uz)], v);
        }
    }
    if (threadIdx.x < 4 && s_cnt[threadIdx.x]) atomicAdd(A.counters + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}

template <bool BRICK>
__global__ void __launch_bounds__(256) sim_emit_queue_kernel(const __grid_constant__ SimArgs A, long long n) {
    const GridDesc &G = A.G;
    const Box box = launch_box<false>(A);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nround = (n + 31) & ~31LL;                 // whole warps stay in the loop (warp-wide queue push)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
        const unsigned long long u = (unsigned long long)(A.unit0 + i);
        const unsigned long long q = u * A.world + A.rank;
        Packet pk; pk.ind = -1; pk.level = 0; pk.eidx = -1; pk.rho = 0.0f;
        QPk o;
        o.ix = o.iy = o.iz = 0;
        if (i < n) {
        RngPhilox rng; rng.seed(A.phx, q);
        const int id = (int)(q / (unsigned)A.batch), III = (int)(q % (unsigned)A.batch);
        if (A.kind == SIM_PS)      emit_ps<SimArgs, RngPhilox, false>(A, rng, III, pk);
        else if (A.kind == SIM_BG) emit_bg<SimArgs, RngPhilox, false>(A, rng, id, pk);
        else                       emit_hp<SimArgs, RngPhilox, false>(A, rng, pk);
        start_packet(A, rng, pk, A.kind != SIM_HP);
        }
        if (pk.ind >= 0) {
            const int ix = clampi((int)floorf(pk.pos.x), 0, G.nx - 1), iy = clampi((int)floorf(pk.pos.y), 0, G.ny - 1),
                      iz = clampi((int)floorf(pk.pos.z), 0, G.nz - 1);
            LeanPk<BRICK> f;
            lean_set_direction<BRICK>(box, f, pk.dir, ix, iy, iz, pk.pos.x - (float)ix, pk.pos.y - (float)iy, pk.pos.z - (float)iz);
            o.tx = f.tx; o.ty = f.ty; o.tz = f.tz; o.rdx = f.rdx; o.rdy = f.rdy; o.rdz = f.rdz;
            o.photons = pk.photons; o.free_path = pk.free_path; o.tau = 0.0f;
            o.ix = ix; o.iy = iy; o.iz = iz;
            o.upm = (unsigned)f.upm; o.sn = 0u; o.u = (unsigned)u; o.pad = 0u;
        }
        q_push_warp(A, o, pk.ind >= 0);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(A.counters + 0, (unsigned long long)n);
}

// with_abu alone stays on the lean path (look-ahead kernel with the per-cell opacity array) unless a border reflects
static bool sim_is_general(const SimArgs &A) {
    return (A.roi.flags & 2) || A.kind == SIM_ROI || A.with_msf || (A.with_abu && (A.mirror != 0 || A.kappa == nullptr)) || A.save_int2 ||
           A.with_ali || A.kind == SIM_CL;
}

static void launch_walk(const SimArgs &A, int blocks, int threads, cudaStream_t stream) {
    const bool general = sim_is_general(A);
    // combining lanes pays only when packets share cells: point-source packets (measured on the bench octree: the
    // background launch runs 4.6e10 cell-steps/s with plain adds, 4.0e10 with the match/reduce path)
    const bool red = (A.save_int2 || A.with_ali) || A.deposit == DEP_RED || A.kind != SIM_PS;
    if (A.nbr != nullptr) {              // neighbour table: no climbs (linkwalk.cuh)
        if (general) {
            if (red) sim_link_kernel<DEP_RED, true><<<blocks, threads, 0, stream>>>(A);
            else     sim_link_kernel<DEP_WARP, true><<<blocks, threads, 0, stream>>>(A);
        } else {
            if (red) sim_link_kernel<DEP_RED, false><<<blocks, threads, 0, stream>>>(A);
            else     sim_link_kernel<DEP_WARP, false><<<blocks, threads, 0, stream>>>(A);
        }
    } else if (general) {
        if (red) sim_walk_kernel<DEP_RED, true><<<blocks, threads, 0, stream>>>(A);
        else     sim_walk_kernel<DEP_WARP, true><<<blocks, threads, 0, stream>>>(A);
    } else {
        if (red) sim_walk_kernel<DEP_RED, false><<<blocks, threads, 0, stream>>>(A);
        else     sim_walk_kernel<DEP_WARP, false><<<blocks, threads, 0, stream>>>(A);
    }
}

// TABS += acc * TW*ADHOC ; INT += acc ; acc = 0   (float4 streams; acc is all-zero again afterwards)
__global__ void __launch_bounds__(256) fold_acc_kernel(float *__restrict__ acc, float *__restrict__ tabs, float *__restrict__ inten,
                                                       float scale, long long n) {
    const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 a = reinterpret_cast<float4 *>(acc)[i];
        if (a.x != 0.0f || a.y != 0.0f || a.z != 0.0f || a.w != 0.0f) {
            float4 t = reinterpret_cast<float4 *>(tabs)[i];
            t.x += a.x * scale; t.y += a.y * scale; t.z += a.z * scale; t.w += a.w * scale;
            reinterpret_cast<float4 *>(tabs)[i] = t;
            if (inten != nullptr) {
                float4 v = reinterpret_cast<float4 *>(inten)[i];
                v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
                reinterpret_cast<float4 *>(inten)[i] = v;
            }
            reinterpret_cast<float4 *>(acc)[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float a = acc[i];
        if (a != 0.0f) { tabs[i] += a * scale; if (inten != nullptr) inten[i] += a; acc[i] = 0.0f; }
    }
}

// MWC64X work-item stream with the item-kernel interface
struct RngMwcItem : RngMwc {
    __device__ __forceinline__ void seed_item(const SimArgs &A, unsigned long long id) { seed(A.mwc, id); }
};

}  // namespace

// name of the packet kernel the last launch_sim / launch_sim_domain dispatched (soc_last_kernel)
static thread_local char g_kernel_name[128] = "";
static void note_kernel(const char *base, int dep, int brick, int extra, const char *extra_name, int dom) {
    static const char *deps[3] = { "DEP_RED", "DEP_WARP", "DEP_TILE" };
    snprintf(g_kernel_name, sizeof(g_kernel_name), "%s<%s,%s,%s=%d%s>", base, deps[dep < 0 || dep > 2 ? 0 : dep], brick ? "brick" : "linear",
             extra_name, extra, dom ? ",domains" : "");
}
const char *sim_last_kernel() { return g_kernel_name; }

template <int DEP, bool BRICK, int CTAS, bool KAPPA, bool DOM>
static void launch_ahead_dep(const SimArgs &A, int blocks, int threads, cudaStream_t stream) {
    static int per_sm = 0, sms = 0;                   // resident CTAs of this instantiation: the persistent grid is sms x per_sm
    if (per_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sim_ahead_kernel<DEP, BRICK, CTAS, KAPPA, DOM>, threads, 0);
        if (per_sm < 1) per_sm = 1;
    }
    if (blocks > sms * per_sm) blocks = sms * per_sm;
    sim_ahead_kernel<DEP, BRICK, CTAS, KAPPA, DOM><<<blocks, threads, 0, stream>>>(A);
    note_kernel("sim_ahead_kernel", DEP, BRICK, KAPPA, "kappa", DOM);
}

template <bool BRICK, int CTAS, bool KAPPA, bool DOM = false>
static void launch_ahead(const SimArgs &A, int dep, int blocks, int threads, cudaStream_t stream) {
    if (dep == DEP_RED)       launch_ahead_dep<DEP_RED, BRICK, CTAS, KAPPA, DOM>(A, blocks, threads, stream);
    else if (dep == DEP_WARP) launch_ahead_dep<DEP_WARP, BRICK, CTAS, KAPPA, DOM>(A, blocks, threads, stream);
    else                      launch_ahead_dep<DEP_TILE, BRICK, CTAS, KAPPA, DOM>(A, blocks, threads, stream);
}

template <bool BRICK, bool PEND>
static void launch_lean(SimArgs A, int dep, int blocks, int threads, cudaStream_t stream) {
    if (BRICK && dep == DEP_TILE) {          // z-slab of the shared-memory tile in bricked order
        const int slab = 2 * A.G.nx * A.G.ny, z0 = A.tile_z0;
        A.tile_lo = (z0 >> 1) * slab;
        A.tile_span = (((z0 + SOC_TILE_N - 1) >> 1) - (z0 >> 1) + 1) * slab;
    }
    // per-cell opacities: only the look-ahead kernel has the variant (one float2 per cell, see sim_ahead_kernel)
    if (A.with_abu) {
        launch_ahead<BRICK, 3, true>(A, dep, blocks, threads, stream);
        return;
    }
    // geometry one cell ahead of the physics (cp.async density ring).  Measured on the bench step: background launch
    // (plain adds) 60.4 -> 57.9 ms; the point-source launch with the shared-memory tile runs 63.9 ms on the lean kernel,
    // 73.6 ms (3 CTAs / SM) or 71.8 ms (4 CTAs / SM) here, so ahead = 1 takes the plain-add launches only
    // Grids whose DENS + ACC exceed the L2 (512^3: 1 GiB) are bound by the rate of random DRAM transactions instead: there the look-ahead wins for
    // every accumulation engine (512^3 point-source launch 368 -> 342 ms, background 310 -> 275 ms)
    const bool beyond_l2 = A.G.nxyz > (1LL << 25);
    if (A.ahead && !PEND && A.mirror == 0 && (dep == DEP_RED || A.ahead > 1 || beyond_l2)) {
        if (A.ahead == 2) launch_ahead<BRICK, 4, false>(A, dep, blocks, threads, stream);
        else              launch_ahead<BRICK, 3, false>(A, dep, blocks, threads, stream);
        return;
    }
    if (dep == DEP_RED)       sim_lean_kernel<DEP_RED, BRICK, PEND, false><<<blocks, threads, 0, stream>>>(A);
    else if (dep == DEP_WARP) sim_lean_kernel<DEP_WARP, BRICK, PEND, false><<<blocks, threads, 0, stream>>>(A);
    else                      sim_lean_kernel<DEP_TILE, BRICK, PEND, false><<<blocks, threads, 0, stream>>>(A);
    note_kernel("sim_lean_kernel", dep, BRICK, PEND, "pend", 0);
}


// the lean kernel keeps the work unit in 32 bits and the step count of a packet in 24
static bool sim_uses_lean(const SimArgs &A) { return !sim_is_general(A) && A.nlocal < (1LL << 32) && A.max_steps < (1 << 24) - 2; }

static void launch_fast(const SimArgs &A, int blocks, int threads, cudaStream_t stream) {
    const int dep = (A.save_int2 || A.with_ali) ? DEP_RED : A.deposit;
    if (!sim_uses_lean(A)) {
        note_kernel("sim_fast_kernel", dep, 0, 1, "general", 0);
        if (dep == DEP_RED)       sim_fast_kernel<DEP_RED, true><<<blocks, threads, 0, stream>>>(A);
        else if (dep == DEP_WARP) sim_fast_kernel<DEP_WARP, true><<<blocks, threads, 0, stream>>>(A);
        else                      sim_fast_kernel<DEP_TILE, true><<<blocks, threads, 0, stream>>>(A);
    } else if (A.brick) {
        if (A.pend) launch_lean<true, true>(A, dep, blocks, threads, stream);
        else        launch_lean<true, false>(A, dep, blocks, threads, stream);
    } else {
        if (A.pend) launch_lean<false, true>(A, dep, blocks, threads, stream);
        else        launch_lean<false, false>(A, dep, blocks, threads, stream);
    }
}

// ---- domain mode ------------------------------------------------------------------------------------------------------
// Eligible: the launches the bricked lean / look-ahead kernels serve, without reflecting borders (a domain face that is a
// face of the grid ends the packet).  Whether it pays (DENS + accumulator beyond the L2) is decided by the caller.
bool sim_domains_eligible(const SimArgs &A, int rng_mode) {
    return sim_uses_bricks(A, rng_mode) && A.brick && A.mirror == 0 && !A.pend;
}

// ---- counting sort of a domain queue by (16^3-cell block of the entry cell, direction octant): neighbouring lanes then
// work on packets that cross the same cells, which is what keeps the L1 / L2 sector traffic per step down (the order in
// which packets are parked is random).  32768 keys; histogram, scan and scatter are three small kernels.
#define Q_SORT_KEYS 32768
__device__ __forceinline__ unsigned q_sort_key(const QPk &v, int lox, int loy, int loz) {
    const unsigned kx = (unsigned)(v.ix - lox) >> 4, ky = (unsigned)(v.iy - loy) >> 4, kz = (unsigned)(v.iz - loz) >> 4;
    return ((((kz & 15u) * 16u + (ky & 15u)) * 16u + (kx & 15u)) << 3) | (v.upm & 7u);
}
__global__ void __launch_bounds__(256) q_hist_kernel(const QPk *__restrict__ q, long long n, unsigned *__restrict__ hist, int lox, int loy, int loz) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        QPk v;
        reinterpret_cast<float4 *>(&v)[2] = reinterpret_cast<const float4 *>(q + i)[2];      // tau, ix, iy, iz
        reinterpret_cast<float4 *>(&v)[3] = reinterpret_cast<const float4 *>(q + i)[3];      // upm, ...
        atomicAdd(hist + q_sort_key(v, lox, loy, loz), 1u);
    }
}
// exclusive scan of hist[Q_SORT_KEYS] in place (one block of 1024 threads, 32 keys each)
__global__ void __launch_bounds__(1024) q_scan_kernel(unsigned *__restrict__ hist) {
    __shared__ unsigned part[1024];
    const int t = threadIdx.x;
    unsigned loc[32], sum = 0;
    #pragma unroll
    for (int k = 0; k < 32; k++) { loc[k] = hist[t * 32 + k]; sum += loc[k]; }
    part[t] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        unsigned v = (t >= off) ? part[t - off] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned run = part[t] - sum;
    #pragma unroll
    for (int k = 0; k < 32; k++) { hist[t * 32 + k] = run; run += loc[k]; }
}
__global__ void __launch_bounds__(256) q_scatter_kernel(const QPk *__restrict__ q, long long n, unsigned *__restrict__ cursor, QPk *__restrict__ out,
                                                        int lox, int loy, int loz) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const QPk v = q_load(q + i);
        const unsigned slot = atomicAdd(cursor + q_sort_key(v, lox, loy, loz), 1u);
        q_store(out + slot, v);
    }
}
void launch_queue_sort(const QPk *q, long long n, QPk *out, unsigned *hist, const int lo[3], cudaStream_t stream) {
    long long b = (n + 255) / 256;
    const int blocks = (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
    cudaMemsetAsync(hist, 0, Q_SORT_KEYS * sizeof(unsigned), stream);
    q_hist_kernel<<<blocks, 256, 0, stream>>>(q, n, hist, lo[0], lo[1], lo[2]);
    q_scan_kernel<<<1, 1024, 0, stream>>>(hist);
    q_scatter_kernel<<<blocks, 256, 0, stream>>>(q, n, hist, out, lo[0], lo[1], lo[2]);
}

// the last parked packets, resumed on the whole grid by the general kernel (q_in, A.dom set by the caller)
void launch_sim_cleanup(const SimArgs &A, int blocks, int threads, cudaStream_t stream) {
    // all queues in one launch on the bricked layout: the look-ahead kernel with re-homing (ROAM); one queue at a time (more than 64
    // boxes) or SOC_AHEAD=0: the general kernel on the reference's cell order, which adds straight into TABS / INT
    if (A.ahead && A.brick && A.q_nparts > 1) {
        static int per_sm[2] = { 0, 0 }, sms = 0;
        const int v = A.with_abu ? 1 : 0;
        if (per_sm[v] == 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (v) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[v], sim_ahead_kernel<DEP_RED, true, 3, true, true, true>, threads, 0);
            else   cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[v], sim_ahead_kernel<DEP_RED, true, 3, false, true, true>, threads, 0);
            if (per_sm[v] < 1) per_sm[v] = 1;
        }
        if (blocks > sms * per_sm[v]) blocks = sms * per_sm[v];
        if (v) sim_ahead_kernel<DEP_RED, true, 3, true, true, true><<<blocks, threads, 0, stream>>>(A);
        else   sim_ahead_kernel<DEP_RED, true, 3, false, true, true><<<blocks, threads, 0, stream>>>(A);
        return;
    }
    sim_fast_kernel<DEP_RED, true><<<blocks, threads, 0, stream>>>(A);
}

void launch_sim_emit(const SimArgs &A, long long nunits, cudaStream_t stream) {
    long long b = (nunits + 255) / 256;
    const long long cap = 148LL * 16;
    sim_emit_queue_kernel<true><<<(int)(b < 1 ? 1 : (b > cap ? cap : b)), 256, 0, stream>>>(A, nunits);
}

// Same choice of kernel as launch_lean(): per-cell opacities and plain adds on the look-ahead kernel, the shared-memory
// tile / lane combining of point-source launches on the lean kernel (a domain is sized to live in the L2).
void launch_sim_domain(const SimArgs &A, int blocks, int threads, cudaStream_t stream) {
    const int dep = A.deposit;
    // without the emission code the look-ahead kernel needs <= 64 registers: 4 CTAs per SM
    static int ctas = 0;
    if (ctas == 0) { const char *e = getenv("SOC_DOM_CTAS"); ctas = (e && atoi(e) == 3) ? 3 : 4; }                    // tuning knob
    if (A.with_abu) {
        if (ctas == 4) launch_ahead<true, 4, true, true>(A, dep, blocks, threads, stream);
        else           launch_ahead<true, 3, true, true>(A, dep, blocks, threads, stream);
    } else if (A.ahead && (dep == DEP_RED || A.ahead > 1)) {
        if (ctas == 4) launch_ahead<true, 4, false, true>(A, dep, blocks, threads, stream);
        else           launch_ahead<true, 3, false, true>(A, dep, blocks, threads, stream);
    }
    else {
        if (dep == DEP_RED)       sim_lean_kernel<DEP_RED, true, false, true><<<blocks, threads, 0, stream>>>(A);
        else if (dep == DEP_WARP) sim_lean_kernel<DEP_WARP, true, false, true><<<blocks, threads, 0, stream>>>(A);
        else                      sim_lean_kernel<DEP_TILE, true, false, true><<<blocks, threads, 0, stream>>>(A);
        note_kernel("sim_lean_kernel", dep, 1, 0, "pend", 1);
    }
}

// ---- two-pass point-source launch (grids that live in the L2) --------------------------------------------------------------
// The one-pass launch runs the lean kernel with the shared-memory tile: the tile test, the lane combining and the emission
// (8 of 32 lanes at a refill) are carried by every iteration of a launch that is bound by instruction issue.  Two passes:
// (1) the lean kernel with the tile as its box -- emission, the first ~8 steps of every packet into shared memory, then the
// packet is parked (complete stepping state) at the border of the tile; (2) the plain-add look-ahead kernel over the whole
// grid, fed from the queue.  Same packets, same paths.
// can the emission and the steps inside the shared-memory tile run as a pass of their own (sim_tile_pass_kernel)?
bool sim_tile_pass_eligible(const SimArgs &A, int rng_mode) {
    return rng_mode != SOC_RNG_REFERENCE && A.kind == SIM_PS && A.deposit == DEP_TILE && A.tile_inside && A.brick && A.ahead && !A.pend &&
           A.mirror == 0 && sim_uses_lean(A) && A.G.nx >= SOC_TILE_N && A.G.ny >= SOC_TILE_N && A.G.nz >= SOC_TILE_N;
}
bool sim_two_pass_eligible(const SimArgs &A, int rng_mode) {
    // larger grids run domain by domain (there the tile pass replaces the emission pass) or, with that switched off, the look-ahead tile kernel
    return sim_tile_pass_eligible(A, rng_mode) && A.G.nxyz <= (1LL << 25);
}
template <bool KAPPA>
static void launch_tile_pass(const SimArgs &A, cudaStream_t stream) {
    static int per_sm = 0, sms = 0;                   // resident CTAs: the grid is sms x per_sm
    const size_t smem = (size_t)SOC_TILE_CELLS * (KAPPA ? 3 : 2) * sizeof(float);
    if (per_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(sim_tile_pass_kernel<KAPPA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sim_tile_pass_kernel<KAPPA>, 256, smem);
        if (per_sm < 1) per_sm = 1;
    }
    long long need = (A.nlocal + 255) / 256;
    const int b = (int)(need < (long long)sms * per_sm ? (need < 1 ? 1 : need) : (long long)sms * per_sm);
    sim_tile_pass_kernel<KAPPA><<<b, 256, smem, stream>>>(A);
}
void launch_sim_tile_pass(const SimArgs &A, int blocks, int threads, cudaStream_t stream) {
    (void)blocks; (void)threads;
    if (A.with_abu) launch_tile_pass<true>(A, stream);
    else            launch_tile_pass<false>(A, stream);
}
void sim_note_two_pass() {
    const char *second = sim_last_kernel();          // what launch_sim_domain dispatched for the second pass
    char tmp[128];
    snprintf(tmp, sizeof(tmp), "sim_tile_pass_kernel + %.96s", second[0] ? second : "sim_ahead_kernel");
    snprintf(g_kernel_name, sizeof(g_kernel_name), "%s", tmp);
}

bool sim_kappa_eligible(const SimArgs &A, int rng_mode) {
    return rng_mode != SOC_RNG_REFERENCE && A.G.levels == 1 && !A.ref_geometry && A.with_abu && !A.with_msf && A.mirror == 0 &&
           !(A.roi.flags & 2) && (A.kind == SIM_PS || A.kind == SIM_BG || A.kind == SIM_HP) && !A.save_int2 && !A.with_ali &&
           A.nlocal < (1LL << 32) && A.max_steps < (1 << 24) - 2;
}

// position i of the (domain-major) brick order -> index of the cell in the reference's x-fastest order
struct LayoutDesc { int nx, ny, ds0, ds1, ds2, ns0, ns1, one32; long long dcells; };
__device__ __forceinline__ long long layout_source(const LayoutDesc &L, long long i) {
    if (L.one32) {                           // one box (grids up to 2^31 cells): 32-bit arithmetic -- the 64-bit divisions below made the
        const unsigned r = (unsigned)i;      // fold kernel compute-bound (0.15 ms for 256^3 where the streams need 0.06 ms)
        const unsigned b = r >> 3, sub = r & 7u;
        const unsigned hx = (unsigned)L.ds0 >> 1, hy = (unsigned)L.ds1 >> 1;
        const unsigned t = b / hx, bx = b - t * hx, bz = t / hy, by = t - bz * hy;
        const unsigned ix = 2u * bx + (sub & 1u), iy = 2u * by + ((sub >> 1) & 1u), iz = 2u * bz + (sub >> 2);
        return (long long)((iz * (unsigned)L.ny + iy) * (unsigned)L.nx + ix);
    }
    const long long d = i / L.dcells, r = i - d * L.dcells;
    const long long b = r >> 3;
    const int sub = (int)(r & 7);
    const int hx = L.ds0 >> 1, hy = L.ds1 >> 1;
    const int bx = (int)(b % hx), by = (int)((b / hx) % hy), bz = (int)(b / ((long long)hx * hy));
    const int dx = (int)(d % L.ns0), dy = (int)((d / L.ns0) % L.ns1), dz = (int)(d / ((long long)L.ns0 * L.ns1));
    const int ix = dx * L.ds0 + 2 * bx + (sub & 1), iy = dy * L.ds1 + 2 * by + ((sub >> 1) & 1), iz = dz * L.ds2 + 2 * bz + (sub >> 2);
    return ((long long)iz * L.ny + iy) * L.nx + ix;
}
static LayoutDesc layout_of(const SimArgs &A) {
    LayoutDesc L;
    L.nx = A.G.nx; L.ny = A.G.ny;
    L.ds0 = A.dsize[0]; L.ds1 = A.dsize[1]; L.ds2 = A.dsize[2]; L.ns0 = A.dsplit[0]; L.ns1 = A.dsplit[1];
    L.dcells = (long long)L.ds0 * L.ds1 * L.ds2;
    L.one32 = (A.dsplit[0] * A.dsplit[1] * A.dsplit[2] == 1 && L.dcells < (1LL << 31)) ? 1 : 0;     // one box, indices fit 32 bits
    return L;
}

// (kabs*n, ksca*n) per cell, in brick order when the launch uses bricks: one thread per output cell
__global__ void __launch_bounds__(256) kappa_kernel(const float *__restrict__ dens, const float2 *__restrict__ opt, float2 *__restrict__ out,
                                                    const LayoutDesc L, long long n, int brick) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long src = brick ? layout_source(L, i) : i;
        const float d = dens[src];
        const float2 o = opt[src];
        out[i] = make_float2(o.x * d, o.y * d);
    }
}

bool sim_uses_bricks(const SimArgs &A, int rng_mode) {
    return sim_uses_lean(A) && rng_mode != SOC_RNG_REFERENCE && A.G.levels == 1 && !A.ref_geometry && !sim_is_general(A) && A.dens_brick != nullptr;
}

void launch_sim(const SimArgs &A, int rng_mode, int blocks, int threads, cudaStream_t stream) {
    const bool oct = A.G.levels > 1, dbl = A.G.dbl_sim != 0;
    if (rng_mode == SOC_RNG_REFERENCE) snprintf(g_kernel_name, sizeof(g_kernel_name), "sim_item_kernel<MWC64X,%s,%s>", oct ? "octree" : "regular", dbl ? "f64" : "f32");
    else if (oct && !A.ref_geometry) snprintf(g_kernel_name, sizeof(g_kernel_name), "%s<%s>", A.nbr != nullptr ? "sim_link_kernel" : "sim_walk_kernel", sim_is_general(A) ? "general" : "lean");
    else if (A.ref_geometry) snprintf(g_kernel_name, sizeof(g_kernel_name), "sim_stream_kernel<%s,%s>", oct ? "octree" : "regular", dbl ? "f64" : "f32");
    if (rng_mode == SOC_RNG_REFERENCE) {
        if (!oct)      sim_item_kernel<RngMwcItem, false, false><<<blocks, threads, 0, stream>>>(A);
        else if (!dbl) sim_item_kernel<RngMwcItem, true, false><<<blocks, threads, 0, stream>>>(A);
        else           sim_item_kernel<RngMwcItem, true, true><<<blocks, threads, 0, stream>>>(A);
    } else {
        if (!oct && !A.ref_geometry) launch_fast(A, blocks, threads, stream);
        else if (!oct) sim_stream_kernel<false, false><<<blocks, threads, 0, stream>>>(A);
        else if (!A.ref_geometry) launch_walk(A, blocks, threads, stream);
        else if (!dbl) sim_stream_kernel<true, false><<<blocks, threads, 0, stream>>>(A);
        else           sim_stream_kernel<true, true><<<blocks, threads, 0, stream>>>(A);
    }
}

int sim_blocks_per_sm(int rng_mode, bool octree, bool dbl, int threads) {
    int n = 0;
    if (rng_mode == SOC_RNG_REFERENCE) {
        if (!octree)   cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sim_item_kernel<RngMwcItem, false, false>, threads, 0);
        else if (!dbl) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sim_item_kernel<RngMwcItem, true, false>, threads, 0);
        else           cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sim_item_kernel<RngMwcItem, true, true>, threads, 0);
    } else {
        if (!octree)   cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sim_fast_kernel<DEP_TILE, true>, threads, 0);   // == the lean kernels (launch bounds 256 x 4)
        else           cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sim_walk_kernel<DEP_WARP, true>, threads, 0);
    }
    return n > 0 ? n : 1;
}

// bricked scratch accumulator -> TABS / INT in the reference's cell order.  One thread per x-pair of a brick: the
// accumulator is read as float2 in its own order (coalesced), TABS / INT are updated 8 bytes at a time.
__global__ void __launch_bounds__(256) fold_acc_brick_kernel(float *__restrict__ acc, float *__restrict__ tabs, float *__restrict__ inten,
                                                             float scale, const LayoutDesc L, long long nquads) {
    // one thread = half a brick: cells (x, y), (x+1, y), (x, y+1), (x+1, y+1) of one z -- 16 contiguous bytes of the accumulator,
    // two 8-byte pieces of TABS / INT one row apart (the loads of a thread are independent: enough bytes in flight to stream)
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nquads; i += stride) {
        const float4 a = reinterpret_cast<float4 *>(acc)[i];
        if (a.x != 0.0f || a.y != 0.0f || a.z != 0.0f || a.w != 0.0f) {
            const long long lin0 = layout_source(L, 4 * i), lin1 = lin0 + L.nx;
            float2 t0 = *reinterpret_cast<float2 *>(tabs + lin0), t1 = *reinterpret_cast<float2 *>(tabs + lin1);
            float2 v0 = make_float2(0.0f, 0.0f), v1 = v0;
            if (inten != nullptr) { v0 = *reinterpret_cast<float2 *>(inten + lin0); v1 = *reinterpret_cast<float2 *>(inten + lin1); }
            t0.x += a.x * scale; t0.y += a.y * scale; t1.x += a.z * scale; t1.y += a.w * scale;
            *reinterpret_cast<float2 *>(tabs + lin0) = t0; *reinterpret_cast<float2 *>(tabs + lin1) = t1;
            if (inten != nullptr) {
                v0.x += a.x; v0.y += a.y; v1.x += a.z; v1.y += a.w;
                *reinterpret_cast<float2 *>(inten + lin0) = v0; *reinterpret_cast<float2 *>(inten + lin1) = v1;
            }
            reinterpret_cast<float4 *>(acc)[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
    }
}

__global__ void __launch_bounds__(256) brick_permute_kernel(const float *__restrict__ dens, float *__restrict__ out, const LayoutDesc L, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = dens[layout_source(L, i)];
}

static int stream_grid(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = 148LL * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

void launch_brick_permute(const SimArgs &A, float *dens_brick, cudaStream_t stream) {
    const long long n = A.G.nxyz;
    brick_permute_kernel<<<stream_grid(n), 256, 0, stream>>>(A.G.dens, dens_brick, layout_of(A), n);
}

void launch_kappa(const SimArgs &A, cudaStream_t stream) {
    const long long n = A.G.nxyz;
    kappa_kernel<<<stream_grid(n), 256, 0, stream>>>(A.G.dens, reinterpret_cast<const float2 *>(A.opt), const_cast<float2 *>(A.kappa),
                                                     layout_of(A), n, A.brick);
}

void launch_fold_acc(const SimArgs &A, cudaStream_t stream) {
    const long long n = A.G.cells;
    const float scale = A.tw * A.adhoc;
    if (A.brick) fold_acc_brick_kernel<<<stream_grid(n >> 2), 256, 0, stream>>>(A.acc, A.tabs, A.use_int ? A.inten : nullptr, scale, layout_of(A), n >> 2);
    else         fold_acc_kernel<<<stream_grid(n >> 2), 256, 0, stream>>>(A.acc, A.tabs, A.use_int ? A.inten : nullptr, scale, n);
}
