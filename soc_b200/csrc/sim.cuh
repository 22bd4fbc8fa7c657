// soc_b200 -- arguments of the Monte Carlo packet kernels (emission / absorption run).
#pragma once
#include "common.cuh"

enum SimKind { SIM_PS = 0, SIM_BG = 1, SIM_HP = 2, SIM_CL = 3, SIM_ROI = 4 };

// accumulation engine of the absorption counters (north-star item b)
enum DepositMode {
    DEP_RED = 0,        // one red.global.add.f32 per lane and step
    DEP_WARP = 1,       // lanes of a warp that hit the same cell are combined first (__match_any_sync)
    DEP_TILE = 2        // DEP_WARP + a shared-memory tile of cells around the (first) point source
};

#define SOC_TILE_N 16                    // shared-memory tile edge (cells); SOC_TILE_N^3 floats = 16 KiB
#define SOC_TILE_CELLS (SOC_TILE_N * SOC_TILE_N * SOC_TILE_N)

// Domain-tiled propagation (regular grids whose DENS + scratch accumulator exceed the L2): the packet kernels work on
// one box of cells ("domain") at a time so that its arrays stay in the L2; a packet that leaves the domain through an
// interior face is parked in the queue of the domain it enters -- its complete DDA state, so that the path continues
// bit for bit -- and picked up when that domain is processed.
struct __align__(16) QPk {
    float tx, ty, tz, rdx;          // face distances, 1/|d|
    float rdy, rdz, photons, free_path;
    float tau; int ix, iy, iz;      // root-grid coordinates of the cell the packet is in
    unsigned upm, sn, u, pad;       // direction signs, steps | scatterings << 24, work unit (Philox stream)
};

struct SimArgs {
    GridDesc G;
    // domain mode
    int dom;                            // 1 = this launch works on one domain and takes its packets from q_in
    int dom_lo[3], dom_hi[3];           // cells of the domain (inclusive)
    int dom_faces;                      // bit 2*axis + (towards +axis ? 1 : 0): that face of the domain is a face of the grid
    int dsplit[3], dsize[3];            // layout of the bricked arrays: domains per axis, cells per domain along each axis (one domain = plain brick order)
    int dom_base;                       // first position of this launch's domain in the bricked arrays
    const QPk *__restrict__ q_in;       // packets of this launch (nlocal of them)
    QPk *q_base; unsigned *q_tail;      // queues of all domains: q_base[d * q_cap + i], i < q_tail[d]
    long long q_cap;
    int q_nparts; long long q_part[65];  // clean-up pass: q_in is q_nparts queues back to back in unit space, queue d holds units q_part[d] .. q_part[d+1]
    long long unit0;                    // emission pass: first work unit of the chunk
    unsigned long long scramble;        // experiment: work unit u runs packet (u * scramble) % nlocal (0 = in order)
    // accumulators [cells]
    float *tabs, *xab, *inten, *intx, *inty, *intz;
    // per-launch scratch accumulator [cells]: the stream kernels add the unweighted absorbed energy here and
    // fold_acc() spreads it to TABS (x TW*ADHOC) and INT afterwards -- one atomic per cell-step instead of two
    float *acc;
    int use_acc;
    // regular grids with even dimensions: a second copy of DENS in 2x2x2 bricks (one 32-byte sector per brick), so
    // that the cell entered next shares the sector of the current one half of the time; `acc` then uses the same
    // order and fold_acc() un-bricks it (brick = 1)
    const float *__restrict__ dens_brick;
    int brick;
    const float2 *__restrict__ kappa;   // WITH_ABU on the lean path: (kabs*n, ksca*n) per cell in the layout of the density array
    const int *__restrict__ nbr; // octrees: neighbour table [6*cells] of linkwalk.cuh (nullptr: climb through PAR, walk.cuh)
    int pend;                    // lean kernel: merge deposits into aligned 4-cell groups (red.global.add.v4.f32)
    int ahead;                   // lean kernel: geometry one cell ahead of the physics (sim_ahead_kernel)
    int slab_xy, brick_by, brick_bz;   // nx*ny; index increments to the next brick along y and z
    // inputs
    const float *__restrict__ emit, *__restrict__ emwei, *__restrict__ opt;
    const float *__restrict__ dsc, *__restrict__ csc;
    const float *__restrict__ pspos, *__restrict__ ps, *__restrict__ xps_area;
    const int *__restrict__ xps_nside, *__restrict__ xps_side;
    const float *__restrict__ hpbg, *__restrict__ hpbgp;
    const float *__restrict__ abu, *__restrict__ scav;   // WITH_MSF: ABU[cells*ndust], SCA[ndust]
    int with_msf, ndust, mirror;
    RoiDesc roi;                 // region of interest (flags 1 load, 2 save)
    int roi_nelem;               // SIM_ROI: number of surface elements of the loaded file (the PACKETS argument)
    float kabs, ksca, bg, tw, adhoc, sw_a, sw_b;
    int kind, batch, global;
    int bins, no_ps, ps_method, with_abu, with_ali, use_int, save_int2, use_emweight, hpbg_weighted, step_weight;
    long long nlocal;            // units of this rank: ceil((nunits - rank) / world)
    long long nunits;            // work units: reference work items (RNG mode 0) or packets / cells (mode 1)
    int rank, world;
    int max_steps;               // guard against packets that never leave (counted as stuck)
    int deposit;                 // DepositMode
    int refill;                  // stream kernel: refill a warp when at least this many lanes are idle
    int ref_geometry;            // regular grids: 1 = step with the reference's GetStep arithmetic instead of the DDA
    int nav_hops;                // octree walk kernel: navigation rounds (climb / cross / descend) per loop iteration
    int sc_batch;                // production kernels: scatter when this many lanes of the warp wait at a scattering point
    int agg_steps;               // stream kernel: combine lanes only while a packet is younger than this
    int tile_x0, tile_y0, tile_z0;   // DEP_TILE: root-grid origin of the shared-memory tile
    int tile_lo, tile_span;          // first root-cell index of the tile's z-slab and the slab's length
    int tile_inside;                 // DEP_TILE: the point source lies inside the grid (and so inside the tile)
    unsigned long long *counters;   // packets, steps, scatterings, stuck
    unsigned long long *work;       // stream kernel: next unit to hand out
    MwcLaunch mwc;
    PhiloxLaunch phx;
};

void launch_sim(const SimArgs &A, int rng_mode, int blocks, int threads, cudaStream_t stream);
bool sim_domains_eligible(const SimArgs &A, int rng_mode);       // can this launch run domain by domain (launch_sim_emit / launch_sim_domain)?
void launch_sim_emit(const SimArgs &A, long long nunits, cudaStream_t stream);    // emits units [unit0, unit0 + nunits) into the domain queues
void launch_sim_domain(const SimArgs &A, int blocks, int threads, cudaStream_t stream);   // propagates q_in[0 .. nlocal) inside the domain
void launch_queue_sort(const QPk *q, long long n, QPk *out, unsigned *hist /* 32768 */, const int lo[3], cudaStream_t stream);   // counting sort by entry block and direction octant
bool sim_tile_pass_eligible(const SimArgs &A, int rng_mode);      // emission + the steps inside the shared-memory tile as a pass of their own (domain mode: instead of the emission pass)
bool sim_two_pass_eligible(const SimArgs &A, int rng_mode);       // point-source launch as tile pass + plain-add pass (sim.cu)
void launch_sim_tile_pass(const SimArgs &A, int blocks, int threads, cudaStream_t stream);   // emits units [unit0, unit0 + nlocal), parks them at the border of the tile (box dom_lo .. dom_hi)
void sim_note_two_pass();            // sets the name sim_last_kernel() reports for a two-pass launch
void launch_sim_cleanup(const SimArgs &A, int blocks, int threads, cudaStream_t stream);  // finishes q_in[0 .. nlocal) on the whole grid (general kernel, adds straight into TABS / INT)
void launch_fold_acc(const SimArgs &A, cudaStream_t stream);     // TABS += acc*TW*ADHOC; INT += acc; acc = 0
void launch_brick_permute(const SimArgs &A, float *dens_brick, cudaStream_t stream);   // DENS -> (domain-major) 2x2x2-brick order, A.dsplit / A.dsize
bool sim_kappa_eligible(const SimArgs &A, int rng_mode);          // per-cell opacities can take the look-ahead kernel (needs A.kappa)
void launch_kappa(const SimArgs &A, cudaStream_t stream);        // fills A.kappa from DENS and OPT (bricked when A.brick)
bool sim_uses_bricks(const SimArgs &A, int rng_mode);            // does launch_sim() take the bricked kernel?
const char *sim_last_kernel();       // name of the packet kernel dispatched last by this thread
int  sim_blocks_per_sm(int rng_mode, bool octree, bool dbl, int threads);
