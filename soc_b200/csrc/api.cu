// soc_b200 -- the C ABI (include/soc_b200.h): context, device buffers, launches.
//
// One context = one CUDA device + one in-order stream, mirroring the reference's single pyopencl
// queue (ASOC_aux.py:1188-1256).  Uploads and launches are asynchronous on that stream; downloads and
// soc_sync() synchronise.  Errors never abort: every entry point returns a negative soc_status and
// leaves a message for soc_last_error().
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <climits>
#include <new>
#include "sim.cuh"
#include "map.cuh"
#include "aux.cuh"
#include "sca.cuh"

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(SOC_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define NEED_CTX(c) do { if ((c) == nullptr) return fail(SOC_ERR_ARG, "null context"); \
    cudaError_t e_ = cudaSetDevice((c)->device); if (e_ != cudaSuccess) \
    return fail(SOC_ERR_CUDA, "cudaSetDevice(%d): %s", (c)->device, cudaGetErrorString(e_)); } while (0)

struct DevBuf { void *ptr; size_t bytes; };

struct soc_context {
    int device, sms;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    bool timed;
    DevBuf buf[SOC_BUF_COUNT];
    unsigned long long *counters;      // device: packets, steps, scatterings, stuck, peels, work, -, -
    float *acc; size_t acc_bytes;      // per-launch scratch accumulator of the stream kernels (all zero between launches)
    void *scratch; size_t scratch_bytes;   // temporary device array of soc_emission2
    int *nbr;                          // octrees: neighbour table [6*cells] of {cell, density} pairs (linkwalk.cuh)
    int use_nbr;
    float *dens_brick;                 // regular grids with even dimensions: DENS in 2x2x2-brick order (lean kernel)
    int layout;                        // 1 = use the bricked copy where the kernel supports it
    int brick_ns[3];                   // domains per axis the bricked copy is laid out for (0 = not permuted yet)
    int pend;                          // 1 = merged deposits (vector reds) in the lean kernel
    float2 *kappa; size_t kappa_cells; // WITH_ABU on regular grids: (kabs*n, ksca*n) per cell, rebuilt before every launch
    int ahead;                         // 1 = look-ahead variant of the lean kernel (geometry one cell ahead, cp.async density ring)
    int domains;                       // domain-tiled propagation: 0 auto, < 0 off, > 0 forced box edge (soc_set_domains)
    int two_pass;                      // point-source launches with the shared-memory tile in two passes (sim_launch_two_pass)
    QPk *queues; size_t queue_bytes;   // domain mode: packet queues of all domains
    unsigned *q_tail, *h_tail;         // queue lengths on the device / in pinned host memory
    QPk *q_sorted; size_t q_sorted_bytes; unsigned *q_hist;   // domain mode: one queue sorted by entry block and direction
    unsigned long long domain_launches, domain_parked;   // statistics of the last domain-mode launch
    unsigned long long launches;
    soc_params P;
    bool have_params, have_grid;
    GridDesc G;
    int rng_mode, rank, world;
    int deposit, refill, agg_steps, geometry, sc_batch, nav_hops;
    int fabs_nfreq;
    RoiDesc roi; bool have_roi;
    uint64_t pow2k[26];
};

static const char *buf_name(int b) {
    static const char *n[SOC_BUF_COUNT] = { "DENS", "PAR", "TABS", "XAB", "INT", "INTX", "INTY", "INTZ", "EMIT", "EMWEI",
        "OPT", "DSC", "CSC", "PSPOS", "PS", "XPS_NSIDE", "XPS_SIDE", "XPS_AREA", "HPBG", "HPBGP", "MAP", "SAVETAU", "OUT",
        "ODIR", "ORA", "ODE", "TTT", "TNEW", "FABS", "ABU", "ABSV", "SCAV", "ROI_LOAD", "ROI_SAVE" };
    return (b >= 0 && b < SOC_BUF_COUNT) ? n[b] : "?";
}

static int ensure(soc_context *c, int b, size_t bytes) {
    if (b < 0 || b >= SOC_BUF_COUNT) return fail(SOC_ERR_ARG, "unknown buffer id %d", b);
    if (c->buf[b].ptr != nullptr && c->buf[b].bytes == bytes) return SOC_OK;
    if (c->buf[b].ptr != nullptr) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->buf[b].ptr)); c->buf[b].ptr = nullptr; c->buf[b].bytes = 0; }
    if (bytes == 0) return SOC_OK;
    CU(cudaMalloc(&c->buf[b].ptr, bytes));
    c->buf[b].bytes = bytes;
    return SOC_OK;
}
static int ensure_zeroed(soc_context *c, int b, size_t bytes) {
    bool fresh = !(c->buf[b].ptr != nullptr && c->buf[b].bytes == bytes);
    int r = ensure(c, b, bytes);
    if (r != SOC_OK) return r;
    if (fresh && bytes) CU(cudaMemsetAsync(c->buf[b].ptr, 0, bytes, c->stream));
    return SOC_OK;
}
template <class T> static T *dptr(soc_context *c, int b) { return reinterpret_cast<T *>(c->buf[b].ptr); }
static int need(soc_context *c, int b, size_t min_bytes, const char *who) {
    if (c->buf[b].ptr == nullptr || c->buf[b].bytes < min_bytes)
        return fail(SOC_ERR_STATE, "%s: buffer %s missing or too small (%zu < %zu bytes)", who, buf_name(b), c->buf[b].bytes, min_bytes);
    return SOC_OK;
}

extern "C" {

const char *soc_last_error(void) { return g_err; }
int soc_version(void) { return 100; }

int soc_create(int device_ordinal, soc_context **out) {
    if (out == nullptr) return fail(SOC_ERR_ARG, "soc_create: null output pointer");
    *out = nullptr;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device_ordinal < 0 || device_ordinal >= ndev) return fail(SOC_ERR_ARG, "soc_create: device %d of %d", device_ordinal, ndev);
    CU(cudaSetDevice(device_ordinal));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device_ordinal));
    if (prop.major < 10) return fail(SOC_ERR_UNSUPPORTED, "soc_create: device %s is sm_%d%d, this library is built for sm_100a", prop.name, prop.major, prop.minor);
    soc_context *c = new (std::nothrow) soc_context();
    if (c == nullptr) return fail(SOC_ERR_ARG, "out of host memory");
    memset(c, 0, sizeof(*c));
    c->device = device_ordinal; c->sms = prop.multiProcessorCount;
    c->rng_mode = SOC_RNG_PACKET; c->rank = 0; c->world = 1;
    c->deposit = DEP_TILE; c->refill = 8; c->agg_steps = 24; c->sc_batch = 0;        // 0 = by grid type
    c->nav_hops = 1; c->layout = 1;
    c->use_nbr = 1;
    if (const char *e = getenv("SOC_NBR")) c->use_nbr = atoi(e) != 0;                                                  // tuning knob
    c->pend = 0;          // measured on the bench step: 58.9 ms with, 59.0 ms without -- the merged reds trade L2 work for issue slots
    if (const char *e = getenv("SOC_PEND")) c->pend = atoi(e) != 0;                                                    // tuning knob
    c->ahead = 1;
    if (const char *e = getenv("SOC_AHEAD")) c->ahead = atoi(e);                                                       // tuning knob
    if (const char *e = getenv("SOC_LAYOUT")) c->layout = atoi(e) != 0;                                               // tuning knob
    if (const char *e = getenv("SOC_DOMAINS")) c->domains = atoi(e);                                                  // tuning knob
    c->two_pass = 1;
    if (const char *e = getenv("SOC_TWO_PASS")) c->two_pass = atoi(e);                                                // tuning knob
    if (const char *e = getenv("SOC_NAV_HOPS")) { int v = atoi(e); if (v >= 1 && v <= 8) c->nav_hops = v; }          // tuning knob
    if (const char *e = getenv("SOC_SC_BATCH")) { int v = atoi(e); if (v >= 1 && v <= 32) c->sc_batch = v; }   // tuning knob
    if (const char *e = getenv("SOC_L2_FETCH")) { int v = atoi(e); if (v == 32 || v == 64 || v == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v); }   // tuning knob
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    CU(cudaMalloc(&c->counters, 8 * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(c->counters, 0, 8 * sizeof(unsigned long long), c->stream));
    uint64_t p = mwc_powmod(MWC_A, 274877906944ull);            // A^(2^38)
    for (int k = 0; k < 26; k++) { c->pow2k[k] = p; p = mwc_mulmod(p, p); }
    *out = c;
    return SOC_OK;
}

int soc_destroy(soc_context *c) {
    if (c == nullptr) return SOC_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int b = 0; b < SOC_BUF_COUNT; b++) if (c->buf[b].ptr) cudaFree(c->buf[b].ptr);
    cudaFree(c->counters);
    if (c->acc) cudaFree(c->acc);
    if (c->dens_brick) cudaFree(c->dens_brick);
    if (c->nbr) cudaFree(c->nbr);
    if (c->scratch) cudaFree(c->scratch);
    if (c->kappa) cudaFree(c->kappa);
    if (c->queues) cudaFree(c->queues);
    if (c->q_tail) cudaFree(c->q_tail);
    if (c->q_sorted) cudaFree(c->q_sorted);
    if (c->q_hist) cudaFree(c->q_hist);
    if (c->h_tail) cudaFreeHost(c->h_tail);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    delete c;
    return SOC_OK;
}

int soc_sync(soc_context *c) {
    NEED_CTX(c);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return SOC_OK;
}

int soc_set_params(soc_context *c, const soc_params *p) {
    NEED_CTX(c);
    if (p == nullptr) return fail(SOC_ERR_ARG, "soc_set_params: null");
    if (p->dir_weight || p->do_split)
        return fail(SOC_ERR_UNSUPPORTED, "soc_set_params: DIR_WEIGHT / DO_SPLIT are not implemented");
    if (p->roi_flags < 0 || p->roi_flags > 7) return fail(SOC_ERR_ARG, "soc_set_params: roi_flags=%d", p->roi_flags);
    if (p->with_msf && (!p->with_abu || p->ndust < 1))
        return fail(SOC_ERR_ARG, "soc_set_params: WITH_MSF needs WITH_ABU and NDUST >= 1 (ASOC.py:1167)");
    if (p->mirror < 0 || p->mirror > 63) return fail(SOC_ERR_ARG, "soc_set_params: MIRROR=%d", p->mirror);
    if (p->map_interpolation < 0 || p->map_interpolation > 2) return fail(SOC_ERR_ARG, "soc_set_params: MAP_INTERPOLATION=%d", p->map_interpolation);
    if (p->ps_method == 3 || p->ps_method < 0 || p->ps_method > 5)
        return fail(SOC_ERR_UNSUPPORTED, "soc_set_params: PS_METHOD %d (the reference's method 3 does not compile either)", p->ps_method);
    if (p->bins < 2) return fail(SOC_ERR_ARG, "soc_set_params: BINS=%d", p->bins);
    if (p->no_ps < 1) return fail(SOC_ERR_ARG, "soc_set_params: NO_PS must be >= 1 (ASOC.py:344 passes max(1,NO_PS))");
    if (p->ref_quirks < 0 || p->ref_quirks > 3) return fail(SOC_ERR_ARG, "soc_set_params: ref_quirks=%d", p->ref_quirks);
    if (p->save_intensity < 0 || p->save_intensity > 2) return fail(SOC_ERR_ARG, "soc_set_params: SAVE_INTENSITY=%d", p->save_intensity);
    if (!(p->length > 0.0f) || !(p->factor > 0.0f) || !(p->adhoc > 0.0f)) return fail(SOC_ERR_ARG, "soc_set_params: LENGTH, FACTOR, ADHOC must be positive");
    c->P = *p;
    c->have_params = true;
    return SOC_OK;
}

int soc_set_grid(soc_context *c, int32_t nx, int32_t ny, int32_t nz, int32_t levels, int64_t cells,
                 const int32_t *lcells, const int32_t *off, const float *dens) {
    NEED_CTX(c);
    if (nx < 1 || ny < 1 || nz < 1 || levels < 1 || levels > SOC_MAX_LEVELS || lcells == nullptr || off == nullptr || dens == nullptr)
        return fail(SOC_ERR_ARG, "soc_set_grid: bad dimensions (levels must be 1..%d)", SOC_MAX_LEVELS);
    if (cells < 1 || cells > INT_MAX) return fail(SOC_ERR_ARG, "soc_set_grid: CELLS=%lld exceeds the int32 cell index of the file formats", (long long)cells);
    long long sum = 0;
    for (int l = 0; l < levels; l++) {
        if (off[l] != sum || lcells[l] < 0) return fail(SOC_ERR_ARG, "soc_set_grid: OFF/LCELLS inconsistent at level %d", l);
        if (l > 0 && (lcells[l] % 8) != 0) return fail(SOC_ERR_ARG, "soc_set_grid: level %d does not consist of octets", l);
        sum += lcells[l];
    }
    if (sum != cells || (long long)nx * ny * nz != lcells[0]) return fail(SOC_ERR_ARG, "soc_set_grid: sum(LCELLS) != CELLS or LCELLS[0] != NX*NY*NZ");
    GridDesc &G = c->G;
    memset(&G, 0, sizeof(G));
    G.nx = nx; G.ny = ny; G.nz = nz; G.levels = levels; G.cells = (int)cells; G.nxyz = nx * ny * nz;
    G.area = 2 * (nx * ny + ny * nz + nz * nx);
    G.dbl_sim = nx > (levels < 3 ? 399 : 100);            // DIMLIM, kernel_ASOC_aux.c:25-37
    G.dbl_map = nx > 100;                                 // kernel_ASOC_map.c:302
    for (int l = 0; l < levels; l++) { G.off[l] = off[l]; G.lcells[l] = lcells[l]; }
    int r = ensure(c, SOC_BUF_DENS, (size_t)cells * 4);
    if (r != SOC_OK) return r;
    CU(cudaMemcpyAsync(c->buf[SOC_BUF_DENS].ptr, dens, (size_t)cells * 4, cudaMemcpyHostToDevice, c->stream));
    size_t npar = (size_t)(cells - G.nxyz);
    r = ensure(c, SOC_BUF_PAR, (npar > 0 ? npar : 1) * 4);
    if (r != SOC_OK) return r;
    G.dens = dptr<float>(c, SOC_BUF_DENS);
    G.par = dptr<int>(c, SOC_BUF_PAR);
    if (levels > 1) { launch_parents(G, dptr<int>(c, SOC_BUF_PAR), c->stream); c->launches += levels - 1; }
    if (c->nbr) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->nbr)); c->nbr = nullptr; }
    if (levels > 1 && c->use_nbr && cells < (1LL << 27)) {
        // neighbour table of the production octree kernel: 48 B per cell (six {cell, density} pairs), built once per grid;
        // 2^27 cells = 6 GiB of table
        if (cudaMalloc(&c->nbr, (size_t)cells * 48) == cudaSuccess) { launch_neighbours(G, c->nbr, c->stream); c->launches++; }
        else { c->nbr = nullptr; cudaGetLastError(); }
    }
    if (c->dens_brick) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->dens_brick)); c->dens_brick = nullptr; }
    if (levels == 1 && nx % 2 == 0 && ny % 2 == 0 && nz % 2 == 0) {
        CU(cudaMalloc(&c->dens_brick, (size_t)cells * 4));       // permuted by the first launch that uses it (sim_launch)
    }
    c->brick_ns[0] = c->brick_ns[1] = c->brick_ns[2] = 0;
    CU(cudaGetLastError());
    c->have_grid = true;
    return SOC_OK;
}

static size_t roi_elems(const int n[3]) { return (size_t)n[0] * n[1] + (size_t)n[1] * n[2] + (size_t)n[2] * n[0]; }

int soc_set_roi(soc_context *c, const int32_t roi[6], int roi_step, int roi_nside, const int32_t roi_dim[3]) {
    NEED_CTX(c);
    if (!c->have_params || !c->have_grid) return fail(SOC_ERR_STATE, "soc_set_roi: grid and params first");
    if (roi == nullptr || roi_nside < 1 || roi_nside > 1024) return fail(SOC_ERR_ARG, "soc_set_roi: bad arguments");
    const int flags = c->P.roi_flags;
    memset(&c->roi, 0, sizeof(c->roi));
    c->roi.flags = flags; c->roi.step = roi_step; c->roi.nside = roi_nside;
    for (int k = 0; k < 6; k++) c->roi.lim[k] = roi[k];
    for (int k = 0; k < 3; k++) c->roi.dim[k] = roi_dim ? roi_dim[k] : 1;
    if (flags & 6) {
        const int dim[3] = { c->G.nx, c->G.ny, c->G.nz };
        for (int k = 0; k < 3; k++)
            if (roi[2 * k] < 0 || roi[2 * k + 1] >= dim[k] || roi[2 * k] > roi[2 * k + 1]) return fail(SOC_ERR_ARG, "soc_set_roi: limits outside the root grid");
    }
    if (flags & 1) { if (!roi_dim || roi_dim[0] < 1 || roi_dim[1] < 1 || roi_dim[2] < 1) return fail(SOC_ERR_ARG, "soc_set_roi: ROI_DIM"); }
    if (flags & 2) {
        if (roi_step < 1) return fail(SOC_ERR_ARG, "soc_set_roi: ROI_STEP=%d", roi_step);
        const int n[3] = { (roi[1] - roi[0] + 1) * roi_step, (roi[3] - roi[2] + 1) * roi_step, (roi[5] - roi[4] + 1) * roi_step };
        int r = soc_clear(c, SOC_BUF_ROI_SAVE, roi_elems(n) * 12 * roi_nside * roi_nside * 4);
        if (r != SOC_OK) return r;
    }
    c->have_roi = true;
    return SOC_OK;
}

// RoiDesc with the device pointers, checked against the options in force
static int roi_args(soc_context *c, RoiDesc &R, bool need_load, const char *who) {
    memset(&R, 0, sizeof(R));
    if (!c->P.roi_flags) return SOC_OK;
    if (!c->have_roi) return fail(SOC_ERR_STATE, "%s: soc_set_roi first (roi_flags=%d)", who, c->P.roi_flags);
    R = c->roi;
    R.load = dptr<float>(c, SOC_BUF_ROI_LOAD); R.save = dptr<float>(c, SOC_BUF_ROI_SAVE);
    const size_t npix = (size_t)12 * R.nside * R.nside * 4;
    if (need_load) {
        int r = need(c, SOC_BUF_ROI_LOAD, roi_elems(R.dim) * npix, who);
        if (r != SOC_OK) return r;
    }
    if ((R.flags & 2) && R.save == nullptr) return fail(SOC_ERR_STATE, "%s: ROI_SAVE buffer missing", who);
    return SOC_OK;
}

int soc_set_rng_mode(soc_context *c, int mode) {
    NEED_CTX(c);
    if (mode != SOC_RNG_REFERENCE && mode != SOC_RNG_PACKET) return fail(SOC_ERR_ARG, "soc_set_rng_mode: %d", mode);
    c->rng_mode = mode;
    return SOC_OK;
}

int soc_set_shard(soc_context *c, int rank, int world) {
    NEED_CTX(c);
    if (world < 1 || rank < 0 || rank >= world) return fail(SOC_ERR_ARG, "soc_set_shard: rank %d of %d", rank, world);
    c->rank = rank; c->world = world;
    return SOC_OK;
}

int soc_set_tuning(soc_context *c, int deposit_mode, int refill_lanes, int aggregate_steps) {
    NEED_CTX(c);
    if (deposit_mode < DEP_RED || deposit_mode > DEP_TILE) return fail(SOC_ERR_ARG, "soc_set_tuning: deposit mode %d", deposit_mode);
    if (refill_lanes < 1 || refill_lanes > 32) return fail(SOC_ERR_ARG, "soc_set_tuning: refill lanes %d", refill_lanes);
    c->deposit = deposit_mode; c->refill = refill_lanes; c->agg_steps = aggregate_steps;
    return SOC_OK;
}

int soc_set_layout(soc_context *c, int mode) {
    NEED_CTX(c);
    if (mode != 0 && mode != 1) return fail(SOC_ERR_ARG, "soc_set_layout: %d", mode);
    c->layout = mode;
    return SOC_OK;
}

int soc_set_domains(soc_context *c, int edge) {
    NEED_CTX(c);
    if (edge > 0 && (edge % 2 != 0 || edge < 4)) return fail(SOC_ERR_ARG, "soc_set_domains: box edge %d must be even and >= 4", edge);
    c->domains = edge;
    return SOC_OK;
}

int soc_set_geometry(soc_context *c, int mode) {
    NEED_CTX(c);
    if (mode != 0 && mode != 1) return fail(SOC_ERR_ARG, "soc_set_geometry: %d", mode);
    c->geometry = mode;
    return SOC_OK;
}

int soc_upload(soc_context *c, int b, const void *host, size_t nbytes) {
    NEED_CTX(c);
    if (b == SOC_BUF_DENS || b == SOC_BUF_PAR) return fail(SOC_ERR_ARG, "soc_upload: %s is set by soc_set_grid", buf_name(b));
    if (host == nullptr && nbytes) return fail(SOC_ERR_ARG, "soc_upload(%s): null host pointer", buf_name(b));
    if (b == SOC_BUF_OPT && c->have_params && c->P.opt_is_half && nbytes) {          // half values in, float on the device
        if (c->scratch_bytes < nbytes) {
            if (c->scratch) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->scratch)); c->scratch = nullptr; c->scratch_bytes = 0; }
            CU(cudaMalloc(&c->scratch, nbytes));
            c->scratch_bytes = nbytes;
        }
        int r = ensure(c, b, 2 * nbytes);
        if (r != SOC_OK) return r;
        CU(cudaMemcpyAsync(c->scratch, host, nbytes, cudaMemcpyHostToDevice, c->stream));
        launch_half_to_float(c->scratch, dptr<float>(c, b), (long long)(nbytes / 2), c->stream);
        c->launches++;
        CU(cudaGetLastError());
        return SOC_OK;
    }
    int r = ensure(c, b, nbytes);
    if (r != SOC_OK) return r;
    if (nbytes) CU(cudaMemcpyAsync(c->buf[b].ptr, host, nbytes, cudaMemcpyHostToDevice, c->stream));
    return SOC_OK;
}

int soc_download(soc_context *c, int b, void *host, size_t nbytes) {
    NEED_CTX(c);
    if (b < 0 || b >= SOC_BUF_COUNT) return fail(SOC_ERR_ARG, "unknown buffer id %d", b);
    if (host == nullptr && nbytes) return fail(SOC_ERR_ARG, "soc_download(%s): null host pointer", buf_name(b));
    if (c->buf[b].ptr == nullptr || c->buf[b].bytes < nbytes)
        return fail(SOC_ERR_ARG, "soc_download(%s): %zu bytes requested, buffer holds %zu", buf_name(b), nbytes, c->buf[b].bytes);
    if (nbytes) CU(cudaMemcpyAsync(host, c->buf[b].ptr, nbytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SOC_OK;
}

int soc_clear(soc_context *c, int b, size_t nbytes) {
    NEED_CTX(c);
    if (b == SOC_BUF_DENS || b == SOC_BUF_PAR) return fail(SOC_ERR_ARG, "soc_clear: %s is set by soc_set_grid", buf_name(b));
    int r = ensure(c, b, nbytes);
    if (r != SOC_OK) return r;
    if (nbytes) CU(cudaMemsetAsync(c->buf[b].ptr, 0, nbytes, c->stream));
    return SOC_OK;
}

void *soc_device_ptr(soc_context *c, int b, size_t *nbytes) {
    if (c == nullptr || b < 0 || b >= SOC_BUF_COUNT) { if (nbytes) *nbytes = 0; return nullptr; }
    if (nbytes) *nbytes = c->buf[b].bytes;
    return c->buf[b].ptr;
}

void *soc_stream(soc_context *c) { return c ? (void *)c->stream : nullptr; }

static bool wants_int(const soc_params &P) { return P.save_intensity == 1 || P.save_intensity == 2 || P.noabsorbed == 0; }

int soc_zero_amc(soc_context *c, int tag) {
    NEED_CTX(c);
    if (!c->have_grid || !c->have_params) return fail(SOC_ERR_STATE, "soc_zero_amc: grid and params first");
    const size_t n = (size_t)c->G.cells * 4;
    int r;
    if (tag == 0) {
        if ((r = soc_clear(c, SOC_BUF_TABS, n)) != SOC_OK) return r;
        if (c->P.with_ali && (r = soc_clear(c, SOC_BUF_XAB, n)) != SOC_OK) return r;
    } else {
        if (wants_int(c->P) && (r = soc_clear(c, SOC_BUF_INT, n)) != SOC_OK) return r;
        if (c->P.save_intensity == 2) {
            if ((r = soc_clear(c, SOC_BUF_INTX, n)) != SOC_OK) return r;
            if ((r = soc_clear(c, SOC_BUF_INTY, n)) != SOC_OK) return r;
            if ((r = soc_clear(c, SOC_BUF_INTZ, n)) != SOC_OK) return r;
        }
    }
    return SOC_OK;
}

// kernel_ASOC.c:77: base offset of the MWC64X streams from the float seed
static uint64_t seed_to_base(float seed) {
    volatile float a = seed * 7.0f;
    volatile float b = a * 3.1415926535897f;
    volatile float f = fmodf(b, 1.0f);
    volatile float g = f * (float)4294967296L;
    return (uint64_t)g;
}

static int sim_common(soc_context *c, SimArgs &A, int kind, int batch, float seed, float kabs, float ksca, float tw,
                      int global, const char *who) {
    if (!c->have_grid || !c->have_params) return fail(SOC_ERR_STATE, "%s: soc_set_grid and soc_set_params first", who);
    if (batch < 1 || global < 1) return fail(SOC_ERR_ARG, "%s: batch=%d global=%d", who, batch, global);
    const soc_params &P = c->P;
    const size_t n = (size_t)c->G.cells * 4;
    int r;
    memset(&A, 0, sizeof(A));
    A.G = c->G;
    if ((r = ensure_zeroed(c, SOC_BUF_TABS, n)) != SOC_OK) return r;
    if (P.with_ali && (r = ensure_zeroed(c, SOC_BUF_XAB, n)) != SOC_OK) return r;
    A.use_int = wants_int(P) ? 1 : 0;
    A.save_int2 = P.save_intensity == 2;
    if (A.use_int && (r = ensure_zeroed(c, SOC_BUF_INT, n)) != SOC_OK) return r;
    if (A.save_int2) {
        if ((r = ensure_zeroed(c, SOC_BUF_INTX, n)) != SOC_OK) return r;
        if ((r = ensure_zeroed(c, SOC_BUF_INTY, n)) != SOC_OK) return r;
        if ((r = ensure_zeroed(c, SOC_BUF_INTZ, n)) != SOC_OK) return r;
    }
    if ((r = need(c, SOC_BUF_CSC, (size_t)P.bins * 4, who)) != SOC_OK) return r;
    if (P.with_abu && (r = need(c, SOC_BUF_OPT, 2 * n, who)) != SOC_OK) return r;
    if (P.with_msf) {
        if ((r = need(c, SOC_BUF_ABU, n * P.ndust, who)) != SOC_OK) return r;
        if ((r = need(c, SOC_BUF_SCAV, (size_t)P.ndust * 4, who)) != SOC_OK) return r;
        if ((r = need(c, SOC_BUF_CSC, (size_t)P.bins * P.ndust * 4, who)) != SOC_OK) return r;
    }
    A.abu = dptr<float>(c, SOC_BUF_ABU); A.scav = dptr<float>(c, SOC_BUF_SCAV);
    A.with_msf = P.with_msf; A.ndust = P.with_msf ? P.ndust : 1; A.mirror = P.mirror;
    if ((r = roi_args(c, A.roi, kind == SIM_ROI, who)) != SOC_OK) return r;
    A.roi.flags &= 3;
    if (kind == SIM_HP) A.roi.flags &= ~2;                     // SimRAM_HP has no ROI_SAVE (kernel_ASOC.c:831-854)
    A.tabs = dptr<float>(c, SOC_BUF_TABS); A.xab = dptr<float>(c, SOC_BUF_XAB);
    A.inten = dptr<float>(c, SOC_BUF_INT); A.intx = dptr<float>(c, SOC_BUF_INTX);
    A.inty = dptr<float>(c, SOC_BUF_INTY); A.intz = dptr<float>(c, SOC_BUF_INTZ);
    A.emit = dptr<float>(c, SOC_BUF_EMIT); A.emwei = dptr<float>(c, SOC_BUF_EMWEI); A.opt = dptr<float>(c, SOC_BUF_OPT);
    A.dsc = dptr<float>(c, SOC_BUF_DSC); A.csc = dptr<float>(c, SOC_BUF_CSC);
    A.pspos = dptr<float>(c, SOC_BUF_PSPOS); A.ps = dptr<float>(c, SOC_BUF_PS); A.xps_area = dptr<float>(c, SOC_BUF_XPS_AREA);
    A.xps_nside = dptr<int>(c, SOC_BUF_XPS_NSIDE); A.xps_side = dptr<int>(c, SOC_BUF_XPS_SIDE);
    A.hpbg = dptr<float>(c, SOC_BUF_HPBG); A.hpbgp = dptr<float>(c, SOC_BUF_HPBGP);
    A.kabs = kabs; A.ksca = ksca; A.tw = tw; A.adhoc = P.adhoc; A.sw_a = P.sw_a; A.sw_b = P.sw_b;
    A.kind = kind; A.batch = batch; A.global = global;
    A.bins = P.bins; A.no_ps = P.no_ps; A.ps_method = P.ps_method; A.with_abu = P.with_abu; A.with_ali = P.with_ali;
    A.use_emweight = P.use_emweight; A.hpbg_weighted = P.hpbg_weighted; A.step_weight = P.step_weight;
    A.rank = c->rank; A.world = c->world;
    long long ms = 100LL * ((long long)c->G.nx + c->G.ny + c->G.nz) << (c->G.levels - 1);
    A.max_steps = (int)(ms > INT_MAX ? INT_MAX : ms);
    A.deposit = c->deposit; A.refill = c->refill; A.agg_steps = c->agg_steps; A.ref_geometry = c->geometry; A.nav_hops = c->nav_hops; A.sc_batch = c->sc_batch > 0 ? c->sc_batch : (c->G.levels > 1 ? 3 : 1);
    A.counters = c->counters; A.work = c->counters + 5;
    // stream layouts
    A.mwc.base_offset = seed_to_base(seed);
    A.mwc.base_state = mwc_mulmod(MWC_BASEID, mwc_powmod(MWC_A, A.mwc.base_offset));
    for (int k = 0; k < 26; k++) A.mwc.pow2k[k] = c->pow2k[k];
    uint32_t sb; memcpy(&sb, &seed, 4);
    A.phx.k0 = sb; A.phx.k1 = 0x534F4332u; A.phx.tag = (uint32_t)kind; A.phx.pad = 0;
    return SOC_OK;
}

// Domain-tiled propagation (soc_set_domains, sim.cuh): emission pass into the queues of the domains, then the domain with
// the longest queue is processed until every queue is empty.  The last few packets (ping-pong between domains after many
// scatterings) are finished on the whole grid by the general kernel, which parks nothing.  A.dsplit / A.dsize hold the
// boxes; DENS (and KAPPA) are already laid out for them.
static int sim_launch_domains(soc_context *c, SimArgs &A, int blocks, int threads) {
    const int dim[3] = { A.G.nx, A.G.ny, A.G.nz };
    const int *ns = A.dsplit, *ds = A.dsize;
    const int D = ns[0] * ns[1] * ns[2];
    if (D > 4096) return fail(SOC_ERR_UNSUPPORTED, "soc_set_domains: %d domains (at most 4096)", D);
    const long long dcells = (long long)ds[0] * ds[1] * ds[2];
    // chunk of work units in flight: every queue must be able to hold all of them
    long long chunk = 1LL << 25;
    if (const char *e = getenv("SOC_DOMAIN_CHUNK")) { long long v = atoll(e); if (v >= 1024) chunk = v; }      // tuning knob
    const long long carry_cap = 1LL << 20;                 // room for the packets carried over into the next chunk (see below)
    if ((size_t)((chunk < A.nlocal ? chunk : A.nlocal) + carry_cap) * D * sizeof(QPk) > c->queue_bytes) {
        // the queues have to grow: at most half of the free memory (asked only then: cudaMemGetInfo is slow)
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        const size_t budget = (free_b + c->queue_bytes) / 2;
        while (chunk > 65536 && (size_t)(chunk + carry_cap) * D * sizeof(QPk) > budget) chunk >>= 1;
    }
    if (chunk > A.nlocal) chunk = A.nlocal > 0 ? A.nlocal : 1;
    const long long q_cap = chunk + carry_cap;
    const size_t need_b = (size_t)q_cap * D * sizeof(QPk);
    if (c->queue_bytes < need_b) {
        if (c->queues) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->queues)); c->queues = nullptr; c->queue_bytes = 0; }
        CU(cudaMalloc(&c->queues, need_b));
        c->queue_bytes = need_b;
    }
    if (c->q_tail == nullptr) { CU(cudaMalloc(&c->q_tail, 4096 * sizeof(unsigned))); CU(cudaMallocHost(&c->h_tail, 4096 * sizeof(unsigned))); }
    // queues sorted by entry block and direction before they are processed.  Background: yes.  Point source: no -- its packets
    // stream radially, sorted neighbours walk the same cells in step and their adds serialise (512^3 launch 177 ms sorted, 133 ms not)
    int sort = A.kind == SIM_PS ? 0 : 1;
    if (const char *e = getenv("SOC_DOMAIN_SORT")) sort = atoi(e);                                               // tuning knob
    if (sort) {
        const size_t sb = (size_t)q_cap * sizeof(QPk);
        if (c->q_sorted_bytes < sb) {
            if (c->q_sorted) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->q_sorted)); c->q_sorted = nullptr; c->q_sorted_bytes = 0; }
            CU(cudaMalloc(&c->q_sorted, sb));
            c->q_sorted_bytes = sb;
        }
        if (c->q_hist == nullptr) CU(cudaMalloc(&c->q_hist, 32768 * sizeof(unsigned)));
    }
    // below this many parked packets the rest runs over the whole grid: a domain launch with few packets costs its
    // latency (~0.5 ms: one packet after the other through ~250 dependent steps) whatever it holds
    long long cleanup = 1LL << 17;
    if (const char *e = getenv("SOC_DOMAIN_CLEANUP")) cleanup = atoll(e);                                         // tuning knob
    // a launch of several chunks: when this few packets of a chunk are left, the next chunk is emitted on top of them, so
    // that the tail of small launches and the clean-up pass are paid once per launch, not once per chunk
    long long carry = 1LL << 20;
    if (const char *e = getenv("SOC_DOMAIN_CARRY")) carry = atoll(e);                                             // tuning knob
    if (carry < cleanup) carry = cleanup;
    if (carry > carry_cap) carry = carry_cap;
    long long sort_min = 1LL << 16;
    static bool first_pass[4096];
    const int verbose = getenv("SOC_DOMAIN_VERBOSE") ? atoi(getenv("SOC_DOMAIN_VERBOSE")) : 0;
    A.dom = 1; A.q_base = c->queues; A.q_tail = c->q_tail; A.q_cap = q_cap;
    const bool tile_first = c->two_pass && sim_tile_pass_eligible(A, c->rng_mode);
    const int deposit = tile_first ? DEP_RED : A.deposit;      // accumulation engine of the box launches
    const int refill0 = A.refill, agg0 = A.agg_steps, deposit0 = A.deposit;
    int tp_refill = 24, tp_agg = 0;
    if (const char *e = getenv("SOC_TILEPASS_REFILL")) { int v = atoi(e); if (v >= 1 && v <= 32) tp_refill = v; }     // tuning knob
    if (const char *e = getenv("SOC_TILEPASS_AGG")) { int v = atoi(e); if (v >= 0) tp_agg = v; }                      // tuning knob
    const long long nlocal = A.nlocal;
    c->domain_launches = 0; c->domain_parked = 0;
    CU(cudaMemsetAsync(c->q_tail, 0, D * sizeof(unsigned), c->stream));
    for (long long u0 = 0; u0 < nlocal; u0 += chunk) {
        const long long n = nlocal - u0 < chunk ? nlocal - u0 : chunk;
        const bool last_chunk = u0 + chunk >= nlocal;
        A.unit0 = u0; A.q_base = c->queues; A.nlocal = nlocal;
        for (int d = 0; d < D; d++) first_pass[d] = !tile_first;
        if (tile_first) {
            // point source with the shared-memory tile: emission and the steps inside the tile in a pass of their own, which parks
            // the packets at the border of the tile; every box launch then runs the plain-add look-ahead kernel
            const int t0[3] = { A.tile_x0, A.tile_y0, A.tile_z0 };
            A.dom_faces = 0;
            for (int k = 0; k < 3; k++) {
                A.dom_lo[k] = t0[k]; A.dom_hi[k] = t0[k] + SOC_TILE_N - 1;
                if (A.dom_lo[k] == 0) A.dom_faces |= 1 << (2 * k);
                if (A.dom_hi[k] == dim[k] - 1) A.dom_faces |= 2 << (2 * k);
            }
            A.deposit = DEP_TILE; A.nlocal = n; A.refill = tp_refill; A.agg_steps = tp_agg;
            CU(cudaMemsetAsync(A.work, 0, sizeof(unsigned long long), c->stream));
            launch_sim_tile_pass(A, blocks, threads, c->stream);
            A.refill = refill0; A.agg_steps = agg0; A.nlocal = nlocal;
        } else launch_sim_emit(A, n, c->stream);
        c->launches++;
        for (;;) {
            CU(cudaMemcpyAsync(c->h_tail, c->q_tail, D * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            long long total = 0; int best = -1;
            for (int d = 0; d < D; d++) { total += c->h_tail[d]; if (c->h_tail[d] > 0 && (best < 0 || c->h_tail[d] > c->h_tail[best])) best = d; }
            if (total == 0) break;
            if (!last_chunk && total <= carry) break;            // the rest travels with the next chunk
            const bool whole = total <= cleanup && D > 1;
            if (whole && D <= 64) {                  // every queue in one launch of the general kernel
                A.q_nparts = D; A.q_part[0] = 0;
                for (int d = 0; d < D; d++) A.q_part[d + 1] = A.q_part[d] + c->h_tail[d];
                A.nlocal = total;
                long long needb = (total + threads - 1) / threads;
                const int b = (int)(needb < blocks ? (needb < 1 ? 1 : needb) : blocks);
                CU(cudaMemsetAsync(c->q_tail, 0, D * sizeof(unsigned), c->stream));
                CU(cudaMemsetAsync(A.work, 0, sizeof(unsigned long long), c->stream));
                launch_sim_cleanup(A, b, threads, c->stream);
                c->launches++; c->domain_launches++; c->domain_parked += total;
                CU(cudaGetLastError());
                break;                               // nothing is parked by this pass
            }
            for (int d = 0; d < D; d++) {
                if (c->h_tail[d] == 0 || (!whole && d != best)) continue;
                const int dd[3] = { d % ns[0], (d / ns[0]) % ns[1], d / (ns[0] * ns[1]) };
                A.dom_faces = 0;
                for (int k = 0; k < 3; k++) {
                    A.dom_lo[k] = dd[k] * ds[k];
                    A.dom_hi[k] = (dd[k] + 1) * ds[k] - 1;
                    if (A.dom_lo[k] == 0) A.dom_faces |= 1 << (2 * k);
                    if (A.dom_hi[k] == dim[k] - 1) A.dom_faces |= 2 << (2 * k);
                }
                A.dom_base = (int)(d * dcells);
                // the shared-memory tile around the point source only where the domain holds a part of it
                A.deposit = deposit;
                if (deposit == DEP_TILE) {
                    const int t0[3] = { A.tile_x0, A.tile_y0, A.tile_z0 };
                    for (int k = 0; k < 3; k++) if (t0[k] > A.dom_hi[k] || t0[k] + SOC_TILE_N - 1 < A.dom_lo[k]) A.deposit = DEP_RED;
                }
                A.q_in = c->queues + (size_t)d * (size_t)q_cap;
                A.nlocal = c->h_tail[d];
                A.q_nparts = 1; A.q_part[0] = 0; A.q_part[1] = A.nlocal;
                A.q_base = whole ? c->queues + (size_t)d * (size_t)q_cap : c->queues;      // the clean-up kernel addresses queue `part` of q_base
                long long needb = (A.nlocal + threads - 1) / threads;
                const int b = (int)(needb < blocks ? (needb < 1 ? 1 : needb) : blocks);
                CU(cudaMemsetAsync(c->q_tail + d, 0, sizeof(unsigned), c->stream));
                CU(cudaMemsetAsync(A.work, 0, sizeof(unsigned long long), c->stream));
                if (verbose > 1) { CU(cudaEventRecord(c->ev0, c->stream)); }
                // not the emission queue of a point source: all its packets start in one cell, and lanes that walk the same cells
                // in step serialise on their adds (same-address RED / shared-memory atomics)
                const bool fresh = first_pass[d];
                first_pass[d] = false;
                if (!whole && sort && A.nlocal >= (long long)sort_min && !(sort < 3 && fresh && A.kind == SIM_PS) && !(sort == 2 && A.kind == SIM_PS) &&
                    !(sort == 4 && fresh)) {
                    launch_queue_sort(A.q_in, A.nlocal, c->q_sorted, c->q_hist, A.dom_lo, c->stream);
                    A.q_in = c->q_sorted;
                    c->launches += 3;
                }
                if (whole) launch_sim_cleanup(A, b, threads, c->stream);
                else       launch_sim_domain(A, b, threads, c->stream);
                if (verbose > 1) {
                    CU(cudaEventRecord(c->ev1, c->stream)); CU(cudaEventSynchronize(c->ev1));
                    float ms = 0.0f; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
                    fprintf(stderr, "soc_b200:   domain %d%s: %u packets, dep %d, %.3f ms\n", d, whole ? " (whole grid)" : "", c->h_tail[d], A.deposit, ms);
                }
                c->launches++; c->domain_launches++; c->domain_parked += c->h_tail[d];
            }
            CU(cudaGetLastError());
        }
    }
    A.deposit = deposit0; A.nlocal = nlocal; A.dom = 0;
    if (verbose)
        fprintf(stderr, "soc_b200: %d domains of %dx%dx%d cells, chunk %lld, %llu domain launches, %llu packet visits for %lld packets\n",
                D, ds[0], ds[1], ds[2], chunk, c->domain_launches, c->domain_parked, nlocal);
    return SOC_OK;
}

// Two-pass point-source launch (sim.cu, sim_two_pass_eligible): per chunk of packets the tile pass (emission, the steps inside
// the shared-memory tile, packets parked at its border), then the plain-add look-ahead kernel over the whole grid from the queue.
static int sim_launch_two_pass(soc_context *c, SimArgs &A, int blocks, int threads) {
    const int dim[3] = { A.G.nx, A.G.ny, A.G.nz };
    long long chunk = 1LL << 25;
    if (const char *e = getenv("SOC_DOMAIN_CHUNK")) { long long v = atoll(e); if (v >= 1024) chunk = v; }      // tuning knob
    if (chunk > A.nlocal) chunk = A.nlocal > 0 ? A.nlocal : 1;
    const size_t need_b = (size_t)chunk * sizeof(QPk);
    if (c->queue_bytes < need_b) {
        if (c->queues) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->queues)); c->queues = nullptr; c->queue_bytes = 0; }
        CU(cudaMalloc(&c->queues, need_b));
        c->queue_bytes = need_b;
    }
    if (c->q_tail == nullptr) { CU(cudaMalloc(&c->q_tail, 4096 * sizeof(unsigned))); CU(cudaMallocHost(&c->h_tail, 4096 * sizeof(unsigned))); }
    const int verbose = getenv("SOC_DOMAIN_VERBOSE") ? atoi(getenv("SOC_DOMAIN_VERBOSE")) : 0;
    const long long nlocal = A.nlocal;
    const int refill0 = A.refill, agg0 = A.agg_steps;
    // tile pass: lanes refilled when 24 of the warp are idle; lanes combined during a packet's first tp_agg steps (beyond the very first)
    int tp_refill = 24, tp_agg = 0;
    if (const char *e = getenv("SOC_TILEPASS_REFILL")) { int v = atoi(e); if (v >= 1 && v <= 32) tp_refill = v; }     // tuning knob
    if (const char *e = getenv("SOC_TILEPASS_AGG")) { int v = atoi(e); if (v >= 0) tp_agg = v; }                      // tuning knob
    const int t0[3] = { A.tile_x0, A.tile_y0, A.tile_z0 };
    A.dom = 1; A.q_base = c->queues; A.q_tail = c->q_tail; A.q_cap = chunk; A.q_in = c->queues; A.dom_base = 0;
    A.q_nparts = 1; A.q_part[0] = 0;
    c->domain_launches = 0; c->domain_parked = 0;
    for (long long u0 = 0; u0 < nlocal; u0 += chunk) {
        const long long n = nlocal - u0 < chunk ? nlocal - u0 : chunk;
        CU(cudaMemsetAsync(c->q_tail, 0, sizeof(unsigned), c->stream));
        CU(cudaMemsetAsync(A.work, 0, sizeof(unsigned long long), c->stream));
        // pass 1: the box is the tile
        A.dom_faces = 0;
        for (int k = 0; k < 3; k++) {
            A.dom_lo[k] = t0[k]; A.dom_hi[k] = t0[k] + SOC_TILE_N - 1;
            if (A.dom_lo[k] == 0) A.dom_faces |= 1 << (2 * k);
            if (A.dom_hi[k] == dim[k] - 1) A.dom_faces |= 2 << (2 * k);
        }
        A.deposit = DEP_TILE; A.unit0 = u0; A.nlocal = n;
        A.refill = tp_refill; A.agg_steps = tp_agg;
        long long needb = (n + threads - 1) / threads;
        cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
        if (verbose > 1) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventRecord(e0, c->stream); }
        launch_sim_tile_pass(A, (int)(needb < blocks ? (needb < 1 ? 1 : needb) : blocks), threads, c->stream);
        if (verbose > 1) cudaEventRecord(e1, c->stream);
        c->launches++; c->domain_launches++;
        CU(cudaMemcpyAsync(c->h_tail, c->q_tail, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        const long long n2 = c->h_tail[0];
        if (verbose == 1) fprintf(stderr, "soc_b200: two-pass launch: %lld packets emitted, %lld parked at the border of the tile\n", n, n2);
        if (n2 > 0) {
            // pass 2: the whole grid as one box, plain adds; nothing is parked
            A.dom_faces = 0x3f;
            for (int k = 0; k < 3; k++) { A.dom_lo[k] = 0; A.dom_hi[k] = dim[k] - 1; }
            A.deposit = DEP_RED; A.unit0 = 0; A.nlocal = n2; A.q_part[1] = n2;
            A.refill = refill0; A.agg_steps = agg0;
            CU(cudaMemsetAsync(A.work, 0, sizeof(unsigned long long), c->stream));
            needb = (n2 + threads - 1) / threads;
            launch_sim_domain(A, (int)(needb < blocks ? (needb < 1 ? 1 : needb) : blocks), threads, c->stream);
            c->launches++; c->domain_launches++; c->domain_parked += n2;
        }
        if (verbose > 1) {
            cudaEventRecord(e2, c->stream); cudaEventSynchronize(e2);
            float t1 = 0.0f, t2 = 0.0f; cudaEventElapsedTime(&t1, e0, e1); cudaEventElapsedTime(&t2, e1, e2);
            fprintf(stderr, "soc_b200: two-pass launch: %lld packets, tile pass %.3f ms, %lld parked, second pass %.3f ms\n", n, t1, n2, t2);
            cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
        }
        CU(cudaGetLastError());
    }
    A.deposit = DEP_TILE; A.nlocal = nlocal; A.dom = 0; A.unit0 = 0; A.refill = refill0; A.agg_steps = agg0;
    sim_note_two_pass();
    return SOC_OK;
}

static int sim_launch(soc_context *c, SimArgs &A, const char *who) {
    A.nlocal = (A.nunits - A.rank + A.world - 1) / A.world;
    const bool oct = A.G.levels > 1, dbl = A.G.dbl_sim != 0;
    int threads, blocks;
    if (c->rng_mode == SOC_RNG_REFERENCE) {
        threads = 128;
        long long local = (A.nunits - A.rank + A.world - 1) / A.world;
        blocks = (int)((local + threads - 1) / threads);
        if (blocks < 1) blocks = 1;
    } else {
        threads = 256;
        blocks = c->sms * sim_blocks_per_sm(c->rng_mode, oct, dbl, threads);
        long long local = (A.nunits - A.rank + A.world - 1) / A.world;
        long long needb = (local + threads - 1) / threads;
        if (needb < blocks) blocks = (int)(needb < 1 ? 1 : needb);
        CU(cudaMemsetAsync(A.work, 0, sizeof(unsigned long long), c->stream));
    }
    if (c->rng_mode != SOC_RNG_REFERENCE) {
        const size_t n = (size_t)A.G.cells * 4;
        if (c->acc == nullptr || c->acc_bytes != n) {
            if (c->acc) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->acc)); c->acc = nullptr; }
            CU(cudaMalloc(&c->acc, n + 16));
            c->acc_bytes = n;
            CU(cudaMemsetAsync(c->acc, 0, n + 16, c->stream));
        }
        A.acc = c->acc; A.use_acc = 1;
        A.dens_brick = c->layout ? c->dens_brick : nullptr;
        A.kappa = nullptr;
        if (sim_kappa_eligible(A, c->rng_mode) && !getenv("SOC_NO_KAPPA")) {
            if (c->kappa == nullptr || c->kappa_cells != (size_t)A.G.cells) {
                if (c->kappa) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->kappa)); c->kappa = nullptr; }
                CU(cudaMalloc(&c->kappa, (size_t)A.G.cells * sizeof(float2)));
                c->kappa_cells = (size_t)A.G.cells;
            }
            A.kappa = c->kappa;
        }
        A.brick = sim_uses_bricks(A, c->rng_mode) ? 1 : 0;
        A.nbr = c->nbr;
        A.pend = c->pend; A.ahead = c->ahead;
        // order of the work units: 0 = as numbered, 1 = background surface elements in 16 x 16 tiles (needs face edges that
        // are multiples of 16), 2 = scattered (experiment)
        A.scramble = 0ull;
        if (const char *e = getenv("SOC_UNIT_ORDER")) { const int v = atoi(e); A.scramble = v == 2 ? 1000003ull : (unsigned long long)(v == 1); }   // tuning knob
        if (A.scramble == 1ull && ((A.G.nx | A.G.ny | A.G.nz) & 15)) A.scramble = 0ull;
        A.slab_xy = A.G.nx * A.G.ny;
    }
    CU(cudaEventRecord(c->ev0, c->stream));
    // layout of the bricked arrays: one box (plain brick order), or the boxes of the domain mode (soc_set_domains)
    int edge = 0;
    A.dsplit[0] = A.dsplit[1] = A.dsplit[2] = 1;
    A.dsize[0] = A.G.nx; A.dsize[1] = A.G.ny; A.dsize[2] = A.G.nz;
    if (c->rng_mode != SOC_RNG_REFERENCE && c->domains >= 0 && A.brick && sim_domains_eligible(A, c->rng_mode))
        edge = c->domains > 0 ? c->domains : (A.G.nxyz > (1LL << 25) ? 256 : 0);
    bool domains = false;
    if (edge > 0 && (c->domains > 0 || A.G.nx > edge || A.G.ny > edge || A.G.nz > edge)) {       // forced: also with a single domain
        const int dim[3] = { A.G.nx, A.G.ny, A.G.nz };
        domains = true;
        for (int k = 0; k < 3; k++) {
            const int ns = (dim[k] + edge - 1) / edge;
            if (dim[k] % ns != 0 || ((dim[k] / ns) & 1)) domains = false;        // boxes of one size with even edges only
            A.dsplit[k] = ns; A.dsize[k] = dim[k] / ns;
        }
        if (!domains) { A.dsplit[0] = A.dsplit[1] = A.dsplit[2] = 1; A.dsize[0] = dim[0]; A.dsize[1] = dim[1]; A.dsize[2] = dim[2]; }
    }
    if (A.brick) {
        if (c->brick_ns[0] != A.dsplit[0] || c->brick_ns[1] != A.dsplit[1] || c->brick_ns[2] != A.dsplit[2]) {
            launch_brick_permute(A, c->dens_brick, c->stream);
            c->launches++;
            for (int k = 0; k < 3; k++) c->brick_ns[k] = A.dsplit[k];
        }
        A.brick_by = 4 * A.dsize[0] - 2; A.brick_bz = 2 * A.dsize[0] * A.dsize[1] - 4;
    }
    if (A.kappa != nullptr) { launch_kappa(A, c->stream); c->launches++; }     // OPT may have changed since the last launch
    if (domains) {
        int r = sim_launch_domains(c, A, blocks, threads);
        if (r != SOC_OK) return r;
    } else if (c->two_pass && sim_two_pass_eligible(A, c->rng_mode)) {
        int r = sim_launch_two_pass(c, A, blocks, threads);
        if (r != SOC_OK) return r;
    } else launch_sim(A, c->rng_mode, blocks, threads, c->stream);
    if (A.use_acc) { launch_fold_acc(A, c->stream); c->launches++; }
    CU(cudaEventRecord(c->ev1, c->stream));
    c->timed = true;
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SOC_ERR_CUDA, "%s: launch failed: %s", who, cudaGetErrorString(e));
    return SOC_OK;
}

int soc_sim_pb(soc_context *c, int source, int packets, int batch, float seed, float abs, float sca, float bg, float tw, int global) {
    NEED_CTX(c);
    (void)packets;
    if (source != 0 && source != 1 && source != 3) return fail(SOC_ERR_ARG, "soc_sim_pb: SOURCE=%d", source);
    if (source == 3 && !(c->have_params && (c->P.roi_flags & 1))) return fail(SOC_ERR_STATE, "soc_sim_pb: SOURCE=3 needs WITH_ROI_LOAD (roi_flags & 1)");
    SimArgs A;
    int r = sim_common(c, A, source == 0 ? SIM_PS : (source == 1 ? SIM_BG : SIM_ROI), batch, seed, abs, sca, tw, global, "soc_sim_pb");
    if (r != SOC_OK) return r;
    A.bg = bg;
    if (source == 3) {
        // PACKETS = number of surface elements, 100 work items each, BATCH a multiple of the Healpix pixel count (ASOC.py:1093-1110)
        if (packets < 1 || (size_t)packets != roi_elems(A.roi.dim)) return fail(SOC_ERR_ARG, "soc_sim_pb: SOURCE=3 takes PACKETS = number of ROI surface elements (%zu)", roi_elems(A.roi.dim));
        A.roi_nelem = packets;
        const long long items = 100LL * packets < global ? 100LL * packets : global;
        A.nunits = (c->rng_mode == SOC_RNG_REFERENCE) ? global : items * batch;
        if (A.deposit == DEP_TILE) A.deposit = DEP_RED;
        return sim_launch(c, A, "soc_sim_pb");
    }
    if (source == 0) {
        const size_t nps = (size_t)c->P.no_ps;
        if ((r = need(c, SOC_BUF_PSPOS, nps * 12, "soc_sim_pb")) != SOC_OK) return r;
        if ((r = need(c, SOC_BUF_PS, nps * 4, "soc_sim_pb")) != SOC_OK) return r;
        if (c->P.ps_method == 2) {
            if ((r = need(c, SOC_BUF_XPS_NSIDE, nps * 4, "soc_sim_pb")) != SOC_OK) return r;
        }
        if (c->P.ps_method == 2 || c->P.ps_method == 5) {
            if ((r = need(c, SOC_BUF_XPS_SIDE, nps * 12, "soc_sim_pb")) != SOC_OK) return r;
            if ((r = need(c, SOC_BUF_XPS_AREA, nps * 12, "soc_sim_pb")) != SOC_OK) return r;
        }
    }
    if (c->rng_mode == SOC_RNG_REFERENCE) A.nunits = global;
    else A.nunits = (source == 1 ? (long long)(8LL * c->G.area < global ? 8LL * c->G.area : global) : (long long)global) * batch;
    // shared-memory tile around the first point source (regular grids)
    if (A.deposit == DEP_TILE) {
        if (source == 0 && c->G.levels == 1 && c->P.no_ps == 1) {
            float p[3];
            CU(cudaMemcpyAsync(p, c->buf[SOC_BUF_PSPOS].ptr, 12, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            int o[3], dim[3] = { c->G.nx, c->G.ny, c->G.nz };
            for (int k = 0; k < 3; k++) {
                o[k] = (int)floorf(p[k]) - SOC_TILE_N / 2;
                if (o[k] > dim[k] - SOC_TILE_N) o[k] = dim[k] - SOC_TILE_N;
                if (o[k] < 0) o[k] = 0;
            }
            A.tile_x0 = o[0]; A.tile_y0 = o[1]; A.tile_z0 = o[2];
            A.tile_inside = p[0] >= 0.0f && p[0] < (float)dim[0] && p[1] >= 0.0f && p[1] < (float)dim[1] && p[2] >= 0.0f && p[2] < (float)dim[2];
            A.tile_lo = o[2] * c->G.nx * c->G.ny;
            A.tile_span = SOC_TILE_N * c->G.nx * c->G.ny;
            // the tile takes the adds within 8 cells of the source: lanes are combined only that long (measured: PS launch
            // of the bench step 66.3 ms with 24 steps, 63.9 ms with 6)
            if (A.agg_steps > 8) A.agg_steps = 8;
        } else A.deposit = (source == 0) ? DEP_WARP : DEP_RED;     // background packets do not share cells: plain adds
    }
    return sim_launch(c, A, "soc_sim_pb");
}

int soc_sim_hp(soc_context *c, int packets, int batch, float seed, float abs, float sca, float tw, int global) {
    NEED_CTX(c);
    (void)packets;
    SimArgs A;
    int r = sim_common(c, A, SIM_HP, batch, seed, abs, sca, tw, global, "soc_sim_hp");
    if (r != SOC_OK) return r;
    if ((r = need(c, SOC_BUF_HPBG, 49152 * 4, "soc_sim_hp")) != SOC_OK) return r;
    if (c->P.hpbg_weighted && (r = need(c, SOC_BUF_HPBGP, 49152 * 4, "soc_sim_hp")) != SOC_OK) return r;
    // kernel_ASOC.c:878: work items beyond 8*AREA return without simulating anything
    long long items = 8LL * c->G.area < global ? 8LL * c->G.area : global;
    A.nunits = (c->rng_mode == SOC_RNG_REFERENCE) ? global : items * batch;
    if (A.deposit == DEP_TILE) A.deposit = DEP_WARP;
    return sim_launch(c, A, "soc_sim_hp");
}

int soc_sim_cl(soc_context *c, int source, int packets, int batch, float seed, float abs, float sca, float tw, int global) {
    NEED_CTX(c);
    (void)packets; (void)source;
    SimArgs A;
    int r = sim_common(c, A, SIM_CL, batch, seed, abs, sca, tw, global, "soc_sim_cl");
    if (r != SOC_OK) return r;
    const size_t n = (size_t)c->G.cells * 4;
    if ((r = need(c, SOC_BUF_EMIT, n, "soc_sim_cl")) != SOC_OK) return r;
    if (c->P.use_emweight && (r = need(c, SOC_BUF_EMWEI, n, "soc_sim_cl")) != SOC_OK) return r;
    if (c->P.use_emweight > 1) return fail(SOC_ERR_UNSUPPORTED, "soc_sim_cl: USE_EMWEIGHT=2 is not implemented");
    A.nunits = (c->rng_mode == SOC_RNG_REFERENCE) ? global : c->G.cells;
    if (A.deposit == DEP_TILE) A.deposit = DEP_WARP;
    return sim_launch(c, A, "soc_sim_cl");
}

int soc_eq_temperature(soc_context *c, int level, float adhoc, float kE, float Emin, int NE) {
    NEED_CTX(c);
    if (!c->have_grid || !c->have_params) return fail(SOC_ERR_STATE, "soc_eq_temperature: grid and params first");
    if (level < 0 || level >= c->G.levels || NE < 2) return fail(SOC_ERR_ARG, "soc_eq_temperature: level %d NE %d", level, NE);
    const size_t n = (size_t)c->G.cells * 4;
    int r;
    if ((r = need(c, SOC_BUF_TTT, (size_t)NE * 4, "soc_eq_temperature")) != SOC_OK) return r;
    if ((r = need(c, SOC_BUF_EMIT, n, "soc_eq_temperature")) != SOC_OK) return r;
    if ((r = ensure_zeroed(c, SOC_BUF_TNEW, n)) != SOC_OK) return r;
    launch_eq_temperature(c->G, level, adhoc, kE, Emin, NE, c->P.factor, c->P.length, dptr<float>(c, SOC_BUF_TTT),
                          dptr<float>(c, SOC_BUF_EMIT), dptr<float>(c, SOC_BUF_TNEW), c->stream);
    c->launches++;
    CU(cudaGetLastError());
    return SOC_OK;
}

int soc_build_opt(soc_context *c, int ndust, const float *kabs, const float *ksca, int first, int single_abu) {
    NEED_CTX(c);
    if (!c->have_grid || !c->have_params) return fail(SOC_ERR_STATE, "soc_build_opt: grid and params first");
    if (!c->P.with_abu) return fail(SOC_ERR_STATE, "soc_build_opt: the run has no variable abundances (WITH_ABU = 0)");
    if (ndust < 1 || ndust > SOC_MAX_DUSTS || !kabs || !ksca || first < 0 || first >= ndust) return fail(SOC_ERR_ARG, "soc_build_opt: ndust=%d first=%d", ndust, first);
    if (single_abu && ndust != 2) return fail(SOC_ERR_ARG, "soc_build_opt: SINGLE_ABU assumes exactly two dust species (ASOC.py:620)");
    const size_t cells = (size_t)c->G.cells;
    int r;
    // SINGLE_ABU reads one abundance per cell, ABU[cells, 1]; else ABU[cells, ndust]
    const int width = single_abu ? (int)(c->buf[SOC_BUF_ABU].bytes / (cells * 4)) : ndust;
    if (width < 1 || (!single_abu && width != ndust)) return fail(SOC_ERR_STATE, "soc_build_opt: ABU holds %zu bytes, expected CELLS x %d floats", c->buf[SOC_BUF_ABU].bytes, ndust);
    if ((r = need(c, SOC_BUF_ABU, cells * 4 * (size_t)width, "soc_build_opt")) != SOC_OK) return r;
    if ((r = ensure(c, SOC_BUF_OPT, cells * 8)) != SOC_OK) return r;
    launch_build_opt(dptr<float>(c, SOC_BUF_ABU), dptr<float>(c, SOC_BUF_OPT), (long long)cells, width, single_abu ? 0 : first, single_abu,
                     c->P.opt_is_half, kabs, ksca, c->stream);
    c->launches++;
    CU(cudaGetLastError());
    return SOC_OK;
}

int soc_absorbed_begin(soc_context *c, int nfreq) {
    NEED_CTX(c);
    if (!c->have_grid || !c->have_params) return fail(SOC_ERR_STATE, "soc_absorbed_begin: grid and params first");
    if (nfreq < 1) return fail(SOC_ERR_ARG, "soc_absorbed_begin: nfreq=%d", nfreq);
    const size_t bytes = (size_t)c->G.cells * nfreq * 4;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    if (c->buf[SOC_BUF_FABS].bytes != bytes && bytes > free_b + c->buf[SOC_BUF_FABS].bytes)
        return fail(SOC_ERR_STATE, "soc_absorbed_begin: %zu bytes for the [CELLS,NFREQ] array do not fit (%zu free)", bytes, free_b);
    int r = soc_clear(c, SOC_BUF_FABS, bytes);
    if (r != SOC_OK) return r;
    c->fabs_nfreq = nfreq;
    return SOC_OK;
}

int soc_absorbed_add(soc_context *c, int ifreq) {
    NEED_CTX(c);
    if (c->fabs_nfreq < 1 || c->buf[SOC_BUF_FABS].ptr == nullptr) return fail(SOC_ERR_STATE, "soc_absorbed_add: soc_absorbed_begin first");
    if (ifreq < 0 || ifreq >= c->fabs_nfreq) return fail(SOC_ERR_ARG, "soc_absorbed_add: ifreq %d of %d", ifreq, c->fabs_nfreq);
    int r = need(c, SOC_BUF_INT, (size_t)c->G.cells * 4, "soc_absorbed_add");
    if (r != SOC_OK) return r;
    launch_absorbed_add(dptr<float>(c, SOC_BUF_FABS), dptr<float>(c, SOC_BUF_INT), c->G.cells, c->fabs_nfreq, ifreq, c->stream);
    c->launches++;
    CU(cudaGetLastError());
    return SOC_OK;
}

int soc_absorbed_finish(soc_context *c, float coeff0, float nnnlimit, int finish_scale, float *host) {
    NEED_CTX(c);
    if (c->fabs_nfreq < 1 || c->buf[SOC_BUF_FABS].ptr == nullptr) return fail(SOC_ERR_STATE, "soc_absorbed_finish: soc_absorbed_begin first");
    if (finish_scale) {
        launch_absorbed_scale(c->G, dptr<float>(c, SOC_BUF_FABS), c->fabs_nfreq, coeff0, nnnlimit, c->stream);
        c->launches++;
        CU(cudaGetLastError());
    }
    if (host != nullptr) {
        CU(cudaMemcpyAsync(host, c->buf[SOC_BUF_FABS].ptr, c->buf[SOC_BUF_FABS].bytes, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return SOC_OK;
}

int soc_split_absorbed(soc_context *c, int idust, int ndust, int nfreq, const double *rabs, float *host) {
    NEED_CTX(c);
    if (!c->have_grid) return fail(SOC_ERR_STATE, "soc_split_absorbed: grid first");
    if (ndust < 1 || idust < 0 || idust >= ndust || nfreq < 1 || rabs == nullptr || host == nullptr) return fail(SOC_ERR_ARG, "soc_split_absorbed: bad arguments");
    const size_t cells = (size_t)c->G.cells, nb = cells * (size_t)nfreq * 4;
    int r;
    if ((r = need(c, SOC_BUF_FABS, nb, "soc_split_absorbed")) != SOC_OK) return r;
    if ((r = need(c, SOC_BUF_ABU, cells * (size_t)ndust * 4, "soc_split_absorbed")) != SOC_OK) return r;
    const size_t rb = (size_t)nfreq * ndust * sizeof(double);
    if (c->scratch_bytes < nb + rb) {
        if (c->scratch) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->scratch)); c->scratch = nullptr; c->scratch_bytes = 0; }
        CU(cudaMalloc(&c->scratch, nb + rb));
        c->scratch_bytes = nb + rb;
    }
    double *d_rabs = reinterpret_cast<double *>(c->scratch);                     // 8-byte aligned: first
    float *d_out = reinterpret_cast<float *>(reinterpret_cast<char *>(c->scratch) + rb);
    CU(cudaMemcpyAsync(d_rabs, rabs, rb, cudaMemcpyHostToDevice, c->stream));
    launch_split_absorbed(idust, (long long)cells, nfreq, ndust, d_rabs, dptr<float>(c, SOC_BUF_ABU), dptr<float>(c, SOC_BUF_FABS), d_out, c->stream);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host, d_out, nb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SOC_OK;
}

int soc_emission(soc_context *c, float freq, float fabs_) {
    NEED_CTX(c);
    if (!c->have_grid || !c->have_params) return fail(SOC_ERR_STATE, "soc_emission: grid and params first");
    const size_t n = (size_t)c->G.cells * 4;
    int r;
    if ((r = need(c, SOC_BUF_TNEW, n, "soc_emission")) != SOC_OK) return r;
    if ((r = ensure(c, SOC_BUF_EMIT, n)) != SOC_OK) return r;
    launch_emission(c->G.cells, freq, fabs_, c->P.factor, c->P.length, dptr<float>(c, SOC_BUF_TNEW), dptr<float>(c, SOC_BUF_EMIT), c->stream);
    c->launches++;
    CU(cudaGetLastError());
    return SOC_OK;
}

int soc_emission2(soc_context *c, int c0, int c1, int nfreq, const float *freq, const float *fabs_, float *emit_out) {
    NEED_CTX(c);
    if (!c->have_grid || !c->have_params) return fail(SOC_ERR_STATE, "soc_emission2: grid and params first");
    if (c0 < 0 || c1 <= c0 || c1 > c->G.cells || nfreq < 1 || !freq || !fabs_ || !emit_out) return fail(SOC_ERR_ARG, "soc_emission2: bad arguments");
    int r;
    if ((r = need(c, SOC_BUF_TNEW, (size_t)c->G.cells * 4, "soc_emission2")) != SOC_OK) return r;
    const size_t nb = (size_t)(c1 - c0) * nfreq * 4;
    if (c->scratch_bytes < nb + 2 * (size_t)nfreq * 4) {
        if (c->scratch) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->scratch)); c->scratch = nullptr; c->scratch_bytes = 0; }
        CU(cudaMalloc(&c->scratch, nb + 2 * (size_t)nfreq * 4));
        c->scratch_bytes = nb + 2 * (size_t)nfreq * 4;
    }
    float *d_emit = reinterpret_cast<float *>(c->scratch);
    float *d_freq = d_emit + (size_t)(c1 - c0) * nfreq, *d_fabs = d_freq + nfreq;
    CU(cudaMemcpyAsync(d_freq, freq, (size_t)nfreq * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_fabs, fabs_, (size_t)nfreq * 4, cudaMemcpyHostToDevice, c->stream));
    launch_emission2(c0, c1, nfreq, c->P.factor, c->P.length, d_freq, d_fabs, dptr<float>(c, SOC_BUF_TNEW), d_emit, c->stream);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(emit_out, d_emit, nb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SOC_OK;
}

static int map_common(soc_context *c, MapArgs &M, size_t npixels, float abs, float sca, int save_colden, const char *who) {
    if (!c->have_grid || !c->have_params) return fail(SOC_ERR_STATE, "%s: grid and params first", who);
    const size_t n = (size_t)c->G.cells * 4;
    int r;
    if ((r = need(c, SOC_BUF_EMIT, n, who)) != SOC_OK) return r;
    if (c->P.with_abu && (r = need(c, SOC_BUF_OPT, 2 * n, who)) != SOC_OK) return r;
    if ((r = ensure(c, SOC_BUF_MAP, npixels * 4)) != SOC_OK) return r;
    if ((r = ensure(c, SOC_BUF_SAVETAU, npixels * 4)) != SOC_OK) return r;
    memset(&M, 0, sizeof(M));
    M.G = c->G;
    M.map = dptr<float>(c, SOC_BUF_MAP); M.savetau = dptr<float>(c, SOC_BUF_SAVETAU);
    M.emit = dptr<float>(c, SOC_BUF_EMIT); M.opt = dptr<float>(c, SOC_BUF_OPT);
    M.kabs = abs; M.ksca = sca; M.length = c->P.length;
    M.with_abu = c->P.with_abu; M.level_threshold = c->P.level_threshold; M.save_colden = save_colden;
    M.map_interpolation = c->P.map_interpolation;
    M.maph_literal = (c->P.ref_quirks & 2) ? 1 : 0;
    if ((r = roi_args(c, M.roi, false, who)) != SOC_OK) return r;
    M.counters = c->counters;
    return SOC_OK;
}

int soc_mapping(soc_context *c, float map_dx, int npix_x, int npix_y, const float dir[3], const float ra[3],
                const float de[3], float abs, float sca, const float centre[3], const float intobs[3], int save_colden) {
    NEED_CTX(c);
    if (npix_x < 1 || npix_y < 1 || !dir || !ra || !de || !centre || !intobs) return fail(SOC_ERR_ARG, "soc_mapping: bad arguments");
    MapArgs M;
    int r = map_common(c, M, (size_t)npix_x * npix_y, abs, sca, save_colden, "soc_mapping");
    if (r != SOC_OK) return r;
    M.dir = { dir[0], dir[1], dir[2] }; M.ra = { ra[0], ra[1], ra[2] }; M.de = { de[0], de[1], de[2] };
    M.centre = { centre[0], centre[1], centre[2] }; M.intobs = { intobs[0], intobs[1], intobs[2] };
    M.map_dx = map_dx; M.npx = npix_x; M.npy = npix_y;
    CU(cudaEventRecord(c->ev0, c->stream));
    launch_mapping(M, false, c->stream);
    CU(cudaEventRecord(c->ev1, c->stream));
    c->timed = true; c->launches++;
    CU(cudaGetLastError());
    return SOC_OK;
}

int soc_mapping_levels(soc_context *c, float map_dx, int npix_x, int npix_y, const float dir[3], const float ra[3],
                       const float de[3], float abs, float sca, const float centre[3], const float intobs[3], int save_colden) {
    NEED_CTX(c);
    if (npix_x < 1 || npix_y < 1 || !dir || !ra || !de || !centre || !intobs) return fail(SOC_ERR_ARG, "soc_mapping_levels: bad arguments");
    if (c->have_grid && c->G.levels > 12) return fail(SOC_ERR_UNSUPPORTED, "soc_mapping_levels: at most 12 hierarchy levels (%d)", c->G.levels);
    MapArgs M;
    int r = map_common(c, M, (size_t)npix_x * npix_y * (size_t)(c->have_grid ? c->G.levels : 1), abs, sca, save_colden, "soc_mapping_levels");
    if (r != SOC_OK) return r;
    if (!save_colden) M.savetau = nullptr;
    M.dir = { dir[0], dir[1], dir[2] }; M.ra = { ra[0], ra[1], ra[2] }; M.de = { de[0], de[1], de[2] };
    M.centre = { centre[0], centre[1], centre[2] }; M.intobs = { intobs[0], intobs[1], intobs[2] };
    M.map_dx = map_dx; M.npx = npix_x; M.npy = npix_y;
    CU(cudaEventRecord(c->ev0, c->stream));
    launch_mapping_levels(M, c->stream);
    CU(cudaEventRecord(c->ev1, c->stream));
    c->timed = true; c->launches++;
    CU(cudaGetLastError());
    return SOC_OK;
}

int soc_healpix_mapping(soc_context *c, int nside, float abs, float sca, const float intobs[3], int save_colden) {
    NEED_CTX(c);
    if (nside < 1 || nside > 8192 || !intobs) return fail(SOC_ERR_ARG, "soc_healpix_mapping: bad arguments");
    MapArgs M;
    int r = map_common(c, M, (size_t)12 * nside * nside, abs, sca, save_colden, "soc_healpix_mapping");
    if (r != SOC_OK) return r;
    M.intobs = { intobs[0], intobs[1], intobs[2] };
    M.nside = nside;
    CU(cudaEventRecord(c->ev0, c->stream));
    launch_mapping(M, true, c->stream);
    CU(cudaEventRecord(c->ev1, c->stream));
    c->timed = true; c->launches++;
    CU(cudaGetLastError());
    return SOC_OK;
}

int soc_ps_tau(soc_context *c, int no, const float dir[3], float abs, float sca, float *colden_out, float *tau_out) {
    NEED_CTX(c);
    if (no < 1 || !dir || !colden_out || !tau_out) return fail(SOC_ERR_ARG, "soc_ps_tau: bad arguments");
    if (!c->have_grid || !c->have_params) return fail(SOC_ERR_STATE, "soc_ps_tau: grid and params first");
    int r;
    if ((r = need(c, SOC_BUF_PSPOS, (size_t)no * 12, "soc_ps_tau")) != SOC_OK) return r;
    if (c->P.with_abu && (r = need(c, SOC_BUF_OPT, (size_t)c->G.cells * 8, "soc_ps_tau")) != SOC_OK) return r;
    if ((r = ensure(c, SOC_BUF_MAP, (size_t)no * 4)) != SOC_OK) return r;
    if ((r = ensure(c, SOC_BUF_SAVETAU, (size_t)no * 4)) != SOC_OK) return r;
    MapArgs M;
    memset(&M, 0, sizeof(M));
    M.G = c->G; M.opt = dptr<float>(c, SOC_BUF_OPT); M.kabs = abs; M.ksca = sca; M.length = c->P.length; M.with_abu = c->P.with_abu;
    M.dir = { dir[0], dir[1], dir[2] };
    launch_pstau(M, no, dptr<float>(c, SOC_BUF_PSPOS), dptr<float>(c, SOC_BUF_MAP), dptr<float>(c, SOC_BUF_SAVETAU), c->stream);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(colden_out, c->buf[SOC_BUF_MAP].ptr, (size_t)no * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(tau_out, c->buf[SOC_BUF_SAVETAU].ptr, (size_t)no * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SOC_OK;
}

// ---- scattered light ------------------------------------------------------------------------------------------
int soc_sca_zero_out(soc_context *c, int ndir, int npix_x, int npix_y) {
    NEED_CTX(c);
    if (ndir < 0) return soc_clear(c, SOC_BUF_OUT, (size_t)12 * ndir * ndir * 4);      // Healpix image, NSIDE = -ndir
    if (ndir < 1 || npix_x < 1 || npix_y < 1) return fail(SOC_ERR_ARG, "soc_sca_zero_out: bad arguments");
    return soc_clear(c, SOC_BUF_OUT, (size_t)ndir * npix_x * npix_y * 4);
}

static int sca_common(soc_context *c, ScaArgs &S, int kind, int flavour, int batch, float seed, float abs, float sca, int ndir,
                      int npix_x, int npix_y, float map_dx, const float centre[3], int global, const char *who) {
    if (!c->have_grid || !c->have_params) return fail(SOC_ERR_STATE, "%s: grid and params first", who);
    if (batch < 1 || global < 1 || ndir == 0) return fail(SOC_ERR_ARG, "%s: bad arguments", who);
    const int nside = ndir < 0 ? -ndir : 0;                    // Healpix image seen from ODIR[0..2] (ASOCS.py:44-47)
    if (nside > 8192) return fail(SOC_ERR_ARG, "%s: NSIDE=%d", who, nside);
    if (nside == 0 && (npix_x < 1 || npix_y < 1 || centre == nullptr)) return fail(SOC_ERR_ARG, "%s: bad map geometry", who);
    const soc_params &P = c->P;
    if (kind == SIM_PS && P.ps_method != 0 && P.ps_method != 1)
        return fail(SOC_ERR_UNSUPPORTED, "%s: PS_METHOD %d reads XPS_* through mistyped pointers in the reference (kernel_ASOC_sca.c:1486)", who, P.ps_method);
    const int nd_eff = nside ? 1 : ndir;
    const size_t n = (size_t)c->G.cells * 4, nd = (size_t)nd_eff * 12;
    const size_t npix = nside ? (size_t)12 * nside * nside : (size_t)ndir * npix_x * npix_y;
    const int nsf = P.with_msf ? P.ndust : 1;
    int r;
    if ((r = need(c, SOC_BUF_CSC, (size_t)P.bins * nsf * 4, who)) != SOC_OK) return r;
    if ((r = need(c, SOC_BUF_DSC, (size_t)P.bins * nsf * 4, who)) != SOC_OK) return r;
    if ((r = need(c, SOC_BUF_ODIR, nd, who)) != SOC_OK) return r;
    if (!nside) {
        if ((r = need(c, SOC_BUF_ORA, nd, who)) != SOC_OK) return r;
        if ((r = need(c, SOC_BUF_ODE, nd, who)) != SOC_OK) return r;
    }
    if (P.with_abu && (r = need(c, SOC_BUF_OPT, 2 * n, who)) != SOC_OK) return r;
    if (P.with_msf) {
        if ((r = need(c, SOC_BUF_ABU, n * P.ndust, who)) != SOC_OK) return r;
        if ((r = need(c, SOC_BUF_SCAV, (size_t)P.ndust * 4, who)) != SOC_OK) return r;
    }
    if ((r = ensure_zeroed(c, SOC_BUF_OUT, npix * 4)) != SOC_OK) return r;
    memset(&S, 0, sizeof(S));
    S.G = c->G;
    S.out = dptr<float>(c, SOC_BUF_OUT);
    S.opt = dptr<float>(c, SOC_BUF_OPT); S.dsc = dptr<float>(c, SOC_BUF_DSC); S.csc = dptr<float>(c, SOC_BUF_CSC);
    S.pspos = dptr<float>(c, SOC_BUF_PSPOS); S.ps = dptr<float>(c, SOC_BUF_PS);
    S.odir = dptr<float>(c, SOC_BUF_ODIR); S.ora = dptr<float>(c, SOC_BUF_ORA); S.ode = dptr<float>(c, SOC_BUF_ODE);
    S.hpbg = dptr<float>(c, SOC_BUF_HPBG); S.hpbgp = dptr<float>(c, SOC_BUF_HPBGP);
    S.emit = dptr<float>(c, SOC_BUF_EMIT); S.emwei = dptr<float>(c, SOC_BUF_EMWEI);
    S.abu = dptr<float>(c, SOC_BUF_ABU); S.scav = dptr<float>(c, SOC_BUF_SCAV);
    S.kabs = abs; S.ksca = sca; S.map_dx = map_dx;
    if (centre) S.centre = { centre[0], centre[1], centre[2] };
    S.kind = kind; S.flavour = flavour; S.batch = batch; S.global = global; S.ndir = nd_eff; S.nside = nside;
    S.npx = nside ? 1 : npix_x; S.npy = nside ? 1 : npix_y;
    S.bins = P.bins; S.no_ps = P.no_ps; S.ps_method = P.ps_method; S.with_abu = P.with_abu; S.ffs = P.ffs;
    S.hpbg_weighted = P.hpbg_weighted; S.use_emweight = P.use_emweight; S.with_ali = 0;
    S.with_msf = P.with_msf; S.ndust = P.with_msf ? P.ndust : 1; S.mirror = P.mirror;
    S.hg_test = (P.ref_quirks & 1) ? 1 : 0;
    S.nbr = c->nbr;
    S.rank = c->rank; S.world = c->world; S.ref_geometry = c->geometry; S.ev_batch = c->sc_batch > 0 ? c->sc_batch : 3; S.nav_hops = c->nav_hops;
    long long ms = 100LL * ((long long)c->G.nx + c->G.ny + c->G.nz) << (c->G.levels - 1);
    S.max_steps = (int)(ms > INT_MAX ? INT_MAX : ms);
    S.counters = c->counters; S.work = c->counters + 5;
    S.mwc.base_offset = seed_to_base(seed);
    S.mwc.base_state = mwc_mulmod(MWC_BASEID, mwc_powmod(MWC_A, S.mwc.base_offset));
    for (int k = 0; k < 26; k++) S.mwc.pow2k[k] = c->pow2k[k];
    uint32_t sb; memcpy(&sb, &seed, 4);
    S.phx.k0 = sb; S.phx.k1 = 0x534F4353u; S.phx.tag = (uint32_t)kind; S.phx.pad = 0;
    return SOC_OK;
}

static int sca_launch(soc_context *c, ScaArgs &S, const char *who) {
    const int threads = (c->rng_mode != SOC_RNG_REFERENCE && !c->geometry) ? 256 : 128;
    long long local = (S.nunits - S.rank + S.world - 1) / S.world;
    int blocks;
    if (c->rng_mode == SOC_RNG_REFERENCE) blocks = (int)((local + threads - 1) / threads);
    else {
        blocks = c->sms * sca_blocks_per_sm(c->G.levels > 1, c->G.dbl_sim != 0, threads);
        long long needb = (local + threads - 1) / threads;
        if (needb < blocks) blocks = (int)needb;
        CU(cudaMemsetAsync(S.work, 0, sizeof(unsigned long long), c->stream));
    }
    if (blocks < 1) blocks = 1;
    CU(cudaEventRecord(c->ev0, c->stream));
    launch_sca(S, c->rng_mode, blocks, threads, c->stream);
    CU(cudaEventRecord(c->ev1, c->stream));
    c->timed = true; c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SOC_ERR_CUDA, "%s: launch failed: %s", who, cudaGetErrorString(e));
    return SOC_OK;
}

int soc_sca_ps(soc_context *c, int packets, int batch, float seed, float abs, float sca, int ndir, int npix_x, int npix_y,
               float map_dx, const float centre[3], int global) {
    NEED_CTX(c);
    (void)packets;
    ScaArgs S;
    int r = sca_common(c, S, SIM_PS, 0, batch, seed, abs, sca, ndir, npix_x, npix_y, map_dx, centre, global, "soc_sca_ps");
    if (r != SOC_OK) return r;
    const size_t nps = (size_t)c->P.no_ps;
    if ((r = need(c, SOC_BUF_PSPOS, nps * 12, "soc_sca_ps")) != SOC_OK) return r;
    if ((r = need(c, SOC_BUF_PS, nps * 4, "soc_sca_ps")) != SOC_OK) return r;
    S.nunits = (c->rng_mode == SOC_RNG_REFERENCE) ? global : (long long)global * batch;
    return sca_launch(c, S, "soc_sca_ps");
}

int soc_sca_pb(soc_context *c, int source, int packets, int batch, float seed, float abs, float sca, float bg, int ndir,
               int npix_x, int npix_y, float map_dx, const float centre[3], int global) {
    NEED_CTX(c);
    (void)packets;
    if (source != 0 && source != 1 && source != 3) return fail(SOC_ERR_ARG, "soc_sca_pb: SOURCE=%d", source);
    if (source == 3 && !(c->have_params && (c->P.roi_flags & 1))) return fail(SOC_ERR_STATE, "soc_sca_pb: SOURCE=3 needs WITH_ROI_LOAD (roi_flags & 1)");
    ScaArgs S;
    int r = sca_common(c, S, source == 0 ? SIM_PS : (source == 1 ? SIM_BG : SIM_ROI), 1, batch, seed, abs, sca, ndir, npix_x, npix_y, map_dx, centre, global, "soc_sca_pb");
    if (r != SOC_OK) return r;
    if (source == 0) {
        const size_t nps = (size_t)c->P.no_ps;
        if ((r = need(c, SOC_BUF_PSPOS, nps * 12, "soc_sca_pb")) != SOC_OK) return r;
        if ((r = need(c, SOC_BUF_PS, nps * 4, "soc_sca_pb")) != SOC_OK) return r;
    }
    S.bg = bg;
    if (source == 3) {
        if ((r = roi_args(c, S.roi, true, "soc_sca_pb")) != SOC_OK) return r;
        if (packets < 1 || (size_t)packets != roi_elems(S.roi.dim)) return fail(SOC_ERR_ARG, "soc_sca_pb: SOURCE=3 takes PACKETS = number of ROI surface elements");
        S.roi_nelem = packets;
    }
    long long items = (source == 1 && 8LL * c->G.area < global) ? 8LL * c->G.area : ((source == 3 && 100LL * packets < global) ? 100LL * packets : global);
    S.nunits = (c->rng_mode == SOC_RNG_REFERENCE) ? global : items * batch;
    return sca_launch(c, S, "soc_sca_pb");
}

int soc_sca_hp(soc_context *c, int packets, int batch, float seed, float abs, float sca, int ndir, int npix_x, int npix_y,
               float map_dx, const float centre[3], int global) {
    NEED_CTX(c);
    (void)packets;
    ScaArgs S;
    int r = sca_common(c, S, SIM_HP, 2, batch, seed, abs, sca, ndir, npix_x, npix_y, map_dx, centre, global, "soc_sca_hp");
    if (r != SOC_OK) return r;
    if ((r = need(c, SOC_BUF_HPBG, 49152 * 4, "soc_sca_hp")) != SOC_OK) return r;
    if (c->P.hpbg_weighted && (r = need(c, SOC_BUF_HPBGP, 49152 * 4, "soc_sca_hp")) != SOC_OK) return r;
    S.nunits = (c->rng_mode == SOC_RNG_REFERENCE) ? global : (long long)global * batch;
    return sca_launch(c, S, "soc_sca_hp");
}

int soc_sca_cl(soc_context *c, int source, int packets, int batch, float seed, float abs, float sca, int ndir, int npix_x,
               int npix_y, float map_dx, const float centre[3], int global) {
    NEED_CTX(c);
    (void)packets; (void)source;
    ScaArgs S;
    int r = sca_common(c, S, SIM_CL, 3, batch, seed, abs, sca, ndir, npix_x, npix_y, map_dx, centre, global, "soc_sca_cl");
    if (r != SOC_OK) return r;
    const size_t n = (size_t)c->G.cells * 4;
    if ((r = need(c, SOC_BUF_EMIT, n, "soc_sca_cl")) != SOC_OK) return r;
    if (c->P.use_emweight && (r = need(c, SOC_BUF_EMWEI, n, "soc_sca_cl")) != SOC_OK) return r;
    if (c->P.use_emweight > 1) return fail(SOC_ERR_UNSUPPORTED, "soc_sca_cl: USE_EMWEIGHT=2 is not implemented");
    S.nunits = (c->rng_mode == SOC_RNG_REFERENCE) ? global : c->G.cells;
    return sca_launch(c, S, "soc_sca_cl");
}

int soc_get_counters(soc_context *c, soc_counters *out) {
    NEED_CTX(c);
    if (out == nullptr) return fail(SOC_ERR_ARG, "soc_get_counters: null");
    unsigned long long h[8];
    CU(cudaMemcpyAsync(h, c->counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memset(out, 0, sizeof(*out));
    out->packets = h[0]; out->steps = h[1]; out->scatterings = h[2]; out->peels = h[4];
    out->launches = c->launches; out->reserved[0] = h[3];       // reserved[0] = packets killed by the step guard
    return SOC_OK;
}

int soc_reset_counters(soc_context *c) {
    NEED_CTX(c);
    CU(cudaMemsetAsync(c->counters, 0, 8 * sizeof(unsigned long long), c->stream));
    c->launches = 0;
    return SOC_OK;
}

const char *soc_last_kernel(soc_context *c) { (void)c; return sim_last_kernel(); }

int soc_last_launch_ms(soc_context *c, float *ms) {
    NEED_CTX(c);
    if (ms == nullptr) return fail(SOC_ERR_ARG, "soc_last_launch_ms: null");
    if (!c->timed) return fail(SOC_ERR_STATE, "soc_last_launch_ms: nothing launched yet");
    CU(cudaEventSynchronize(c->ev1));
    CU(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return SOC_OK;
}

}  // extern "C"
