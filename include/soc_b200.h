/* soc_b200 -- C ABI of the B200 (sm_100a) implementation of SOC's photon-packet hot path.
 *
 * This is the drop-in boundary: every entry point replaces one use of pyopencl by the reference
 * drivers ASOC.py / ASOCS.py (context + queue, cl.Buffer + enqueue_copy, Program.build options,
 * and the kernel launches).  Plain pointers and sizes only; all functions return 0 on success and
 * a negative soc_status otherwise, with a human-readable message in soc_last_error().
 * Calls are made from one host thread per context (like the reference's single in-order queue);
 * launches are asynchronous on the context's stream, downloads and soc_sync() synchronise.
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   soc_create / soc_destroy / soc_sync     cl.Context, cl.CommandQueue, queue.finish()
 *                                           (ASOC_aux.py:1188-1256, ASOC.py:336, 1461)
 *   soc_set_params                          the -D macro list of Program.build (ASOC.py:344-396,
 *                                           ASOCS.py:133-152)
 *   soc_set_grid                            LCELLS/OFF/DENS/PAR buffers + kernel Parents
 *                                           (ASOC.py:428-454, 524-527, 580-586; kernel_ASOC_aux.c:688)
 *   soc_upload / soc_download / soc_clear   cl.Buffer + cl.enqueue_copy (ASOC.py:1174-1236, 1484, 1533)
 *   soc_zero_amc                            kernel ZeroAMC (kernel_ASOC_aux.c:657; ASOC.py:1115, 1183)
 *   soc_sim_pb / soc_sim_hp / soc_sim_cl    kernels SimRAM_PB / SimRAM_HP / SimRAM_CL
 *                                           (kernel_ASOC.c:15, 831, 1223; ASOC.py:1317-1419, 1847)
 *   soc_build_opt                           host loop OPT = sum_d ABU*K_d + upload (ASOC.py:1146-1175)
 *   soc_absorbed_begin / _add / _finish     FABSORBED[:,f] += TMP and the final scaling loop
 *   soc_split_absorbed                      kernel split_absorbed of kernel_A2E_MABU_aux.c (A2E_MABU.py:700-705)
 *                                           (ASOC.py:1482-1497, 2782-2878), kept on the device
 *   soc_eq_temperature / soc_emission       kernels EqTemperature / Emission / Emission2
 *     / soc_emission2                       (kernel_ASOC_aux.c:745, 793, 862; ASOC.py:2027-2040, 2154-2197)
 *   soc_ps_tau                              kernel PSTau (kernel_ASOC_map.c:1545; ASOC.py:3576-3644)
 *   soc_mapping / soc_healpix_mapping       kernels Mapping / HealpixMapping
 *                                           (kernel_ASOC_map.c:496, 890; ASOC.py:3127-3139)
 *   soc_mapping_levels                      kernel Mapping of kernel_ASOC_map_H.c:380 (ASOC.py:3320-3440)
 *   soc_sca_zero_out / soc_sca_ps / _pb     kernels zero_out / SimRAM_PS / SimRAM_PB / SimRAM_HP / SimRAM_CL of
 *     / soc_sca_hp / soc_sca_cl             kernel_ASOC_sca.c:14, 1462, 471, 40, 1098 (ASOCS.py:515, 665-708)
 * New (no counterpart in the single-device reference):
 *   soc_set_shard, soc_device_ptr           packet sharding over ranks and the device addresses a
 *                                           host-side NCCL all-reduce needs
 *   soc_set_rng_mode, soc_set_tuning, soc_set_geometry, soc_set_layout, soc_set_domains,
 *   soc_get_counters, soc_last_launch_ms, soc_last_kernel,
 *   soc_stream                              stream layout, accumulation engine, work counters, device timing
 */
#ifndef SOC_B200_H
#define SOC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct soc_context soc_context;

enum soc_status {
    SOC_OK = 0,
    SOC_ERR_CUDA = -1,         /* CUDA runtime error (message has the CUDA error string)      */
    SOC_ERR_ARG = -2,          /* bad argument / unknown buffer / size mismatch               */
    SOC_ERR_STATE = -3,        /* call order: grid or params or a required buffer is missing  */
    SOC_ERR_UNSUPPORTED = -4   /* option of the reference that this library does not implement */
};

/* The compile-time options of the reference kernels (ASOC.py:344-362) as run-time values.
 * Options that cannot work in the reference as shipped (DIR_WEIGHT, PS_METHOD 3) or that are out of
 * scope (DO_SPLIT, POLSTAT) are rejected with SOC_ERR_UNSUPPORTED.  The region of interest (roi_flags: ROI_LOAD,
 * ROI_SAVE, ROI_MAP) is implemented; its limits and buffers are set with soc_set_roi(). */
typedef struct soc_params {
    int32_t bins;              /* BINS: length of the DSC / CSC tables                        */
    int32_t no_ps;             /* NO_PS (>=1)                                                 */
    int32_t ps_method;         /* PS_METHOD 0,1,2,4,5                                         */
    int32_t with_abu;          /* WITH_ABU: per-cell opacities in OPT[2*CELLS]                */
    int32_t with_ali;          /* WITH_ALI: self-absorptions of SimRAM_CL go to XAB           */
    int32_t noabsorbed;        /* NOABSORBED==0 -> per-frequency absorptions into INT         */
    int32_t save_intensity;    /* SAVE_INTENSITY 0,1,2                                        */
    int32_t use_emweight;      /* USE_EMWEIGHT 0,1                                            */
    int32_t hpbg_weighted;     /* HPBG_WEIGHTED                                               */
    int32_t ffs;               /* FFS: forced first scattering (scattered-light kernels)      */
    int32_t step_weight;       /* STEP_WEIGHT <=0,1,2                                         */
    int32_t level_threshold;   /* LEVEL_THRESHOLD (maps)                                      */
    int32_t with_msf;          /* WITH_MSF: one scattering function per dust species; needs with_abu, ABU, ABSV, SCAV and
                                  DSC / CSC of ndust*bins entries (kernel_ASOC.c:777-794)        */
    int32_t mirror;            /* MIRROR bit mask: 1 x=0, 2 x=NX, 4 y=0, 8 y=NY, 16 z=0, 32 z=NZ are reflecting borders
                                  (ASOC.py:319-321, kernel_ASOC_aux.c:1054)                      */
    int32_t dir_weight, do_split;              /* must be 0                                      */
    int32_t roi_flags;         /* 1 WITH_ROI_LOAD, 2 WITH_ROI_SAVE, 4 ROI_MAP; geometry through soc_set_roi()     */
    int32_t map_interpolation; /* MAP_INTERPOLATION 0,1,2 (kernel_ASOC_map.c:656-811)            */
    float   sw_a, sw_b;        /* SW_A, SW_B                                                  */
    float   length;            /* LENGTH = GL*PARSEC rounded as "%.5e" (ASOC.py:347,356)      */
    float   factor;            /* FACTOR (1e20)                                               */
    float   adhoc;             /* ADHOC (1.0)                                                 */
    float   reserved;
    int32_t ndust;             /* NDUST (only read when with_msf != 0)                            */
    int32_t opt_is_half;       /* OPT_IS_HALF: soc_upload(SOC_BUF_OPT) takes IEEE half values (ASOC.py:1155); they are
                                  widened on the device, the kernels see exactly the half-rounded opacities  */
    int32_t ref_quirks;        /* behaviour of the shipped reference where this library deliberately differs (ini key REFQUIRKS):
                                  bit 0 (1): scattered-light SimRAM_HP / SimRAM_CL weight their peel-off rays with the `#ifdef
                                  HG_TEST` branch that is live in kernel_ASOC_sca.c (analytic g = 0.65 phase function times
                                  1-exp(-tau), :349-355, 1343-1349) instead of DSC * exp(-tau);
                                  bit 1 (2): per-level maps step with the Index() copy of kernel_ASOC_map_H.c, which drops the
                                  root coordinates when a ray climbs into a root-grid leaf (:250).  0 = the intended behaviour */
    int32_t reserved2[1];
} soc_params;

/* Device buffers.  Names are those of the reference's kernel arguments. */
enum soc_buffer {
    SOC_BUF_DENS = 0, SOC_BUF_PAR, SOC_BUF_TABS, SOC_BUF_XAB, SOC_BUF_INT, SOC_BUF_INTX, SOC_BUF_INTY,
    SOC_BUF_INTZ, SOC_BUF_EMIT, SOC_BUF_EMWEI, SOC_BUF_OPT, SOC_BUF_DSC, SOC_BUF_CSC, SOC_BUF_PSPOS,
    SOC_BUF_PS, SOC_BUF_XPS_NSIDE, SOC_BUF_XPS_SIDE, SOC_BUF_XPS_AREA, SOC_BUF_HPBG, SOC_BUF_HPBGP,
    SOC_BUF_MAP, SOC_BUF_SAVETAU, SOC_BUF_OUT, SOC_BUF_ODIR, SOC_BUF_ORA, SOC_BUF_ODE, SOC_BUF_TTT,
    SOC_BUF_TNEW, SOC_BUF_FABS,
    SOC_BUF_ABU,               /* WITH_MSF: abundances [CELLS*NDUST], dust index fastest        */
    SOC_BUF_ABSV, SOC_BUF_SCAV, /* WITH_MSF: the ABS / SCA kernel arguments as vectors [NDUST]  */
    SOC_BUF_ROI_LOAD,          /* WITH_ROI_LOAD: external field [elements * 12*ROI_NSIDE^2] of one frequency */
    SOC_BUF_ROI_SAVE,          /* WITH_ROI_SAVE: photons entering ROI [elements * 12*ROI_NSIDE^2]           */
    SOC_BUF_COUNT
};

/* Stream layout of the Monte Carlo kernels. */
enum soc_rng_mode {
    SOC_RNG_REFERENCE = 0,     /* MWC64X, one stream per reference work item (bit-compatible seeding) */
    SOC_RNG_PACKET = 1         /* counter-based Philox4x32-10, one stream per photon packet            */
};

/* Work counters accumulated by the kernels. */
typedef struct soc_counters {
    uint64_t packets;          /* emitted photon packets                                  */
    uint64_t steps;            /* cell-steps: absorption updates incl. partial steps; for the
                                  scattered-light and map kernels every GetStep counts   */
    uint64_t scatterings;
    uint64_t peels;            /* peel-off rays (scattered-light kernels)                 */
    uint64_t launches;         /* kernel launches issued by this context                  */
    uint64_t reserved[3];
} soc_counters;

const char *soc_last_error(void);
int  soc_version(void);

int  soc_create(int device_ordinal, soc_context **ctx);
int  soc_destroy(soc_context *ctx);
int  soc_sync(soc_context *ctx);

int  soc_set_params(soc_context *ctx, const soc_params *p);
int  soc_set_grid(soc_context *ctx, int32_t nx, int32_t ny, int32_t nz, int32_t levels, int64_t cells,
                  const int32_t *lcells, const int32_t *off, const float *dens);
/* Region of interest (ASOC.py:906-945; kernel_ASOC.c:44-51, 141-179, 469-502, 615-643; kernel_ASOC_aux.c:1031):
 * roi = [x0,x1,y0,y1,z0,z1] inclusive root-cell limits, roi_step = subdivision of a root cell on the ROI surface,
 * roi_nside = Healpix NSIDE of the stored directions, roi_dim = dimensions (nx,ny,nz) of the file loaded with
 * WITH_ROI_LOAD.  Which of the three ROI options are active is soc_params.roi_flags.  With WITH_ROI_SAVE the buffer
 * SOC_BUF_ROI_SAVE is (re)allocated and cleared here; soc_sim_pb(source = 3) reads SOC_BUF_ROI_LOAD. */
int  soc_set_roi(soc_context *ctx, const int32_t roi[6], int roi_step, int roi_nside, const int32_t roi_dim[3]);
int  soc_set_rng_mode(soc_context *ctx, int mode);
int  soc_set_shard(soc_context *ctx, int rank, int world);
/* Accumulation engine of the absorption counters: deposit_mode 0 = one red.global.add.f32 per lane and step,
 * 1 = lanes of a warp that hit the same cell are combined first, 2 = 1 + a shared-memory tile of cells around
 * the point source.  refill_lanes: a warp takes new packets when at least this many lanes are idle (1..32).
 * aggregate_steps: lanes are combined only while some packet of the warp is younger than this many steps. */
int  soc_set_tuning(soc_context *ctx, int deposit_mode, int refill_lanes, int aggregate_steps);

/* Cell stepping of the production layout on regular grids: 0 (default) = exact incremental DDA, 1 = the
 * reference's GetStep arithmetic (PEPS overshoot, 2*PEPS pull-back at scatterings, kernel_ASOC_aux.c:282-315).
 * The two differ only where the mean free path is comparable to PEPS = 1e-4 cells. */
int  soc_set_geometry(soc_context *ctx, int mode);

/* Cell order of the density / scratch accumulator inside the production kernel for regular grids with even
 * dimensions: 1 (default) = 2x2x2 bricks (one 32-byte sector per brick), 0 = the reference's x-fastest order.
 * Buffers seen through this ABI always use the reference order. */
int  soc_set_layout(soc_context *ctx, int mode);

/* Domain-tiled propagation on regular grids (production layout): the grid is cut into boxes of at most `edge` cells per
 * axis, the packet kernels work on one box at a time so that its density / accumulator arrays stay in the L2, and a
 * packet leaving a box through an interior face is parked (its full stepping state) until the box it enters is
 * processed -- same packets, same paths, same results up to the order of the float additions.  edge = 0 (default):
 * automatic, boxes of <= 256 cells per axis when DENS + the scratch accumulator exceed the L2 (more than 2^25 cells);
 * edge < 0: off; edge > 0 (even): forced with that box size.  Launches in this mode return when the packets are done.
 * (Point-source launches with deposit_mode 2 on smaller regular grids use the same queues: emission and the steps inside
 * the shared-memory tile run as a pass of their own that parks the packets at the border of the tile, the rest of the paths
 * through the plain-add kernel -- DESIGN.md 4.1; also synchronous.) */
int  soc_set_domains(soc_context *ctx, int edge);

/* OPT[CELLS,2] = per-cell (KABS, KSCA) built on the device from the abundances in buffer ABU ([CELLS, ndust] floats,
 * uploaded once) and the per-species cross sections of one frequency: replaces the host loop of ASOC.py:1146-1161 and
 * its 8*CELLS-byte upload per frequency.  Species `first` .. ndust-1 are summed (the reference starts from species 1 in
 * its cell-emission loop, ASOC.py:1673); single_abu: OPT = a*K[0] + (1-a)*K[1] with one abundance per cell.  Same
 * float32 operation order as the host code; with OPT_IS_HALF the values are rounded to half precision. */
int  soc_build_opt(soc_context *ctx, int ndust, const float *kabs, const float *ksca, int first, int single_abu);

int  soc_upload(soc_context *ctx, int buffer, const void *host, size_t nbytes);
int  soc_download(soc_context *ctx, int buffer, void *host, size_t nbytes);
int  soc_clear(soc_context *ctx, int buffer, size_t nbytes);     /* allocate (if needed) and zero */
void *soc_device_ptr(soc_context *ctx, int buffer, size_t *nbytes);

int  soc_zero_amc(soc_context *ctx, int tag);

/* `global` is the reference launch's global work size (number of work items); it defines the packet
 * decomposition exactly as ASOC.py:1032-1092 does. */
int  soc_sim_pb(soc_context *ctx, int source, int packets, int batch, float seed, float abs, float sca,
                float bg, float tw, int global);
int  soc_sim_hp(soc_context *ctx, int packets, int batch, float seed, float abs, float sca, float tw, int global);
int  soc_sim_cl(soc_context *ctx, int source, int packets, int batch, float seed, float abs, float sca, float tw,
                int global);

/* EqTemperature: reads EMIT (absorbed energy) and DENS, writes TNEW for cells of `level`; TTT holds the
 * E->T table.  Emission: reads TNEW, writes EMIT. */
int  soc_eq_temperature(soc_context *ctx, int level, float adhoc, float kE, float Emin, int NE);
int  soc_emission(soc_context *ctx, float freq, float fabs);
/* Emission2 (kernel_ASOC_aux.c:862; ASOC.py:2157-2180, key EBATCH): emission of cells [c0,c1[ at all `nfreq` frequencies from
 * TNEW in one launch; emit_out[(cell-c0)*nfreq + ifreq] on the host -- the layout of the emitted file. */
int  soc_emission2(soc_context *ctx, int c0, int c1, int nfreq, const float *freq, const float *fabs, float *emit_out);

/* Device-resident absorbed file (the [CELLS, NFREQ] array of ASOC.py:1482-1497 and its final scaling,
 * ASOC.py:2782-2878): _begin allocates and clears FABS[cells*nfreq] (frequency fastest, the file's layout),
 * _add does FABS[:, ifreq] += INT, _finish multiplies every cell by coeff0 * 8^level / DENS, writes -1e20 into
 * cells with DENS <= nnnlimit (parents are links, i.e. <= 0) and copies the array to `host` (cells*nfreq floats;
 * NULL = leave it on the device, e.g. to all-reduce it first with finish_scale = 0). */
int  soc_absorbed_begin(soc_context *ctx, int nfreq);
/* Hand-off of the absorptions to the dust solver of one species (A2E_MABU.py:700-705, kernel split_absorbed of
 * kernel_A2E_MABU_aux.c:3-24): host[cell, f] = FABS[cell, f] * rabs[f, idust] / sum_d ABU[cell, d] * rabs[f, d], on the
 * [CELLS, nfreq] array in buffer FABS (left there by soc_absorbed_finish, or uploaded) and the abundances in buffer
 * ABU [CELLS, ndust]; rabs = [nfreq, ndust] doubles.  Same arithmetic as the reference kernel. */
int  soc_split_absorbed(soc_context *ctx, int idust, int ndust, int nfreq, const double *rabs, float *host);
int  soc_absorbed_add(soc_context *ctx, int ifreq);
int  soc_absorbed_finish(soc_context *ctx, float coeff0, float nnnlimit, int finish_scale, float *host);

/* Mapping: reads EMIT, DENS (OPT); writes MAP and SAVETAU [npix_y*npix_x]. intobs[0] <= -1e10 selects the
 * orthographic projection. */
int  soc_mapping(soc_context *ctx, float map_dx, int npix_x, int npix_y, const float dir[3], const float ra[3],
                 const float de[3], float abs, float sca, const float centre[3], const float intobs[3],
                 int save_colden);
int  soc_healpix_mapping(soc_context *ctx, int nside, float abs, float sca, const float intobs[3], int save_colden);
/* Per-level Mapping (kernel_ASOC_map_H.c:380; ASOC.py:3320-3440, ini `mapping nx ny dx 999`): one image per hierarchy
 * level, MAP[LEVELS*npix_y*npix_x] (level-major); with save_colden the column density (x LENGTH) goes to SAVETAU.
 * The reference's copy of Index() in that file drops rays that climb into root-grid leaves; this
 * entry point steps like soc_mapping. */
int  soc_mapping_levels(soc_context *ctx, float map_dx, int npix_x, int npix_y, const float dir[3], const float ra[3],
                        const float de[3], float abs, float sca, const float centre[3], const float intobs[3],
                        int save_colden);

/* PSTau (kernel_ASOC_map.c:1545; ASOC.py:3576-3644): optical depth and column density (x LENGTH) from each of the
 * first `no` point sources in PSPOS towards the observer direction `dir`; results are copied to the host arrays. */
int  soc_ps_tau(soc_context *ctx, int no, const float dir[3], float abs, float sca, float *colden_out, float *tau_out);

/* Scattered light (ASOCS.py).  ODIR/ORA/ODE are [ndir*3] floats (x,y,z), OUT is [ndir*npix_y*npix_x].
 * ndir < 0 (ASOCS.py:44-47): one Healpix image of NSIDE = -ndir, RING order, seen by an observer at the position
 * ODIR[0..2] (root-grid units); OUT is [12*NSIDE*NSIDE] and npix / map_dx / centre are ignored. */
int  soc_sca_zero_out(soc_context *ctx, int ndir, int npix_x, int npix_y);
int  soc_sca_ps(soc_context *ctx, int packets, int batch, float seed, float abs, float sca, int ndir, int npix_x,
                int npix_y, float map_dx, const float centre[3], int global);
int  soc_sca_pb(soc_context *ctx, int source, int packets, int batch, float seed, float abs, float sca, float bg,
                int ndir, int npix_x, int npix_y, float map_dx, const float centre[3], int global);
/* Healpix background (HPBG [+ HPBGP]) and emission from the cells (EMIT [+ EMWEI]) as sources of scattered light */
int  soc_sca_hp(soc_context *ctx, int packets, int batch, float seed, float abs, float sca, int ndir, int npix_x,
                int npix_y, float map_dx, const float centre[3], int global);
int  soc_sca_cl(soc_context *ctx, int source, int packets, int batch, float seed, float abs, float sca, int ndir,
                int npix_x, int npix_y, float map_dx, const float centre[3], int global);

int  soc_get_counters(soc_context *ctx, soc_counters *out);
int  soc_reset_counters(soc_context *ctx);
int  soc_last_launch_ms(soc_context *ctx, float *ms);     /* device time of the most recent kernel launch */
const char *soc_last_kernel(soc_context *ctx);            /* name (with its template arguments) of the packet kernel the most recent soc_sim_* call dispatched */
void *soc_stream(soc_context *ctx);                       /* the context's cudaStream_t (for NCCL / event timing) */

#ifdef __cplusplus
}
#endif
#endif
