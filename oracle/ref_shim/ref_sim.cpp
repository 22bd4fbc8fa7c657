// TEST INFRASTRUCTURE ONLY.  Runs the reference's kernel_ASOC.c (compiled in
// place from /root/reference through cl_shim.h) on the host cores.  The macro
// set (-D NX=.. etc., ASOC.py:344-362) is given on the compiler command line by
// oracle/build_ref.py, one shared object per grid/config like the reference JIT.
#include "ref_common.h"

thread_local size_t clshim_gid = 0, clshim_gsize = 1;
thread_local unsigned long clshim_atomic_ok = 0;
static unsigned long g_atomic_total = 0;
long ref_stride = 1, ref_offset = 0, ref_gsize = 0;

// override the shim's CAS so that successful updates are counted
static inline unsigned int counted_cmpxchg(volatile unsigned int *p, unsigned int cmp, unsigned int val) {
    unsigned int old = __sync_val_compare_and_swap(p, cmp, val);
    if (old == cmp) clshim_atomic_ok++;
    return old;
}
#define atomic_cmpxchg counted_cmpxchg

namespace refk {
#include "kernel_ASOC.c"
}
#undef atomic_cmpxchg

static void flush_counter() {
    unsigned long tot = 0;
    #pragma omp parallel reduction(+:tot)
    { tot += clshim_atomic_ok; clshim_atomic_ok = 0; }
    g_atomic_total += tot;
}

extern "C" {

int ref_threads(void) { return omp_get_max_threads(); }
void ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
// work items handed to a host thread at a time (OpenMP dynamic schedule); 256 unless changed
void ref_set_chunk(int n) { omp_set_schedule(omp_sched_dynamic, n > 0 ? n : 256); }
static struct RefInitSchedule { RefInitSchedule() { omp_set_schedule(omp_sched_dynamic, 256); } } ref_init_schedule;
void ref_set_sampling(long stride, long offset, long gsize) { ref_stride = stride > 0 ? stride : 1; ref_offset = offset; ref_gsize = gsize; }
unsigned long ref_atomic_count(int reset) { flush_counter(); unsigned long v = g_atomic_total; if (reset) g_atomic_total = 0; return v; }

// MWC64X known-answer helper: first n outputs of the stream of work item `id`
void ref_rng_stream(float seed, long id, long gsize, int n, unsigned int *out, unsigned int *state_xc) {
    clshim_gid = (size_t)id; clshim_gsize = (size_t)gsize;
    refk::mwc64x_state_t rng;
    ulong samplesPerStream = 274877906944L;
    refk::MWC64X_SeedStreams(&rng, (unsigned long)(fmod(seed * 7.0f * 3.1415926535897f, 1.0f) * 4294967296L), samplesPerStream);
    state_xc[0] = rng.x; state_xc[1] = rng.c;
    for (int i = 0; i < n; i++) out[i] = refk::MWC64X_NextUint(&rng);
}
void ref_rng_stream_base(unsigned long base, long id, int n, unsigned int *out, unsigned int *state_xc) {
    clshim_gid = (size_t)id; clshim_gsize = 1;
    refk::mwc64x_state_t rng;
    refk::MWC64X_SeedStreams(&rng, base, 274877906944L);
    state_xc[0] = rng.x; state_xc[1] = rng.c;
    for (int i = 0; i < n; i++) out[i] = refk::MWC64X_NextUint(&rng);
}

void ref_parents(int global, float *DENS, const int *LCELLS, const int *OFF, int *PAR) {
    REF_PARALLEL_FOR(global, refk::Parents(DENS, LCELLS, OFF, PAR));
}

void ref_sim_pb(int global, int SOURCE, int PACKETS, int BATCH, float SEED, float *ABS, float *SCA, float BG,
                float *PSPOS_xyz, float *PS, float TW, const int *LCELLS, const int *OFF, int *PAR, float *DENS,
                float *EMIT, float *TABS, const float *DSC, const float *CSC, float *XAB, float *EMWEI,
                float *INT, float *INTX, float *INTY, float *INTZ, float *OPT, float *ABU,
                int *XPS_NSIDE, int *XPS_SIDE, float *XPS_AREA, const int *ROI_DIM, float *ROI_LOAD, const int *ROI,
                float *ROI_SAVE) {
    // WITH_ROI_LOAD / WITH_ROI_SAVE append arguments to the kernel (kernel_ASOC.c:44-51)
    REF_PARALLEL_FOR(global,
        refk::SimRAM_PB(SOURCE, PACKETS, BATCH, SEED, ABS, SCA, BG, (float3 *)PSPOS_xyz, PS, TW, LCELLS, OFF, PAR,
                        DENS, EMIT, TABS, DSC, CSC, XAB, EMWEI, INT, INTX, INTY, INTZ, OPT, ABU,
                        XPS_NSIDE, XPS_SIDE, XPS_AREA
#if (WITH_ROI_LOAD)
                        , ROI_DIM, ROI_LOAD
#endif
#if (WITH_ROI_SAVE)
                        , ROI, ROI_SAVE
#endif
                        ));
    flush_counter();
}

void ref_sim_hp(int global, int PACKETS, int BATCH, float SEED, float *ABS, float *SCA, float TW,
                const int *LCELLS, const int *OFF, int *PAR, float *DENS, float *EMIT, float *TABS,
                const float *DSC, const float *CSC, float *XAB, float *INT, float *INTX, float *INTY, float *INTZ,
                float *OPT, float *BG, float *HPBGP, float *ABU) {
    REF_PARALLEL_FOR(global,
        refk::SimRAM_HP(PACKETS, BATCH, SEED, ABS, SCA, TW, LCELLS, OFF, PAR, DENS, EMIT, TABS, DSC, CSC, XAB,
                        INT, INTX, INTY, INTZ, OPT, BG, HPBGP, ABU));
    flush_counter();
}

void ref_sim_cl(int global, int SOURCE, int PACKETS, int BATCH, float SEED, float *ABS, float *SCA, float TW,
                const int *LCELLS, const int *OFF, int *PAR, float *DENS, float *EMIT, float *TABS,
                const float *DSC, const float *CSC, float *XAB, float *EMWEI, float *INT, float *INTX, float *INTY,
                float *INTZ, int *EMINDEX, float *OPT, float *ABU, const int *ROI, float *ROI_SAVE) {
    REF_PARALLEL_FOR(global,
        refk::SimRAM_CL(SOURCE, PACKETS, BATCH, SEED, ABS, SCA, TW, LCELLS, OFF, PAR, DENS, EMIT, TABS, DSC, CSC,
                        XAB, EMWEI, INT, INTX, INTY, INTZ, EMINDEX, OPT, ABU
#if (WITH_ROI_SAVE)
                        , ROI, ROI_SAVE
#endif
                        ));
    flush_counter();
}

void ref_eq_temperature(int global, int level, float adhoc, float kE, float Emin, int NE, int *OFF, int *LCELLS,
                        float *TTT, float *DENS, float *EMIT, float *TNEW) {
    REF_PARALLEL_FOR(global, refk::EqTemperature(level, adhoc, kE, Emin, NE, OFF, LCELLS, TTT, DENS, EMIT, TNEW));
}

void ref_emission(int global, float FREQ, float FABS, float *DENS, float *T, float *EMIT) {
    REF_PARALLEL_FOR(global, refk::Emission(FREQ, FABS, DENS, T, EMIT));
}

void ref_emission2(int global, int c0, int c1, int nfreq, float *FREQ, float *FABS, float *DENS, float *T, float *EMIT) {
    REF_PARALLEL_FOR(global, refk::Emission2(c0, c1, nfreq, FREQ, FABS, DENS, T, EMIT));
}

}  // extern "C"
