// soc_b200 -- cell-to-cell stepping through the octree hierarchy as an incremental ray walk.
//
// The reference finds the next cell from floating-point positions in level-local coordinates: GetStep does
// 3 fmod + 3 divisions, Index climbs through PAR testing the position against the octet and descends through
// the links re-deriving child indices from the position (kernel_ASOC_aux.c:198-315).  The walker below carries,
// instead of a position, the distances tx,ty,tz along the ray (in root-grid units) to the next x/y/z face of
// the current leaf, plus 1/|d|.  Crossing a face is then integer work:
//   * neighbour inside the same octet  -> flip one bit of the cell index (no PAR read),
//   * otherwise climb one level through PAR; the face distances of the other two axes grow by one child size
//     when the child sat in the half away from the ray's exit side,
//   * descend through links: the child is picked by comparing each face distance with the child size.
// No division, fmod, float->int conversion or position arithmetic per step; geometry is exact (no PEPS
// overshoot).  The position inside a cell, when needed (scattering), follows from frac = d>0 ? 1-t|d|/s : t|d|/s.
#pragma once
#include "common.cuh"

struct Walker {
    float tx, ty, tz;          // distance along the ray to the next face per axis, root-grid units
    float rdx, rdy, rdz;       // 1/|d|
    vec3 d;
    int level, ind;            // current leaf: level and index within the level; ind < 0 = outside the cloud
    int ix, iy, iz;            // root-grid coordinates of the level-0 ancestor
    int up;                    // bit b set: the ray moves towards +axis b
    float rho;                 // density of the current leaf
};

__device__ __forceinline__ float cell_size(int level) { return __int_as_float((127 - level) << 23); }   // 2^-level

__device__ __forceinline__ void walker_set_direction(Walker &w, const vec3 &d, float fx, float fy, float fz) {
    const float s = cell_size(w.level);
    w.d = d;
    w.up = (d.x > 0.0f ? 1 : 0) | (d.y > 0.0f ? 2 : 0) | (d.z > 0.0f ? 4 : 0);
    w.rdx = __fdividef(1.0f, fabsf(d.x)); w.rdy = __fdividef(1.0f, fabsf(d.y)); w.rdz = __fdividef(1.0f, fabsf(d.z));
    fx = fminf(fmaxf(fx, 0.0f), 1.0f); fy = fminf(fmaxf(fy, 0.0f), 1.0f); fz = fminf(fmaxf(fz, 0.0f), 1.0f);
    w.tx = ((d.x > 0.0f) ? 1.0f - fx : fx) * s * w.rdx;
    w.ty = ((d.y > 0.0f) ? 1.0f - fy : fy) * s * w.rdy;
    w.tz = ((d.z > 0.0f) ? 1.0f - fz : fz) * s * w.rdz;
}

// fractional position inside the current cell, from the face distances
__device__ __forceinline__ void walker_fraction(const Walker &w, float &fx, float &fy, float &fz) {
    const float is = __int_as_float((127 + w.level) << 23);       // 2^level = 1/size
    fx = w.tx * fabsf(w.d.x) * is; fy = w.ty * fabsf(w.d.y) * is; fz = w.tz * fabsf(w.d.z) * is;
    if (w.d.x > 0.0f) fx = 1.0f - fx;
    if (w.d.y > 0.0f) fy = 1.0f - fy;
    if (w.d.z > 0.0f) fz = 1.0f - fz;
}

// Start a walk at a point located by index_global(): `pos` in the level-local coordinates of (level, ind).
template <bool OCT>
__device__ __forceinline__ void walker_init(const GridDesc &G, Walker &w, const vec3 &pos, const vec3 &dir, int level, int ind, float rho) {
    w.level = level; w.ind = ind; w.rho = rho;
    walker_set_direction(w, dir, pos.x - floorf(pos.x), pos.y - floorf(pos.y), pos.z - floorf(pos.z));
    int root = ind;
    if (OCT) for (int l = level; l > 0; l--) root = G.par[G.off[l] + root - G.nxyz];
    w.ix = root % G.nx; w.iy = (root / G.nx) % G.ny; w.iz = root / (G.nx * G.ny);
}

// Cross the face of axis `ax` into the neighbouring leaf.  Written without per-axis code paths (lanes of a warp
// cross different axes): the three axes are handled by predicated updates driven by bit masks --
// `up` (ray direction), `sid ^ up` (child on the near side of the octet, seen along the ray).
template <bool OCT>
__device__ __forceinline__ void walker_cross(const GridDesc &G, Walker &w, const int ax) {
    const int abit = 1 << ax;
    const bool up = (w.up & abit) != 0;
    const float rda = (ax == 0) ? w.rdx : ((ax == 1) ? w.rdy : w.rdz);
    float size = OCT ? cell_size(w.level) : 1.0f;
    float ta;                                                      // new face distance of the crossed axis
    for (;;) {
        if (!OCT || w.level == 0) {
            const int sgn = up ? 1 : -1;
            const int c = ((ax == 0) ? w.ix : ((ax == 1) ? w.iy : w.iz)) + sgn;
            const int lim = (ax == 0) ? G.nx : ((ax == 1) ? G.ny : G.nz);
            if ((unsigned)c >= (unsigned)lim) { w.ind = -1; return; }
            w.ix += (ax == 0) ? sgn : 0; w.iy += (ax == 1) ? sgn : 0; w.iz += (ax == 2) ? sgn : 0;
            w.ind = (w.iz * G.ny + w.iy) * G.nx + w.ix;
            ta = rda;
            break;
        }
        const int near = (w.ind ^ w.up) & 7;                        // bit b: the child sits on the near side of axis b
        if (near & abit) {                                          // the neighbour is a sibling in the same octet
            w.ind ^= abit;
            ta = size * rda;
            break;
        }
        // leave the octet: one level up; faces of the other axes move out by one child size where the child was near
        const int grow = near & ~abit;
        if (grow & 1) w.tx += size * w.rdx;
        if (grow & 2) w.ty += size * w.rdy;
        if (grow & 4) w.tz += size * w.rdz;
        w.ind = G.par[G.off[w.level] + w.ind - G.nxyz];
        w.level--;
        size *= 2.0f;
    }
    float rho = G.dens[(OCT ? G.off[w.level] : 0) + w.ind];
    if (OCT) {
        while (!is_leaf(rho)) {                                     // descend through the links
            const int base = link_index(rho);
            w.level++;
            size *= 0.5f;
            ta = size * rda;
            const float fx = size * w.rdx, fy = size * w.rdy, fz = size * w.rdz;
            int near = abit;                                        // entered through the near face of the crossed axis
            if (ax != 0 && w.tx > fx) { w.tx -= fx; near |= 1; }
            if (ax != 1 && w.ty > fy) { w.ty -= fy; near |= 2; }
            if (ax != 2 && w.tz > fz) { w.tz -= fz; near |= 4; }
            w.ind = base + ((w.up ^ near) & 7);
            rho = G.dens[G.off[w.level] + w.ind];
        }
    }
    if (ax == 0) w.tx = ta; else if (ax == 1) w.ty = ta; else w.tz = ta;
    w.rho = rho;
}

// Advance to the next leaf along the ray.  Returns the path length inside the cell that is left (root-grid
// units).  Afterwards w.ind < 0 if the ray has left the cloud.
template <bool OCT>
__device__ __forceinline__ float walker_advance(const GridDesc &G, Walker &w) {
    const float tmin = fminf(w.tx, fminf(w.ty, w.tz));
    const int ax = (w.tx <= w.ty && w.tx <= w.tz) ? 0 : ((w.ty <= w.tz) ? 1 : 2);
    w.tx -= tmin; w.ty -= tmin; w.tz -= tmin;
    walker_cross<OCT>(G, w, ax);
    return fmaxf(tmin, 0.0f);
}

// ---- the same walk, one hop at a time ---------------------------------------------------------------------------
// Lanes of a warp need different numbers of climbs and descents per cell crossing; running them as inner loops
// leaves most lanes idle (measured: 8.7 of 32 lanes active).  The kernels therefore keep a per-lane phase and
// perform at most one climb, one crossing attempt and one descent per iteration of their main loop, so that
// all lanes always execute the same short blocks.
enum WalkPhase { WALK_LEAF = 0, WALK_CLIMB = 1, WALK_DESCEND = 2, WALK_CROSS = 3, WALK_SCATTER = 4, WALK_END = 5 };

// one level up (phase CLIMB -> CROSS): faces of the other two axes move out where the child was on the near side
__device__ __forceinline__ void nav_climb(const GridDesc &G, Walker &w, const int ax) {
    const float size = cell_size(w.level);
    const int grow = ((w.ind ^ w.up) & 7) & ~(1 << ax);
    if (grow & 1) w.tx += size * w.rdx;
    if (grow & 2) w.ty += size * w.rdy;
    if (grow & 4) w.tz += size * w.rdz;
    w.ind = G.par[G.off[w.level] + w.ind - G.nxyz];
    w.level--;
}

// try to cross the face of axis `ax` at the current level.  Returns the new phase: LEAF / DESCEND when the
// neighbour was found (its value is loaded into w.rho), CLIMB when the octet has to be left first; w.ind < 0
// when the ray leaves the cloud.
// `mirror`: MIRROR bit mask (1 x=0, 2 x=NX, 4 y=0, 8 y=NY, 16 z=0, 32 z=NZ): at a reflecting border the ray stays in the
// root cell it has climbed to, the crossed direction component changes sign and the walk descends again towards the
// leaf at the border (the descent enters "through the near face" of the crossed axis, which now is the border).
__device__ __forceinline__ int nav_cross(const GridDesc &G, Walker &w, const int ax, const int mirror = 0) {
    const int abit = 1 << ax;
    const bool up = (w.up & abit) != 0;
    const float rda = (ax == 0) ? w.rdx : ((ax == 1) ? w.rdy : w.rdz);
    float ta;
    if (w.level == 0) {
        const int sgn = up ? 1 : -1;
        const int c = ((ax == 0) ? w.ix : ((ax == 1) ? w.iy : w.iz)) + sgn;
        const int lim = (ax == 0) ? G.nx : ((ax == 1) ? G.ny : G.nz);
        if ((unsigned)c >= (unsigned)lim) {
            if (mirror & ((up ? 2 : 1) << (2 * ax))) {
                w.up ^= abit;
                if (ax == 0) w.d.x = -w.d.x; else if (ax == 1) w.d.y = -w.d.y; else w.d.z = -w.d.z;
            } else { w.ind = -1; return WALK_LEAF; }
        } else {
            w.ix += (ax == 0) ? sgn : 0; w.iy += (ax == 1) ? sgn : 0; w.iz += (ax == 2) ? sgn : 0;
            w.ind = (w.iz * G.ny + w.iy) * G.nx + w.ix;
        }
        ta = rda;
    } else {
        if ((((w.ind ^ w.up) & 7) & abit) == 0) return WALK_CLIMB;
        w.ind ^= abit;
        ta = cell_size(w.level) * rda;
    }
    if (ax == 0) w.tx = ta; else if (ax == 1) w.ty = ta; else w.tz = ta;
    w.rho = G.dens[G.off[w.level] + w.ind];
    return is_leaf(w.rho) ? WALK_LEAF : WALK_DESCEND;
}

// one level down through the link in w.rho (phase DESCEND); the cell was entered through the near face of `ax`
__device__ __forceinline__ int nav_descend(const GridDesc &G, Walker &w, const int ax) {
    const int base = link_index(w.rho);
    w.level++;
    const float size = cell_size(w.level);
    const float fx = size * w.rdx, fy = size * w.rdy, fz = size * w.rdz;
    int near = 1 << ax;
    if (ax == 0) w.tx = fx; else if (w.tx > fx) { w.tx -= fx; near |= 1; }
    if (ax == 1) w.ty = fy; else if (w.ty > fy) { w.ty -= fy; near |= 2; }
    if (ax == 2) w.tz = fz; else if (w.tz > fz) { w.tz -= fz; near |= 4; }
    w.ind = base + ((w.up ^ near) & 7);
    w.rho = G.dens[G.off[w.level] + w.ind];
    return is_leaf(w.rho) ? WALK_LEAF : WALK_DESCEND;
}
