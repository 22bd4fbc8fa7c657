// soc_b200 -- octree stepping through a per-cell neighbour table.
//
// walk.cuh finds the cell behind a face by climbing through PAR until the face is inside an octet: one loop
// iteration of the kernel per level climbed, with only the lanes that climb active.  Here every cell carries the
// six cells behind its faces -- the neighbour of the same level where the hierarchy has one, else the coarser leaf
// that covers it, together with that cell's density or link -- built once per grid on the device (48 B per cell).  A face crossing is then one table look-up
// plus integer arithmetic on the cell coordinates (cx,cy,cz at the cell's own level): no climb, the only level
// changes left are descents through the links of a refined neighbour (one DENS read per level, as before).
#pragma once
#include "common.cuh"

#define SOC_NBR_LEVEL_SHIFT 27                      // entry = level << 27 | global cell index; -1 = outside the cloud
#define SOC_NBR_INDEX_MASK ((1 << SOC_NBR_LEVEL_SHIFT) - 1)

struct LWalker {
    float tx, ty, tz;          // distance along the ray to the next face per axis, root-grid units
    float rdx, rdy, rdz;       // 1/|d|; the direction itself is not kept (lw_dir: |d_i| = 1/rd_i, sign from `up`)
    int level, cell;           // current cell: level and GLOBAL index (OFF[level] + index within the level)
    int cx, cy, cz;            // integer coordinates of the cell at its own level (root coordinate * 2^level + ...)
    int up;                    // bit b set: the ray moves towards +axis b
    float rho;                 // density of the current leaf (or the link while descending)
};

__device__ __forceinline__ float lw_size(int level) { return __int_as_float((127 - level) << 23); }   // 2^-level

__device__ __forceinline__ void lw_set_direction(LWalker &w, const vec3 &d, float fx, float fy, float fz) {
    const float s = lw_size(w.level);
    w.up = (d.x > 0.0f ? 1 : 0) | (d.y > 0.0f ? 2 : 0) | (d.z > 0.0f ? 4 : 0);
    w.rdx = __fdividef(1.0f, fabsf(d.x)); w.rdy = __fdividef(1.0f, fabsf(d.y)); w.rdz = __fdividef(1.0f, fabsf(d.z));
    fx = fminf(fmaxf(fx, 0.0f), 1.0f); fy = fminf(fmaxf(fy, 0.0f), 1.0f); fz = fminf(fmaxf(fz, 0.0f), 1.0f);
    w.tx = ((d.x > 0.0f) ? 1.0f - fx : fx) * s * w.rdx;
    w.ty = ((d.y > 0.0f) ? 1.0f - fy : fy) * s * w.rdy;
    w.tz = ((d.z > 0.0f) ? 1.0f - fz : fz) * s * w.rdz;
}

// direction of the ray, rebuilt from 1/|d| and the sign bits (used at scatterings and ray ends only)
__device__ __forceinline__ vec3 lw_dir(const LWalker &w) {
    const float ax = __fdividef(1.0f, w.rdx), ay = __fdividef(1.0f, w.rdy), az = __fdividef(1.0f, w.rdz);
    vec3 d = { (w.up & 1) ? ax : -ax, (w.up & 2) ? ay : -ay, (w.up & 4) ? az : -az };
    return d;
}

__device__ __forceinline__ void lw_fraction(const LWalker &w, float &fx, float &fy, float &fz) {
    const float is = __int_as_float((127 + w.level) << 23);       // 2^level = 1/size
    fx = w.tx * __fdividef(1.0f, w.rdx) * is; fy = w.ty * __fdividef(1.0f, w.rdy) * is; fz = w.tz * __fdividef(1.0f, w.rdz) * is;
    if (w.up & 1) fx = 1.0f - fx;
    if (w.up & 2) fy = 1.0f - fy;
    if (w.up & 4) fz = 1.0f - fz;
}

// Start at a point located by index_global(): `pos` in the level-local coordinates of (level, ind).
__device__ __forceinline__ void lw_init(const GridDesc &G, LWalker &w, const vec3 &pos, const vec3 &dir, int level, int ind, float rho) {
    w.level = level; w.cell = G.off[level] + ind; w.rho = rho;
    lw_set_direction(w, dir, pos.x - floorf(pos.x), pos.y - floorf(pos.y), pos.z - floorf(pos.z));
    // coordinates at the cell's level: walk up collecting the child bits
    int cx = 0, cy = 0, cz = 0, i = ind;
    for (int l = level, sh = 0; l > 0; l--, sh++) {
        cx |= (i & 1) << sh; cy |= ((i >> 1) & 1) << sh; cz |= ((i >> 2) & 1) << sh;
        i = G.par[G.off[l] + i - G.nxyz];
    }
    w.cx = ((i % G.nx) << level) | cx; w.cy = (((i / G.nx) % G.ny) << level) | cy; w.cz = ((i / (G.nx * G.ny)) << level) | cz;
}

// One level down through the link in w.rho; the cell was entered through the near face of `ax`.  Written without
// per-axis branches: lanes of a warp descend behind different faces.
__device__ __forceinline__ bool lw_descend(const GridDesc &G, LWalker &w, const int ax) {
    const int base = link_index(w.rho);
    w.level++;
    const float size = lw_size(w.level);
    const float fx = size * w.rdx, fy = size * w.rdy, fz = size * w.rdz;
    const bool ex = ax == 0, ey = ax == 1, ez = ax == 2;
    const bool nx_ = ex || w.tx > fx, ny_ = ey || w.ty > fy, nz_ = ez || w.tz > fz;     // entry on the near side of the axis
    w.tx = ex ? fx : (nx_ ? w.tx - fx : w.tx);
    w.ty = ey ? fy : (ny_ ? w.ty - fy : w.ty);
    w.tz = ez ? fz : (nz_ ? w.tz - fz : w.tz);
    const int near = (nx_ ? 1 : 0) | (ny_ ? 2 : 0) | (nz_ ? 4 : 0);
    const int child = (w.up ^ near) & 7;
    w.cx = (w.cx << 1) | (child & 1); w.cy = (w.cy << 1) | ((child >> 1) & 1); w.cz = (w.cz << 1) | (child >> 2);
    w.cell = G.off[w.level] + base + child;
    w.rho = __ldg(G.dens + w.cell);
    return is_leaf(w.rho);
}

// Cross the face of axis `ax` (e2: the table entry behind it).  Returns true when the cell entered is a leaf (w.rho = density), false when it is
// refined (w.rho = link; call lw_descend until it returns true).  w.cell < 0: the ray has left the cloud.
// `mirror`: MIRROR bit mask, the ray is reflected at such a border and stays in its cell.
// A neighbour that is refined by one level (the common case at a refinement boundary) is resolved right here.
// entry of the neighbour table behind the face of axis `ax` the ray leaves through: {level << 27 | cell, density (or link) of that
// cell} -- the density comes with the look-up, the walk has one dependent gather per crossing instead of two
__device__ __forceinline__ int2 lw_entry(const int *__restrict__ nbr, const LWalker &w, const int ax) {
    return __ldg(reinterpret_cast<const int2 *>(nbr) + 6 * (size_t)w.cell + 2 * ax + ((w.up >> ax) & 1));
}
// e2 = lw_entry(nbr, w, ax), requested by the caller as early as the exit face is known (before the physics of the cell)
__device__ __forceinline__ bool lw_cross(const GridDesc &G, const int2 e2, LWalker &w, const int ax, const int mirror) {
    const int abit = 1 << ax;
    const bool up = (w.up & abit) != 0;
    const int e = e2.x;
    const float rda = (ax == 0) ? w.rdx : ((ax == 1) ? w.rdy : w.rdz);
    if (e < 0) {                                                    // border of the cloud (rare: once per packet)
        if (mirror & ((up ? 2 : 1) << (2 * ax))) {
            w.up ^= abit;
            const float t = lw_size(w.level) * rda;
            if (ax == 0) w.tx = t; else if (ax == 1) w.ty = t; else w.tz = t;
            return true;
        }
        w.cell = -1;
        return true;
    }
    const int nl = e >> SOC_NBR_LEVEL_SHIFT, dl = w.level - nl;
    const float so = lw_size(w.level);
    const int sgn = up ? 1 : -1;
    const bool ex = ax == 0, ey = ax == 1, ez = ax == 2;
    const int nx_ = (w.cx + (ex ? sgn : 0)) >> dl, ny_ = (w.cy + (ey ? sgn : 0)) >> dl, nz_ = (w.cz + (ez ? sgn : 0)) >> dl;
    // the faces of the other two axes move out to those of a coarser cell: whole fine cells in between (0 when dl == 0)
    const int kx = (w.up & 1) ? (((nx_ + 1) << dl) - (w.cx + 1)) : (w.cx - (nx_ << dl));
    const int ky = (w.up & 2) ? (((ny_ + 1) << dl) - (w.cy + 1)) : (w.cy - (ny_ << dl));
    const int kz = (w.up & 4) ? (((nz_ + 1) << dl) - (w.cz + 1)) : (w.cz - (nz_ << dl));
    const float ta = lw_size(nl) * rda;
    w.tx = ex ? ta : fmaf((float)kx * so, w.rdx, w.tx);
    w.ty = ey ? ta : fmaf((float)ky * so, w.rdy, w.ty);
    w.tz = ez ? ta : fmaf((float)kz * so, w.rdz, w.tz);
    w.cx = nx_; w.cy = ny_; w.cz = nz_;
    w.level = nl; w.cell = e & SOC_NBR_INDEX_MASK;
    w.rho = __int_as_float(e2.y);
    return is_leaf(w.rho);
}
