// Spot -> screen-tile binning shared by the SAT render (render.cu, 16-pixel tiles) and the
// Gaussian tensor-core render (gaussian_tc.cu, 128-pixel tiles): footprint + pixel-edge
// arithmetic of overlay_signal_ (/root/reference/src/scopyon/_epifm.py:228-259), a tile
// census, an exclusive scan and a scatter of spot indices (counting sort by tile).
#pragma once
#include "scb_common.cuh"

namespace {


struct __align__(16) SpotRec {
    double ox, oy;      // table origin in camera coordinates: W/2 + x - sw/2   (_epifm.py:233,236)
    double w;           // normalization * unit_area / table scale
    int imin, imax;     // pixel rows [imin, imax) touched on axis 0 (clipped to the image)
    int jmin, jmax;     // pixel cols [jmin, jmax) touched on axis 1
    int slot;           // SAT index, <0: skip
    int pad;
};

struct Geo {
    int n_w, n_h, nti, ntj;
    int tile;           // screen-tile edge in pixels (16: SAT render, 128: Gaussian tensor-core render)
    int side;           // table samples per axis (2*(n_radial-1)+1)
    int n_depth_keys;
    int modulus, blocks, pitch;   // SAT column interleave: (a, b) at a*pitch + (b % modulus)*blocks + b/modulus
    double pl, res, sw, half_w, half_h, depth_cutoff;
    double f0, f1, f2;
};

__device__ __forceinline__ int clamp_to_int(double v, int lo, int hi) {
    if (!(v > (double)lo)) return lo;   // also catches NaN
    if (v > (double)hi) return hi;
    return (int)v;
}

// ceil((i*pl - o)/res) with the reference's first/last clamps (_epifm.py:236-253).
__device__ __forceinline__ int edge_index(int i, int i_first, int i_last, double o, const Geo &g) {
    double v = __ddiv_rn(__dsub_rn(__dmul_rn((double)i, g.pl), o), g.res);
    int e = clamp_to_int(ceil(v), -1, g.side + 1);
    if (i == i_first) e = max(e, 0);
    if (i == i_last) e = min(e, g.side);
    return min(max(e, 0), g.side);  // no-op for interior edges of a valid footprint
}

// One thread per spot: footprint, depth key, tile census.
__global__ void __launch_bounds__(256)
spot_prepare_kernel(Geo g, int64_t n, const double *__restrict__ depth, const double *__restrict__ x,
                    const double *__restrict__ y, const double *__restrict__ weight,
                    const double *__restrict__ inv_scale, const int32_t *__restrict__ slot_of_key,
                    SpotRec *__restrict__ spots, int *__restrict__ tile_count, int32_t *__restrict__ errors) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    SpotRec rec;
    rec.slot = -1;
    rec.imin = rec.imax = rec.jmin = rec.jmax = 0;
    rec.ox = rec.oy = rec.w = 0.0;
    rec.pad = 0;
    const double w = weight[s];
    const double xi = __dsub_rn(x[s], g.f1);
    const double yi = __dsub_rn(y[s], g.f2);
    const double dz = depth ? fabs(__dsub_rn(depth[s], g.f0)) : 0.0;
    if (w > 0.0 && isfinite(xi) && isfinite(yi) && isfinite(dz)) {   // _epifm.py:217-218
        // depth key, _epifm.py:76-84
        int key;
        if (dz < __dadd_rn(g.depth_cutoff, g.res)) {
            key = clamp_to_int(__ddiv_rn(dz, g.res), 0, g.n_depth_keys - 1);
        } else {
            key = g.n_depth_keys;  // frozen at the cutoff ("key -1")
        }
        int slot = slot_of_key ? slot_of_key[key] : 0;
        if (slot < 0) {
            atomicAdd(errors, 1);
        } else {
            // _epifm.py:233-235 and 255-257
            const double cx = __dadd_rn(g.half_w, xi), cy = __dadd_rn(g.half_h, yi);
            const double hs = __dmul_rn(g.sw, 0.5);
            rec.ox = __dsub_rn(cx, hs);
            rec.oy = __dsub_rn(cy, hs);
            int imin = clamp_to_int(floor(__ddiv_rn(rec.ox, g.pl)), -1, g.n_w + 1);
            int imax = clamp_to_int(ceil(__ddiv_rn(__dadd_rn(cx, hs), g.pl)), -1, g.n_w + 1);
            int jmin = clamp_to_int(floor(__ddiv_rn(rec.oy, g.pl)), -1, g.n_h + 1);
            int jmax = clamp_to_int(ceil(__ddiv_rn(__dadd_rn(cy, hs), g.pl)), -1, g.n_h + 1);
            rec.imin = max(0, imin); rec.imax = min(g.n_w, imax);
            rec.jmin = max(0, jmin); rec.jmax = min(g.n_h, jmax);
            if (rec.imax > rec.imin && rec.jmax > rec.jmin) {
                rec.slot = slot;
                rec.w = inv_scale ? w * (g.res * g.res) * inv_scale[slot] : w;
                const int t0 = rec.imin / g.tile, t1 = (rec.imax - 1) / g.tile;
                const int u0 = rec.jmin / g.tile, u1 = (rec.jmax - 1) / g.tile;
                for (int ti = t0; ti <= t1; ++ti)
                    for (int tj = u0; tj <= u1; ++tj) atomicAdd(&tile_count[ti * g.ntj + tj], 1);
            }
        }
    }
    spots[s] = rec;
}

// Table sample index of every pixel edge a footprint touches (_epifm.py:236-253): one thread
// per (spot, axis, edge).  Rows are stored as sample indices, columns as SAT storage offsets
// (column interleave); the render kernels only look them up.
__global__ void __launch_bounds__(256)
spot_edges_kernel(Geo g, int64_t n, const SpotRec *__restrict__ spots, uint16_t *__restrict__ edges, int edge_cap) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per_spot = 2 * edge_cap;
    const int64_t s = idx / per_spot;
    if (s >= n) return;
    const int e = (int)(idx - s * per_spot);
    const int axis = e >= edge_cap, k = e - axis * edge_cap;
    const SpotRec rec = spots[s];
    if (rec.slot < 0) return;
    const int first = axis ? rec.jmin : rec.imin, last = axis ? rec.jmax : rec.imax;
    if (k > last - first) return;
    const int b = edge_index(first + k, first, last, axis ? rec.oy : rec.ox, g);
    edges[s * per_spot + e] = (uint16_t)(axis ? (b % g.modulus) * g.blocks + b / g.modulus : b);
}

// Exclusive scan of the tile census (one CTA; n_tiles is at most a few 10^5).  Each thread
// owns `per` consecutive counters; they are fetched up front (independent loads, 16-byte
// when possible) so the pass costs one memory round trip instead of `per`.
constexpr int kScanPerMax = 32;

__global__ void __launch_bounds__(1024)
tile_scan_kernel(int n_tiles, const int *__restrict__ tile_count, int *__restrict__ tile_start) {
    __shared__ int warp_tot[32];
    const int per = (n_tiles + 1023) / 1024;
    const int b0 = threadIdx.x * per;
    int run = 0;
    int local[kScanPerMax];
    const bool cached = per <= kScanPerMax;
    if (cached) {
        if ((per & 3) == 0) {
            const int4 *src = reinterpret_cast<const int4 *>(tile_count + b0);
#pragma unroll
            for (int i = 0; i < kScanPerMax / 4; ++i) {
                if (4 * i < per) {
                    int4 v = make_int4(0, 0, 0, 0);
                    if (b0 + 4 * i + 3 < n_tiles) v = src[i];
                    else {
                        if (b0 + 4 * i + 0 < n_tiles) v.x = tile_count[b0 + 4 * i + 0];
                        if (b0 + 4 * i + 1 < n_tiles) v.y = tile_count[b0 + 4 * i + 1];
                        if (b0 + 4 * i + 2 < n_tiles) v.z = tile_count[b0 + 4 * i + 2];
                    }
                    local[4 * i + 0] = v.x; local[4 * i + 1] = v.y; local[4 * i + 2] = v.z; local[4 * i + 3] = v.w;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < kScanPerMax; ++i)
                if (i < per) local[i] = (b0 + i < n_tiles) ? tile_count[b0 + i] : 0;
        }
#pragma unroll
        for (int i = 0; i < kScanPerMax; ++i)
            if (i < per) run += local[i];
    } else {
        for (int i = 0; i < per; ++i)
            if (b0 + i < n_tiles) run += tile_count[b0 + i];
    }
    int incl = run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int base = incl - run;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    if (cached) {
#pragma unroll
        for (int i = 0; i < kScanPerMax; ++i) {
            if (i < per && b0 + i < n_tiles) {
                tile_start[b0 + i] = base;
                base += local[i];
            }
        }
    } else {
        for (int i = 0; i < per; ++i) {
            if (b0 + i < n_tiles) {
                tile_start[b0 + i] = base;
                base += tile_count[b0 + i];
            }
        }
    }
    if (threadIdx.x == 1023) tile_start[n_tiles] = base;
}

// Scatter spot indices into their tiles' segments (arrival order; the render kernel
// orders each segment by spot index, so the outcome is deterministic).
__global__ void __launch_bounds__(256)
tile_fill_kernel(Geo g, int64_t n, const SpotRec *__restrict__ spots,
                 const int *__restrict__ tile_start, int *__restrict__ tile_cursor,
                 int *__restrict__ pair_spot) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int slot = spots[s].slot;
    if (slot < 0) return;
    const int imin = spots[s].imin, imax = spots[s].imax, jmin = spots[s].jmin, jmax = spots[s].jmax;
    const int t0 = imin / g.tile, t1 = (imax - 1) / g.tile;
    const int u0 = jmin / g.tile, u1 = (jmax - 1) / g.tile;
    for (int ti = t0; ti <= t1; ++ti)
        for (int tj = u0; tj <= u1; ++tj) {
            const int tile = ti * g.ntj + tj;
            pair_spot[tile_start[tile] + atomicAdd(&tile_cursor[tile], 1)] = (int)s;
        }
}

struct Workspace {
    SpotRec *spots;
    uint16_t *edges;
    int edge_cap;            // edge slots per axis per spot
    int *tile_count, *tile_cursor, *tile_start, *pair_spot;
    size_t bytes;
    int64_t pair_capacity;
};

inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

Geo make_geo(const scb_geometry *geom, int tile) {
    Geo g;
    g.tile = tile;
    g.n_w = geom->n_w; g.n_h = geom->n_h;
    g.nti = (geom->n_w + tile - 1) / tile;
    g.ntj = (geom->n_h + tile - 1) / tile;
    g.side = 2 * (geom->n_radial - 1) + 1;
    g.n_depth_keys = geom->n_depth_keys;
    g.pl = geom->pixel_length; g.res = geom->resolution;
    g.sw = geom->resolution * (double)(g.side - 1);            // _epifm.py:228
    g.half_w = ((double)geom->n_w * geom->pixel_length) * 0.5;  // _epifm.py:230,233
    g.half_h = ((double)geom->n_h * geom->pixel_length) * 0.5;
    g.depth_cutoff = geom->depth_cutoff;
    g.modulus = geom->sat_modulus < 1 ? 1 : geom->sat_modulus;
    g.blocks = (g.side + 1 + g.modulus - 1) / g.modulus;
    g.pitch = g.modulus * g.blocks;
    g.f0 = geom->focal[0]; g.f1 = geom->focal[1]; g.f2 = geom->focal[2];
    return g;
}

// most tiles one spot can touch: footprint rows <= ceil(sw/pl)+1
int64_t max_tiles_per_spot(const Geo &g) {
    double rows = ceil(g.sw / g.pl) + 2.0;
    int64_t per_axis = (int64_t)((rows + g.tile - 2) / g.tile) + 1;
    int64_t a = per_axis < g.nti ? per_axis : g.nti;
    int64_t b = per_axis < g.ntj ? per_axis : g.ntj;
    return a * b;
}

Workspace carve(const Geo &g, int64_t n, void *base) {
    Workspace w;
    const size_t n_tiles = (size_t)g.nti * g.ntj;
    w.pair_capacity = (n > 0 ? n : 1) * max_tiles_per_spot(g);
    char *p = (char *)base;
    size_t off = 0;
    w.spots = (SpotRec *)(p + off); off += align_up((size_t)(n > 0 ? n : 1) * sizeof(SpotRec));
    // pixel edges per axis: footprint rows <= ceil(sw/pl) + 1, plus one closing edge, rounded up to 8
    {
        double rows = ceil(g.sw / g.pl) + 3.0;
        int64_t cap = (int64_t)rows;
        const int64_t most = (g.n_w > g.n_h ? g.n_w : g.n_h) + 1;
        if (cap > most) cap = most;
        w.edge_cap = (int)((cap + 7) & ~(int64_t)7);
    }
    w.edges = (uint16_t *)(p + off); off += align_up((size_t)(n > 0 ? n : 1) * 2 * w.edge_cap * sizeof(uint16_t));
    w.tile_count = (int *)(p + off); off += align_up(n_tiles * sizeof(int));
    w.tile_cursor = (int *)(p + off); off += align_up(n_tiles * sizeof(int));
    w.tile_start = (int *)(p + off); off += align_up((n_tiles + 1) * sizeof(int));
    w.pair_spot = (int *)(p + off); off += align_up((size_t)w.pair_capacity * sizeof(int));
    w.bytes = off;
    return w;
}

int check_geometry(const scb_geometry *geom) {
    SCB_REQUIRE(geom != nullptr, SCB_E_NULL, "geometry is NULL");
    SCB_REQUIRE(geom->n_w > 0 && geom->n_h > 0 && geom->n_w <= 32768 && geom->n_h <= 32768, SCB_E_INVALID,
                "image_size %d x %d out of range", geom->n_w, geom->n_h);
    SCB_REQUIRE(geom->n_radial >= 2 && geom->n_radial <= 2048, SCB_E_INVALID, "n_radial=%d", geom->n_radial);
    SCB_REQUIRE(geom->pixel_length > 0 && geom->resolution > 0, SCB_E_INVALID,
                "pixel_length=%g resolution=%g", geom->pixel_length, geom->resolution);
    SCB_REQUIRE(geom->n_depth_keys >= 1, SCB_E_INVALID, "n_depth_keys=%d", geom->n_depth_keys);
    SCB_REQUIRE(geom->sat_modulus >= 1 && geom->sat_modulus <= 4096, SCB_E_INVALID, "sat_modulus=%d", geom->sat_modulus);
    return 0;
}

}  // namespace
