// Spot -> screen-tile binning shared by the box-table / SAT render (render.cu, 8 x 64-pixel
// strips, one or many frames per call) and the Gaussian tensor-core render (gaussian_tc.cu,
// 128-pixel tiles): footprint + pixel-edge arithmetic of overlay_signal_
// (/root/reference/src/scopyon/_epifm.py:228-259), a tile census, an exclusive scan and a
// scatter of list entries (counting sort by tile).
#pragma once
#include "scb_common.cuh"
#include <cooperative_groups.h>
#include <climits>

namespace {


struct __align__(16) SpotRec {
    double ox, oy;      // table origin in camera coordinates: W/2 + x - sw/2   (_epifm.py:233,236)
    double w;           // normalization * unit_area / table scale
    int imin, imax;     // pixel rows [imin, imax) touched on axis 0 (clipped to the image)
    int jmin, jmax;     // pixel cols [jmin, jmax) touched on axis 1
    int slot;           // SAT index, <0: skip
    // "regular" footprints (written by spot_edges_kernel): the table samples of the pixel edges
    // along an axis share one phase and occupy consecutive slots, edge e at slot0 + e (slot -1 =
    // the zero sample before the table).  phase | (slot0 + 1) << 16, or -1 when irregular.
    int row_run, col_run;
    int walk;           // 1: spot_edges_kernel must walk this footprint's edges (irregular, or no box-table path)
    int frame;          // image of a multi-frame call the spot belongs to
    int pad;
};

struct Geo {
    int n_w, n_h, nti, ntj;
    int frames;         // images rendered by one call (a block of movie frames); strips of frame f follow those of f - 1
    int64_t spots_per_frame;
    int tile_h, tile_w; // screen tile in pixels (8 x 128: SAT render strips, 128 x 128: Gaussian tensor-core render)
    int th_shift, tw_shift, chunk_shift;   // their logarithms (and the chunk's) when they are powers of two, else -1
    int chunk;          // >0: a (spot, tile) pair is listed once per `chunk` columns of the overlap
    int stripes;        // copies of the per-tile counters (power of two): spreads the census atomics
    uint32_t modulus_magic;   // ceil(2^32 / modulus): b / modulus == umulhi(b, magic) for b < 2^16
    int special_edges;  // 1: column edges carry kEdgeZero / last-column codes (SAT render); 0: plain
    int quick_runs;     // 1: spot_prepare may prove an axis regular from its two end edges (box-table path available)
    int side;           // table samples per axis (2*(n_radial-1)+1)
    int n_depth_keys;
    int modulus, slots;           // SAT block layout (scb_common.cuh): phases per axis, slots per phase
    double pl, res, inv_res, sw, half_w, half_h, depth_cutoff;
    double box_peak;    // largest fraction of a spot's photons on one pixel (32-bit accumulators), 1 if unknown
    double f0, f1, f2;
};

__device__ __forceinline__ int clamp_to_int(double v, int lo, int hi) {
    if (!(v > (double)lo)) return lo;   // also catches NaN
    if (v > (double)hi) return hi;
    return (int)v;
}

// ceil((i*pl - o)/res) with the reference's first/last clamps (_epifm.py:236-253).
// The IEEE quotient is only needed when it lies within rounding distance of an integer:
// t * (1/res) is within 2 ulp of t / res, so if no integer lies within 4 ulp of it both have
// the same ceiling; otherwise the correctly rounded division decides.
__device__ __forceinline__ int edge_index(int i, int i_first, int i_last, double o, const Geo &g) {
    const double t = __dsub_rn(__dmul_rn((double)i, g.pl), o);
    const double q = __dmul_rn(t, g.inv_res);
    double c = ceil(q);
    const double slack = __dmul_rn(fabs(q), 8.9e-16);      // 4 * 2^-52
    if (!(c - q > slack && q - (c - 1.0) > slack)) c = ceil(__ddiv_rn(t, g.res));   // also NaN / inf
    int e = clamp_to_int(c, -1, g.side + 1);
    if (i == i_first) e = max(e, 0);
    if (i == i_last) e = min(e, g.side);
    return min(max(e, 0), g.side);  // no-op for interior edges of a valid footprint
}

// List entries of one (spot, tile column) overlap: one, or one per `chunk` columns.
__device__ __forceinline__ int overlap_entries(const Geo &g, int jmin, int jmax, int tj) {
    if (g.chunk <= 0) return 1;
    const int c_lo = max(jmin, tj * g.tile_w), c_hi = min(jmax, (tj + 1) * g.tile_w);
    if (g.chunk_shift >= 0) return (c_hi - c_lo + g.chunk - 1) >> g.chunk_shift;
    return (c_hi - c_lo + g.chunk - 1) / g.chunk;
}
// first / last strip row and column of a footprint: non-negative pixel indices over the tile size
__device__ __forceinline__ int strip_row_of(const Geo &g, int i) { return g.th_shift >= 0 ? i >> g.th_shift : i / g.tile_h; }
__device__ __forceinline__ int strip_col_of(const Geo &g, int j) { return g.tw_shift >= 0 ? j >> g.tw_shift : j / g.tile_w; }

// Unclamped ceil((i*pl - o)/res) when it is safely away from a rounding decision (|distance to
// an integer| > 1e-9 samples, far above the ~1e-12 the evaluation order can move it); false otherwise.
__device__ __forceinline__ bool safe_edge(int i, double o, const Geo &g, int &c) {
    const double q = __dmul_rn(__dsub_rn(__dmul_rn((double)i, g.pl), o), g.inv_res);
    const double up = ceil(q);
    c = (int)up;
    return fabs(q) < 1.0e6 && up - q > 1e-9 && q - (up - 1.0) > 1e-9;
}

// Regularity of one axis from its two end edges.  The exact edge positions are linear in the
// pixel index, (i*pl - o)/res = q0 + k*(pl/res), so if the ceilings of the first and the last
// edge differ by exactly (n-1)*M -- both evaluated safely away from an integer -- every edge in
// between has ceiling c0 + k*M (the fractional parts move monotonically between the two ends).
// Returns the run code (phase | (slot0 + 1) << 16) or -1 when the edges must be walked.
__device__ __forceinline__ int quick_run(int first, int last, double o, const Geo &g) {
    const int n = last - first + 1;                 // edges
    int c0, cl;
    if (n < 2 || !safe_edge(first, o, g, c0) || !safe_edge(last, o, g, cl)) return -1;
    if (cl - c0 != (n - 1) * g.modulus) return -1;
    // interior edges only between an optional opening edge at sample 0 and an optional closing edge at `side`
    int lead = c0;                                  // first edge that lies inside the table
    int slot_shift = 0;
    if (c0 <= 0) {
        lead = c0 + g.modulus;
        slot_shift = 1;
        if (lead <= 0) return -1;
    }
    if (lead >= g.side) return -1;                  // nothing interior
    const int before_last = c0 + (n - 2) * g.modulus;
    if (cl >= g.side && (before_last >= g.side || before_last <= 0)) return -1;   // closing edge needs an interior predecessor
    if (before_last >= g.side) return -1;
    const int slot = g.modulus > 1 ? (int)__umulhi((uint32_t)lead, g.modulus_magic) : lead;
    const int phase = lead - slot * g.modulus;
    const int slot0 = slot - slot_shift;            // slot of edge 0 (-1: the zero sample before the table)
    if (slot0 + n - 1 >= g.slots) return -1;
    return phase | (slot0 + 1) << 16;
}

// Counter copy of a spot: one per group of 256 spots of its frame, round robin.  Counted inside the frame so
// that the threads of a census CTA (one frame, 256 consecutive in-frame indices) always share their copy,
// whatever the number of spots per frame.
constexpr int kCensusThreads = 256;
__device__ __forceinline__ int stripe_of(const Geo &g, int64_t spot, int frame) {
    const int64_t in_frame = g.frames > 1 ? spot - (int64_t)frame * g.spots_per_frame : spot;
    return (int)(in_frame >> 8) & (g.stripes - 1);
}

// The record of spot s (its footprint, depth table, weight, regular runs); true when it touches the image.
// _epifm.py:217-218, 76-84, 233-235, 255-257
__device__ __forceinline__ bool spot_record(const Geo &g, int64_t s, int64_t n_here, int64_t in_frame, int frame, int64_t stride,
                                            const double *__restrict__ depth, const double *__restrict__ x,
                                            const double *__restrict__ y, const double *__restrict__ weight,
                                            const double *__restrict__ inv_scale, const int32_t *__restrict__ slot_of_key,
                                            const int32_t *__restrict__ order, int32_t *__restrict__ errors, SpotRec &rec,
                                            double &w_seen) {
    rec.slot = -1;
    rec.frame = frame;
    rec.pad = 0;
    rec.imin = rec.imax = rec.jmin = rec.jmax = 0;
    rec.ox = rec.oy = rec.w = 0.0;
    rec.walk = 0;
    rec.row_run = rec.col_run = -1;
    w_seen = 0.0;
    bool counted = false;
    if (s < n_here) {
        // the particle this slot shows: the same index of every frame of a block
        const int64_t src = order ? (g.frames > 1 ? (int64_t)frame * g.spots_per_frame + order[in_frame] : order[in_frame]) : s;
        const double w = weight[src];
        const double xi = __dsub_rn(x[src * stride], g.f1);
        const double yi = __dsub_rn(y[src * stride], g.f2);
        const double dz = depth ? fabs(__dsub_rn(depth[src * stride], g.f0)) : 0.0;
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
                    w_seen = w;
                    if (g.quick_runs) {
                        rec.row_run = quick_run(rec.imin, rec.imax, rec.ox, g);
                        rec.col_run = quick_run(rec.jmin, rec.jmax, rec.oy, g);
                    }
                    rec.walk = rec.row_run < 0 || rec.col_run < 0;
                    counted = true;
                }
            }
        }
    }
    return counted;
}

// One thread per spot: footprint, depth key, tile census.
//
// Census.  Every (spot, tile) overlap takes `entries` places in its tile's list; the counter's old
// value is the overlap's position (`ranks`, in the order the fill kernel walks a footprint's tiles), so
// the fill needs no atomics.  The 32 spots of a warp vote before they count: overlaps with the same
// tile (`__match_any_sync`) are added by one lane and share the returned base.  With `order` -- slot
// s of the spot list reads particle order[s], a tile-major ordering the caller refreshes now and then
// (molecules move a pixel or so per frame) -- the spots of a warp are neighbours on the screen and share
// most of their tiles: a tenth of the atomics.  Without it the vote finds no partners and costs little.
constexpr int kLocalStrips = 4096;     // strip counters of a CTA's shared-memory census (a 2048 x 2048 frame has 4096 strips)
constexpr int kLocalRanks = 16;        // strips per footprint it can hold (the rest: global atomics)
#ifndef SCB_PREPARE_CTAS
#define SCB_PREPARE_CTAS 5     // 48 registers, no spills, 40 warps per SM (6: 40 registers with spills, 1 % slower)
#endif
template <int CTAS, bool LOCAL>      // CTAs per SM; LOCAL: the shared-memory census is compiled in
__global__ void __launch_bounds__(256, CTAS)
spot_prepare_kernel(Geo g, int64_t n, int64_t stride, const double *__restrict__ depth, const double *__restrict__ x,
                    const double *__restrict__ y, const double *__restrict__ weight,
                    const double *__restrict__ inv_scale, const int32_t *__restrict__ slot_of_key,
                    SpotRec *__restrict__ spots, int *__restrict__ tile_count,
                    unsigned long long *__restrict__ wmax_bits, int32_t *__restrict__ errors,
                    int *__restrict__ ranks, int rank_cap, const int32_t *__restrict__ order,
                    int *__restrict__ walk_list, unsigned *__restrict__ walk_count) {
    // grid: x over the spots of one frame, y over frames
    const int64_t in_frame = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int frame = blockIdx.y;
    const int64_t s = g.frames > 1 ? (int64_t)frame * g.spots_per_frame + in_frame : in_frame;
    const int64_t n_here = g.frames > 1 ? ((int64_t)(frame + 1) * g.spots_per_frame < n ? (int64_t)(frame + 1) * g.spots_per_frame : n) : n;
    SpotRec rec;
    double w_seen = 0.0;
    const bool counted = spot_record(g, s, n_here, in_frame, frame, stride, depth, x, y, weight, inv_scale, slot_of_key,
                                     order, errors, rec, w_seen);
    if (s < n_here) spots[s] = rec;
    // ---- census, first choice: per CTA in shared memory.  The strips the CTA's 256 spots touch lie in a box of
    // strips (a few dozen when the caller hands the spots over in screen order, a whole 2048 x 2048 frame at
    // most); the overlaps are counted there with shared-memory atomics, whose old values are the positions inside
    // the CTA's share of each list, and one global atomic per (CTA, touched strip) -- all of them in flight at
    // once, nobody waits for a chain of them -- turns the shares into list positions.
    bool census_done = false;
    if (LOCAL && ranks != nullptr && rank_cap <= kLocalRanks) {  // CTA uniform
        __shared__ int s_count[kLocalStrips];
        __shared__ unsigned short s_local[kLocalRanks * kCensusThreads];
        __shared__ int s_box[4];
        const int tid = threadIdx.x;
        if (tid == 0) { s_box[0] = INT_MAX; s_box[1] = -1; s_box[2] = INT_MAX; s_box[3] = -1; }
        __syncthreads();
        const int t0 = rec.imin / g.tile_h, t1 = (rec.imax - 1) / g.tile_h;
        const int u0 = rec.jmin / g.tile_w, u1 = (rec.jmax - 1) / g.tile_w;
        {
            const int lo_i = __reduce_min_sync(0xffffffffu, counted ? t0 : INT_MAX);
            const int hi_i = __reduce_max_sync(0xffffffffu, counted ? t1 : -1);
            const int lo_j = __reduce_min_sync(0xffffffffu, counted ? u0 : INT_MAX);
            const int hi_j = __reduce_max_sync(0xffffffffu, counted ? u1 : -1);
            if ((tid & 31) == 0 && hi_i >= 0) {
                atomicMin(&s_box[0], lo_i); atomicMax(&s_box[1], hi_i);
                atomicMin(&s_box[2], lo_j); atomicMax(&s_box[3], hi_j);
            }
        }
        __syncthreads();
        const int bi0 = s_box[0], bi1 = s_box[1], bj0 = s_box[2], bj1 = s_box[3];
        const int span_j = bj1 - bj0 + 1;
        const int cells = bi1 >= bi0 ? (bi1 - bi0 + 1) * span_j : 0;
        if (cells <= kLocalStrips) {                                // CTA uniform
            census_done = true;
            for (int c = tid; c < cells; c += kCensusThreads) s_count[c] = 0;
            __syncthreads();
            if (counted) {
                int k = 0;
                for (int tj = u0; tj <= u1; ++tj) {                 // the fill kernel's walk: columns of strips, rows inside
                    const int entries = overlap_entries(g, rec.jmin, rec.jmax, tj);
                    if (entries >= 8) atomicAdd(errors, 1);         // cannot happen: scb_render_* refuse such geometries
                    for (int ti = t0; ti <= t1; ++ti, ++k) {
                        const int at = atomicAdd(&s_count[(ti - bi0) * span_j + (tj - bj0)], entries);
                        if (k < rank_cap) s_local[k * kCensusThreads + tid] = (unsigned short)at;
                    }
                }
                if (k > rank_cap) atomicAdd(errors, 1);             // cannot happen: rank_cap bounds the strips of a footprint
            }
            __syncthreads();
            int *count = tile_count + ((size_t)stripe_of(g, s, frame) * g.frames + frame) * g.nti * g.ntj;
            for (int c = tid; c < cells; c += kCensusThreads) {
                const int mine = s_count[c];
                if (mine > 0) {
                    const int ci = c / span_j;
                    s_count[c] = atomicAdd(&count[(bi0 + ci) * g.ntj + bj0 + (c - ci * span_j)], mine);
                }
            }
            __syncthreads();
            if (counted) {
                int *my_rank = ranks + (size_t)s * rank_cap;
                int k = 0;
                for (int tj = u0; tj <= u1; ++tj)
                    for (int ti = t0; ti <= t1 && k < rank_cap; ++ti, ++k)
                        my_rank[k] = s_count[(ti - bi0) * span_j + (tj - bj0)] + (int)s_local[k * kCensusThreads + tid];
            }
        }
    }
    // ---- census, otherwise (no list positions wanted, or strips spread too far): global atomics, warps vote first
    if (!census_done) {
        // the counters exist in `stripes` copies (one per group of 256 spots, round robin) so that the
        // atomics of a frame spread over more L2 sectors; a CTA's spots share their copy
        int *count = tile_count + ((size_t)stripe_of(g, s, frame) * g.frames + frame) * g.nti * g.ntj;
        const int t0 = rec.imin / g.tile_h, t1 = counted ? (rec.imax - 1) / g.tile_h : t0 - 1;
        const int u0 = rec.jmin / g.tile_w, u1 = counted ? (rec.jmax - 1) / g.tile_w : u0 - 1;
        const int n_ti = t1 - t0 + 1;
        const int mine = counted ? n_ti * (u1 - u0 + 1) : 0;
        const int most = __reduce_max_sync(0xffffffffu, mine);
        const unsigned lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
        int *my_rank = ranks ? ranks + (size_t)s * rank_cap : nullptr;
        int ti = t0, tj = u0;                               // the fill kernel's walk: columns of tiles, rows inside
        for (int k = 0; k < most; ++k) {                    // warp uniform
            const bool have = k < mine;
            const int tile = have ? ti * g.ntj + tj : -1 - (int)lane;         // lanes without a tile match nobody
            const int entries = have ? overlap_entries(g, rec.jmin, rec.jmax, tj) : 0;
            const unsigned peers = __match_any_sync(0xffffffffu, tile);
            // entries of the partners below this lane and of all partners (entries < 8: three vote rounds)
            const unsigned b0 = __ballot_sync(0xffffffffu, entries & 1), b1 = __ballot_sync(0xffffffffu, entries & 2),
                           b2 = __ballot_sync(0xffffffffu, entries & 4);
            const int before = __popc(b0 & peers & below) + 2 * __popc(b1 & peers & below) + 4 * __popc(b2 & peers & below);
            const int total = __popc(b0 & peers) + 2 * __popc(b1 & peers) + 4 * __popc(b2 & peers);
            const int leader = __ffs(peers) - 1;
            int base = 0;
            if (have && (int)lane == leader) {
                if (my_rank) base = atomicAdd(&count[tile], total);
                else atomicAdd(&count[tile], total);
            }
            base = __shfl_sync(0xffffffffu, base, leader);
            if (have) {
                if (entries >= 8) atomicAdd(errors, 1);     // cannot happen: scb_render_* refuse such geometries
                if (my_rank) {
                    if (k < rank_cap) my_rank[k] = base + before;
                    else atomicAdd(errors, 1);              // cannot happen: rank_cap bounds the tiles of a footprint
                }
                if (++ti > t1) { ti = t0; ++tj; }
            }
        }
    }
    if (walk_list) {
        // footprints whose edges must be walked (spot_edges_kernel) are listed, one atomic per warp: with box
        // tables they are rare, and the edge kernel then reads a short list instead of every spot record
        const unsigned walks = __ballot_sync(0xffffffffu, counted && rec.walk != 0);
        if (walks) {
            const unsigned lane = threadIdx.x & 31u;
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(walk_count, (unsigned)__popc(walks));
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((walks >> lane) & 1u) walk_list[base + __popc(walks & ((1u << lane) - 1u))] = (int)s;
        }
    }
    if (wmax_bits) {
        // largest weight of the frame: positive doubles order like their bit patterns, so an integer max
        // is exact and order free; one atomic per warp
        unsigned long long bits = (unsigned long long)__double_as_longlong(w_seen);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, d);
            bits = other > bits ? other : bits;
        }
        if ((threadIdx.x & 31) == 0 && bits != 0) atomicMax(wmax_bits + frame, bits);      // one slot per frame of the call
    }
}

// Table sample of every pixel edge a footprint touches (_epifm.py:236-253): one thread per
// (spot, axis) walks the axis' edges in order.  Output per edge: its offset inside a SAT table
// in the block layout -- rows contribute (phase * M * B + slot) * B, columns
// phase * B * B + slot, a corner is table + row + column -- or kEdgeZero for table sample 0
// (S[0][.] == S[.][0] == 0: never fetched).  The clamped closing edge takes the slot after its
// predecessor when that slot repeats the last sample.  While walking, the thread also decides
// whether the axis is "regular" (SpotRec::row_run / col_run: one phase, consecutive slots),
// which lets the render kernel fetch the footprint from the box table as dense rows.  With
// special_edges == 0 the plain sample index is stored instead (Gaussian tensor-core path).
constexpr uint32_t kEdgeZero = 0x80000000u;

__device__ __forceinline__ void spot_edges_of(const Geo &g, int64_t s, int axis, SpotRec *__restrict__ spots,
                                              uint32_t *__restrict__ edges, int edge_cap) {
    const SpotRec rec = spots[s];
    if (rec.slot < 0 || !rec.walk) return;
    const int first = axis ? rec.jmin : rec.imin, last = axis ? rec.jmax : rec.imax;
    const double o = axis ? rec.oy : rec.ox;
    uint32_t *out = edges + (s * 2 + axis) * edge_cap;
    const uint32_t phase_stride = axis ? (uint32_t)(g.slots * g.slots) : (uint32_t)(g.modulus * g.slots) * (uint32_t)g.slots;
    const uint32_t slot_stride = axis ? 1u : (uint32_t)g.slots;
    bool regular = true;
    int before = -1, before_phase = 0, before_slot = 0, run_phase = 0, run_slot0 = 0;
    const int n_edges = last - first + 1;
    for (int k0 = 0; k0 < n_edges; k0 += 4) {          // four edges per 16-byte store (edge_cap is a multiple of 8)
        uint32_t code[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + j;
            code[j] = 0;
            if (k >= n_edges) continue;
            const int b = edge_index(first + k, first, last, o, g);
            int slot = g.modulus > 1 ? (int)__umulhi((uint32_t)b, g.modulus_magic) : b;     // b / modulus
            int phase = b - slot * g.modulus;
            // closing edge: the slot after an interior predecessor repeats the last sample when
            // predecessor + M reaches it
            if (b == g.side && before > 0 && before < g.side && before + g.modulus >= g.side &&
                before_slot + 1 < g.slots) {
                phase = before_phase;
                slot = before_slot + 1;
            }
            if (k == 0) {
                run_phase = phase;
                run_slot0 = b > 0 ? slot : -1;
            } else if (before == 0) {
                // opening edge at sample 0: virtual slot -1 before slot 0, or the real zero slot (0, 0) before (0, 1)
                regular = regular && k == 1 && b > 0 && (slot == 0 || (slot == 1 && phase == 0));
                run_phase = phase;
                run_slot0 = slot - 1;
            } else {
                regular = regular && b > 0 && phase == before_phase && slot == before_slot + 1;
            }
            before = b; before_phase = phase; before_slot = slot;
            code[j] = (uint32_t)b;
            if (g.special_edges) code[j] = b > 0 ? (uint32_t)phase * phase_stride + (uint32_t)slot * slot_stride : kEdgeZero;
        }
        *reinterpret_cast<uint4 *>(out + k0) = make_uint4(code[0], code[1], code[2], code[3]);
    }
    const int run = regular ? (run_phase | (run_slot0 + 1) << 16) : -1;
    if (axis) spots[s].col_run = run; else spots[s].row_run = run;
}

__global__ void __launch_bounds__(128)
spot_edges_kernel(Geo g, int64_t n, SpotRec *__restrict__ spots, uint32_t *__restrict__ edges, int edge_cap,
                  const int *__restrict__ walk_list = nullptr, const unsigned *__restrict__ walk_count = nullptr) {
    // with a list (written by spot_prepare_kernel): a grid-stride loop over its (spot, axis) pairs; without: one
    // thread per (spot, axis) of all spots
    const int64_t pairs = walk_list ? 2 * (int64_t)*walk_count : 2 * n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < pairs; t += (int64_t)gridDim.x * blockDim.x)
        spot_edges_of(g, walk_list ? walk_list[t >> 1] : (t >> 1), (int)(t & 1), spots, edges, edge_cap);
}

// launch shape of spot_edges_kernel
inline void edges_launch_shape(int edge_cap, int64_t n, dim3 &grid, dim3 &block) {
    (void)edge_cap;
    block = dim3(128, 1, 1);
    grid = dim3((unsigned)((2 * n + 127) / 128), 1, 1);
}

// Exclusive scan of the tile census by one thread-block CLUSTER of 8 CTAs.  Logical entry
// j = tile * stripes + stripe is stored at stripe * n_tiles + tile; start[j] gets the scan,
// start[n] the total, so a tile's list is [start[tile * stripes], start[(tile + 1) * stripes]).
// CTA c scans its contiguous share in rounds of 1024 entries (coalesced, one entry per thread),
// publishes its total in shared memory, and after a cluster barrier reads the totals of the
// CTAs before it through distributed shared memory -- one launch, no global round trip.
constexpr int kScanCtas = 8;

// PLAN: the scan of every list's ROOM instead of its length -- a quarter more than the census counted plus 32 units --
// for a following block of frames whose lists are filled while they are counted (render.cu)
template <bool PLAN>
__global__ void __cluster_dims__(kScanCtas, 1, 1) __launch_bounds__(1024)
tile_scan_kernel(int n_tiles, int stripes, const int *__restrict__ tile_count, int *__restrict__ tile_start,
                 int tight = 0) {
    // tight (a test hook): half of what was counted, so that units overflow
    auto room = [tight](int c) { return PLAN ? (tight ? c >> 1 : c + (c >> 2) + 32) : c; };
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ int warp_tot[32];
    __shared__ int cta_total;
    const int n = n_tiles * stripes;
    const int shift = 31 - __clz(stripes), mask = stripes - 1;
    const int rounds = (n + kScanCtas * 1024 - 1) / (kScanCtas * 1024);
    const int share = rounds * 1024;
    const int first = (int)cluster.block_rank() * share;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // all of this thread's entries are fetched first (independent loads): up to kScanRounds rounds
    // live in registers, longer lists fall back to loading round by round
    constexpr int kScanRounds = 8;
    int held[kScanRounds];
#pragma unroll
    for (int r = 0; r < kScanRounds; ++r) {
        const int j = first + r * 1024 + threadIdx.x;
        held[r] = (r < rounds && j < n) ? room(tile_count[(size_t)(j & mask) * n_tiles + (j >> shift)]) : 0;
    }
    // CTA total first (one block reduction), so that every CTA knows its offset before it writes
    int mine = 0;
#pragma unroll
    for (int r = 0; r < kScanRounds; ++r) mine += held[r];
    for (int r = kScanRounds; r < rounds; ++r) {
        const int j = first + r * 1024 + threadIdx.x;
        if (j < n) mine += room(tile_count[(size_t)(j & mask) * n_tiles + (j >> shift)]);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
    if (lane == 0) warp_tot[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 32; ++w) t += warp_tot[w];
        cta_total = t;
    }
    cluster.sync();
    int carry = 0;
    for (unsigned c = 0; c < cluster.block_rank(); ++c) carry += *cluster.map_shared_rank(&cta_total, c);
    if (cluster.block_rank() == kScanCtas - 1 && threadIdx.x == 0) tile_start[n] = carry + cta_total;

    auto scan_round = [&](int r, int v) {
        const int j = first + r * 1024 + threadIdx.x;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        __syncthreads();                      // warp_tot of the previous round (or of the reduction) has been read
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            const int t = warp_tot[w];
            if (w < warp) before += t;
            all += t;
        }
        if (j < n) tile_start[j] = carry + before + incl - v;
        carry += all;
    };
#pragma unroll
    for (int r = 0; r < kScanRounds; ++r)
        if (r < rounds) scan_round(r, held[r]);
    for (int r = kScanRounds; r < rounds; ++r) {
        const int j = first + r * 1024 + threadIdx.x;
        scan_round(r, j < n ? room(tile_count[(size_t)(j & mask) * n_tiles + (j >> shift)]) : 0);
    }
    cluster.sync();      // keep cta_total alive until every CTA has read it
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
    const int t0 = imin / g.tile_h, t1 = (imax - 1) / g.tile_h;
    const int u0 = jmin / g.tile_w, u1 = (jmax - 1) / g.tile_w;
    const int stripe = stripe_of(g, s, 0);
    int *cursor = tile_cursor + (size_t)stripe * g.nti * g.ntj;
    for (int ti = t0; ti <= t1; ++ti)
        for (int tj = u0; tj <= u1; ++tj) {
            const int tile = ti * g.ntj + tj;
            pair_spot[tile_start[tile * g.stripes + stripe] + atomicAdd(&cursor[tile], 1)] = (int)s;
        }
}

constexpr int kOverflowCap = 65536;    // units a planned block render may put aside (more: an error, the caller renders the block again unplanned)

struct Workspace {
    SpotRec *spots;
    uint32_t *edges;
    int edge_cap;            // edge slots per axis per spot
    int *tile_count, *tile_cursor, *tile_start;
    unsigned long long *wmax_bits;   // [frames] bit pattern of the frame's largest spot weight (cleared with the census)
    int *next_tile;                  // dynamic tile queue of the persistent render kernel (cleared likewise)
    int *pair_spot;                  // list entries: spot indices (entry_bytes = 4) or render units
    int *ranks;                      // [n][rank_cap] list positions handed out by the census (render path)
    int rank_cap;
    int *walk_list;                  // spots whose edges must be walked (render path)
    unsigned *walk_count;            // its length (cleared with the census)
    // capacity plan of a block render (render.cu: spot_bin_fused_kernel): where every strip's list starts when the
    // lists are filled in the same pass that counts them -- sized from the previous block's census, kept in the
    // workspace from call to call -- and the units that did not fit their strip's room
    int *plan_base;                  // [n_tiles + 1]
    unsigned *overflow_count;        // cleared with the census
    int *overflow_tile;              // [kOverflowCap]
    void *overflow_units;            // [kOverflowCap] render units
    size_t bytes;
    int64_t pair_capacity;
};

inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

Geo make_geo(const scb_geometry *geom, int tile_h, int tile_w, int chunk = 0, int frames = 1) {
    Geo g;
    g.special_edges = 0;
    g.quick_runs = 0;
    g.tile_h = tile_h; g.tile_w = tile_w; g.chunk = chunk;
    auto log2_of = [](int v) { int k = 0; while (v > 1 && !(v & 1)) { v >>= 1; ++k; } return v == 1 ? k : -1; };
    g.th_shift = log2_of(tile_h); g.tw_shift = log2_of(tile_w); g.chunk_shift = chunk > 0 ? log2_of(chunk) : -1;
    g.n_w = geom->n_w; g.n_h = geom->n_h;
    g.frames = frames < 1 ? 1 : frames; g.spots_per_frame = 0;
    g.box_peak = 1.0;
    g.nti = (geom->n_w + tile_h - 1) / tile_h;
    g.ntj = (geom->n_h + tile_w - 1) / tile_w;
    g.side = 2 * (geom->n_radial - 1) + 1;
    g.n_depth_keys = geom->n_depth_keys;
    g.pl = geom->pixel_length; g.res = geom->resolution;
    g.inv_res = 1.0 / geom->resolution;
    g.sw = geom->resolution * (double)(g.side - 1);            // _epifm.py:228
    g.half_w = ((double)geom->n_w * geom->pixel_length) * 0.5;  // _epifm.py:230,233
    g.half_h = ((double)geom->n_h * geom->pixel_length) * 0.5;
    g.depth_cutoff = geom->depth_cutoff;
    g.modulus = geom->sat_modulus < 1 ? 1 : geom->sat_modulus;
    g.slots = scb_sat_layout(geom->n_radial, g.modulus).slots;
    g.modulus_magic = (uint32_t)((((uint64_t)1 << 32) + g.modulus - 1) / g.modulus);
    // counter copies: as many as keep the scan's register-resident path (<= 32768 entries), at most 8
    g.stripes = 1;
    while (g.stripes < 8 && (int64_t)g.frames * g.nti * g.ntj * g.stripes * 2 <= 32768) g.stripes *= 2;
    g.f0 = geom->focal[0]; g.f1 = geom->focal[1]; g.f2 = geom->focal[2];
    return g;
}

// most list entries one spot can produce: footprint rows/columns <= ceil(sw/pl)+1
int64_t max_tiles_per_spot(const Geo &g) {
    const double rows = ceil(g.sw / g.pl) + 2.0;
    int64_t a = (int64_t)((rows + g.tile_h - 2) / g.tile_h) + 1;
    int64_t b = (int64_t)((rows + g.tile_w - 2) / g.tile_w) + 1;
    if (a > g.nti) a = g.nti;
    if (b > g.ntj) b = g.ntj;
    if (g.chunk > 0) b += (int64_t)(rows / g.chunk) + 1;   // sum over touched tiles of ceil(overlap / chunk)
    return a * b;
}

// tiles one footprint can touch (rows x columns of tiles)
int tiles_per_spot_bound(const Geo &g) {
    const double rows = ceil(g.sw / g.pl) + 2.0;
    int64_t a = (int64_t)((rows + g.tile_h - 2) / g.tile_h) + 1;
    int64_t b = (int64_t)((rows + g.tile_w - 2) / g.tile_w) + 1;
    if (a > g.nti) a = g.nti;
    if (b > g.ntj) b = g.ntj;
    return (int)(a * b);
}

Workspace carve(const Geo &g, int64_t n, void *base, size_t entry_bytes = sizeof(int), bool with_ranks = false) {
    Workspace w;
    const size_t n_tiles = (size_t)g.frames * g.nti * g.ntj;
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
    w.edges = (uint32_t *)(p + off); off += align_up((size_t)(n > 0 ? n : 1) * 2 * w.edge_cap * sizeof(uint32_t));
    w.tile_count = (int *)(p + off); off += align_up(n_tiles * g.stripes * sizeof(int));
    w.tile_cursor = (int *)(p + off); off += align_up(n_tiles * g.stripes * sizeof(int));
    // per frame, so that a frame's accumulator LSBs do not depend on the frames it shares a launch with
    w.wmax_bits = (unsigned long long *)(p + off);
    w.next_tile = (int *)(p + off + 8 * (size_t)g.frames);
    w.walk_count = (unsigned *)(p + off + 8 * (size_t)g.frames + 4);
    w.overflow_count = (unsigned *)(p + off + 8 * (size_t)g.frames + 8); off += align_up(8 * (size_t)g.frames + 16);
    w.tile_start = (int *)(p + off); off += align_up((n_tiles * g.stripes + 1) * sizeof(int));
    w.pair_spot = (int *)(p + off); off += align_up((size_t)w.pair_capacity * entry_bytes);
    w.ranks = nullptr;
    w.rank_cap = 0;
    w.walk_list = nullptr;
    w.plan_base = nullptr;
    w.overflow_tile = nullptr;
    w.overflow_units = nullptr;
    if (with_ranks) {
        w.plan_base = (int *)(p + off); off += align_up((n_tiles + 1) * sizeof(int));
        w.overflow_tile = (int *)(p + off); off += align_up((size_t)kOverflowCap * sizeof(int));
        w.overflow_units = (void *)(p + off); off += align_up((size_t)kOverflowCap * entry_bytes);
        w.rank_cap = tiles_per_spot_bound(g);
        w.ranks = (int *)(p + off); off += align_up((size_t)(n > 0 ? n : 1) * w.rank_cap * sizeof(int));
        w.walk_list = (int *)(p + off); off += align_up((size_t)(n > 0 ? n : 1) * sizeof(int));
    }
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
