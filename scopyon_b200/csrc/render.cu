// Expected-image rendering: spot -> screen-strip binning (counting sort) and warp-per-strip
// accumulation from summed-area-table corners staged by TMA.
//
// Reference: _EPIFMSimulator.get_molecule_plane + PointSpreadingFunction.overlay_signal_
// (/root/reference/src/scopyon/_epifm.py:1262-1264, 224-282).  The reference walks, per
// spot, the ~31x31 pixels under the 1999x1999-sample PSF table and sums a ~66x66 slice of
// the table for each.  Here
//   * the pixel-edge -> table-index arithmetic is evaluated with the same IEEE fp64
//     operations in the same order (explicit __d*_rn intrinsics: never contracted to FMA);
//   * each slice sum is S[i1][j1] - S[i0][j1] - S[i1][j0] + S[i0][j0] on the int64 SAT
//     built by scb_psf_sat_build -- exact, so results do not depend on how a footprint is cut;
//   * one WARP owns one strip of the image -- 8 x 128 pixels with 32-bit accumulators (fp32 box
//     tables, the default) or 8 x 64 with 64-bit ones (fp64 tables, exact mode) -- and keeps its
//     accumulators in 4 KB of shared memory (28 strips in flight per SM).  The strip's work list holds
//     one 32-byte unit per (spot, <= 8 rows, <= 32 columns) overlap; the census of spot_prepare hands
//     every (spot, strip) its place in the list, so the lists are filled without atomics.  For a
//     footprint whose pixel edges are evenly spaced (whole number of table samples per pixel) the box
//     sums of all its pixels are one dense rectangle of one block of the "box table" (psf.cu), so a
//     unit's (<= 8) x 32 values arrive by TMA: one bulk copy of up to 1 KB (2 KB in fp64) into a
//     three-stage shared-memory ring, completion on an mbarrier, two units ahead of the arithmetic.
//     Lane l owns column l of the unit and adds `box * weight` to its accumulators: one shared-memory
//     load, one multiply and one conversion per pixel.  Every lane of a unit carries a footprint pixel,
//     there is no block-wide barrier and no atomic on the image.  Footprints whose edges are not evenly
//     spaced (pixel pitch not a whole number of samples, or a rounding step in the edge arithmetic)
//     gather the four SAT corners per pixel through per-edge offsets -- same integer box sums,
//     bit-identical result;
//   * accumulators are fixed point: 64-bit with the LSB 2^-K photons chosen per call from the largest
//     spot weight (exact mode), 32-bit with the LSB chosen per strip from its list length (fp32 mode),
//     in both cases so that no sum can overflow.  Integer addition is associative, so the image is
//     bitwise reproducible whatever order the work list was filled in.
#include "binning.cuh"

#include <cuda.h>      // CUtensorMap (types only: the encoder is looked up at run time)
#include <string.h>
#include <type_traits>
#include <atomic>
#include <mutex>

#ifndef SCB_FILL_CTAS
#define SCB_FILL_CTAS 3        // 80 registers, 24 warps per SM: the ten list positions a footprint needs are all in flight at once
#endif

namespace {

// Accumulators and strip shape.  A strip's accumulators fill 4 KB of shared memory: 64-bit fixed point
// (fp64 box tables, exact mode: LSB chosen per call) or 32-bit fixed point (fp32 tables, fp32 frames: LSB
// chosen per strip from its list length).  ROWS x kCols pixels: 8 x 64 (exact mode), and for fp32 either
// 8 x 128 or 16 x 64 -- taller strips cut a ~31 x 31 footprint into fewer units (4.3 instead of 7.0 per
// spot: fewer census atomics, fewer unit records, less per-unit bookkeeping in the render) at the price
// of 2 KB ring stages.
template <typename BoxT> struct AccOf;
template <> struct AccOf<double> { using type = long long; };
template <> struct AccOf<float> { using type = int; };
template <typename BoxT, int ROWS> struct Mode {
    using Acc = typename AccOf<BoxT>::type;
    static constexpr int kCols = 4096 / (ROWS * (int)sizeof(Acc));
};
constexpr int kUnitCols = 32;         // columns per unit = lanes
constexpr int kMaxWarps = 7;          // warps (= strips in flight) per CTA
// units fetched per round (one per lane): 16, or 8 with 16-row strips, whose larger ring then still lets
// 21 warps share an SM's shared memory
template <int ROWS> constexpr int batch_units() { return ROWS == 16 ? 8 : 16; }
constexpr int kFastSlots = 32;        // widest box-table block row the ring holds
constexpr int kDefaultRows = 8;       // fp32 strip shape and copy engine unless the environment says otherwise
constexpr int kDefaultCopy = 0;
constexpr int kStages = 3;            // ring depth: the copy of unit u + 2 is issued while unit u is consumed
// COPY 3 = a scene without box tables (every footprint gathers SAT corners): no ring, hence more warps per SM
constexpr int kCopyNone = 3;
constexpr int kStageTail = 16;        // bytes behind a ring stage: its mbarrier (the stage's address gives both)
template <typename BoxT, int ROWS, int COPY = 0> constexpr size_t warp_smem_bytes() {
    return 4096 + (COPY == kCopyNone ? 0 : kStages * (ROWS * kFastSlots * sizeof(BoxT) + kStageTail)) + batch_units<ROWS>() * 32 + 32;
}
// CTAs per SM from the shared memory one warp needs (accumulators + ring + unit batch)
template <typename BoxT, int ROWS, int COPY = 0> constexpr int ctas_per_sm() {
    constexpr int fit = (int)(232448 / (kMaxWarps * warp_smem_bytes<BoxT, ROWS, COPY>() + 1024));
    return fit > 4 ? 4 : fit;
}
constexpr uint32_t kUnitFast = 0x80000000u;

// One work-list entry: the overlap of a spot with (<= 8 rows) x (<= 32 columns) of a strip.
struct __align__(16) Unit {
    double ws;             // weight * res^2 / table scale (the render scales it to accumulator LSBs)
    const void *src;       // fast: first box-table row of the unit (`rows` rows of `slots` doubles);  gather: the spot's SAT
    uint32_t erow, ecol;   // gather: first row / column edge of the overlap inside `edges`
    uint32_t shape;        // rows | cols << 8 | first strip row << 16 | first strip column << 24
    uint32_t extra;        // fast: kUnitFast | box-table column of lane 0
};

// Fixed-point exponent of the accumulators: every pixel is below 2 * n_spots * max weight
// (a spot spreads at most ~1.00001 of its weight over its footprint), so with
// LSB = 2^-K, K = 62 - ceil(log2(bound)), a pixel sum stays below 2^62.
__device__ __forceinline__ int accumulator_shift(unsigned long long wmax_bits, int64_t n_spots) {
    const double wmax = __longlong_as_double((long long)wmax_bits);
    if (!(wmax > 0.0)) return 0;
    int e;
    frexp(2.0 * (double)n_spots * wmax, &e);
    return max(-900, min(62 - e, 900));
}

// 32-bit accumulators: a strip's pixels stay below (units in its list) * (largest weight) *
// (largest fraction of a spot's photons one pixel can receive, box_peak); the factor 1.25 covers
// the slightly wider pixels of unevenly spaced footprints and the half-LSB rounding of every
// term, so with K = 31 - ceil(log2(bound)) every sum stays below 2^31 / 1.25.  The LSB is then
// about (bound / actual maximum) * 2^-31 of the strip's brightest pixel: ~1e-7 of it for evenly
// lit strips, a few 1e-6 where one strip holds both a dense cluster and faint background.
__device__ __forceinline__ int strip_shift(int n_units, unsigned long long wmax_bits, double box_peak) {
    const double wmax = __longlong_as_double((long long)wmax_bits);
    if (!(wmax > 0.0) || n_units <= 0) return 0;
    int e;
    frexp(1.25 * (double)n_units * wmax * box_peak, &e);
    return max(-900, min(31 - e, 900));
}

// One thread per spot: write the spot's units into the strips' list segments, at the places the
// census handed out (`ranks`, in the same tile order) -- no atomics here.
template <int CTAS, int kFillAhead>
__global__ void __launch_bounds__(256, CTAS)
strip_fill_kernel(Geo g, int64_t n, const SpotRec *__restrict__ spots, int edge_cap,
                  const int *__restrict__ ranks, int rank_cap,
                  const int64_t *__restrict__ sat, const void *__restrict__ box_table, int box_bytes,
                  const int *__restrict__ tile_start, const unsigned long long *__restrict__ wmax_bits,
                  Unit *__restrict__ units) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const SpotRec rec = spots[s];
    if (rec.slot < 0) return;
    Unit u;
    u.ws = rec.w;
    const size_t table_at = (size_t)rec.slot * ((size_t)g.modulus * g.modulus * g.slots * g.slots);
    const bool fast = g.quick_runs && rec.row_run >= 0 && rec.col_run >= 0;   // quick_runs: a usable box table exists
    // edge e of a regular axis sits at slot slot0 + e (>= -1); the pixel between edges e and e + 1
    // is box-table row / column slot0 + e + 1
    const int row_slot0 = (rec.row_run >> 16) - 1, col_slot0 = (rec.col_run >> 16) - 1;
    const char *block = static_cast<const char *>(box_table) +
                        (table_at + ((size_t)(rec.row_run & 0xffff) * g.modulus + (rec.col_run & 0xffff)) *
                                        (size_t)(g.slots * g.slots)) * (size_t)box_bytes;
    const uint32_t ebase = (uint32_t)s * 2u * (uint32_t)edge_cap;
    const int stripe = stripe_of(g, s, rec.frame);
    const int frame_tile0 = rec.frame * g.nti * g.ntj;     // first strip of the spot's frame
    const int *my_rank = ranks + (size_t)s * rank_cap;
    const int t0 = rec.imin / g.tile_h, t1 = (rec.imax - 1) / g.tile_h;
    const int u0 = rec.jmin / g.tile_w, u1 = (rec.jmax - 1) / g.tile_w;
    const int n_ti = t1 - t0 + 1, n_strips = n_ti * (u1 - u0 + 1);
    // the list positions of the footprint's first kFillAhead strips (all of them for ~31-row footprints) are
    // requested before the first unit is written: one round of load latency per spot instead of one per strip
    int pos[kFillAhead > 0 ? kFillAhead : 1];
    if (kFillAhead > 0) {
        int ci = 0, cj = 0;
#pragma unroll
        for (int k = 0; k < kFillAhead; ++k) {
            pos[k] = 0;
            if (k < n_strips) {
                const int tile = frame_tile0 + (t0 + ci) * g.ntj + (u0 + cj);
                pos[k] = tile_start[tile * g.stripes + stripe] + __ldg(my_rank + k);
                if (++ci == n_ti) { ci = 0; ++cj; }
            }
        }
    }
    auto emit = [&](int ti, int tj, Unit *dst) {
        const int c_lo = max(rec.jmin, tj * g.tile_w), c_hi = min(rec.jmax, (tj + 1) * g.tile_w);
        const int entries = (c_hi - c_lo + g.chunk - 1) / g.chunk;
        const int r_lo = max(rec.imin, ti * g.tile_h), r_hi = min(rec.imax, (ti + 1) * g.tile_h);
        const int rows = r_hi - r_lo;
        u.erow = ebase + (uint32_t)(r_lo - rec.imin);
        const int first_box_row = row_slot0 + (r_lo - rec.imin) + 1;
        for (int q = 0; q < entries; ++q) {
            const int c = c_lo + q * g.chunk;
            u.ecol = ebase + (uint32_t)(edge_cap + c - rec.jmin);
            u.shape = (uint32_t)rows | (uint32_t)min(g.chunk, c_hi - c) << 8 |
                      (uint32_t)(r_lo - ti * g.tile_h) << 16 | (uint32_t)(c - tj * g.tile_w) << 24;
            if (fast) {
                u.src = block + (size_t)first_box_row * g.slots * (size_t)box_bytes;
                u.extra = kUnitFast | (uint32_t)(col_slot0 + (c - rec.jmin) + 1);
            } else {
                u.src = sat + table_at;
                u.extra = 0;
            }
            dst[q] = u;
        }
    };
    // the census' walk: columns of strips, rows inside
    int ci = 0, cj = 0;
#pragma unroll
    for (int k = 0; k < kFillAhead; ++k) {
        if (k < n_strips) {
            emit(t0 + ci, u0 + cj, units + pos[k]);
            if (++ci == n_ti) { ci = 0; ++cj; }
        }
    }
    for (int k = kFillAhead; k < n_strips; ++k) {
        const int tile = frame_tile0 + (t0 + ci) * g.ntj + (u0 + cj);
        emit(t0 + ci, u0 + cj, units + tile_start[tile * g.stripes + stripe] + __ldg(my_rank + k));
        if (++ci == n_ti) { ci = 0; ++cj; }
    }
}

// ---- binning of a block of frames in ONE pass ------------------------------------------------------------------
// When the strips' list regions are laid out beforehand -- `plan_base`, the scan of every list's room, sized a quarter
// above what the previous block's census counted (tile_scan_kernel<true>; molecules move about a pixel per frame) --
// the census, the hand-out of list positions and the writing of the units need no scan between them: this kernel is
// spot_prepare_kernel and strip_fill_kernel in one, without the spot records (kept only for footprints whose edges
// must be walked), the rank table and the second pass over the spots.  A CTA counts its 256 spots' overlaps in shared
// memory, claims its share of every touched strip's list with one global atomic (the counter's old value is the share's
// offset) and its threads write their units there.  A unit beyond its strip's room -- the plan is a forecast -- goes to
// a short overflow list the render kernel also reads; the image does not depend on where a unit was listed.
constexpr int kFusedStrips = 1024;     // strip counters of a CTA (more: its spots claim their places one by one)
constexpr int kFusedKeep = 12;         // strips of a footprint whose places inside the CTA's shares are kept (three packed words)
struct ListPlan {
    const int *base;                   // [n_tiles + 1] start of every strip's list region; NULL: lists are tile_start's
    const int *count;                  // [n_tiles] units counted per strip (the census)
    int capacity;                      // units the list buffer holds (clamped to INT_MAX: list positions are 32-bit)
    unsigned *overflow_count;
    int *overflow_tile;
    Unit *overflow_units;
};

// the units of one spot: what strip_fill_kernel writes, for a footprint whose record is still in registers
struct UnitMaker {
    Unit u;
    bool fast;
    int row_slot0, col_slot0;
    const char *block;
    const int64_t *table;
    uint32_t ebase;
    __device__ __forceinline__ UnitMaker(const Geo &g, const SpotRec &rec, int64_t s, int edge_cap, const int64_t *sat,
                                         const void *box_table, int box_bytes) {
        u.ws = rec.w;
        const size_t table_at = (size_t)rec.slot * ((size_t)g.modulus * g.modulus * g.slots * g.slots);
        fast = g.quick_runs && rec.row_run >= 0 && rec.col_run >= 0;
        row_slot0 = (rec.row_run >> 16) - 1;
        col_slot0 = (rec.col_run >> 16) - 1;
        block = static_cast<const char *>(box_table) +
                (table_at + ((size_t)(rec.row_run & 0xffff) * g.modulus + (rec.col_run & 0xffff)) * (size_t)(g.slots * g.slots)) *
                    (size_t)box_bytes;
        table = sat + table_at;
        ebase = (uint32_t)s * 2u * (uint32_t)edge_cap;
    }
    // chunk q of the overlap with strip (ti, tj)
    __device__ __forceinline__ const Unit &make(const Geo &g, const SpotRec &rec, int ti, int tj, int q, int edge_cap, int box_bytes) {
        const int c_lo = max(rec.jmin, tj * g.tile_w), c_hi = min(rec.jmax, (tj + 1) * g.tile_w);
        const int r_lo = max(rec.imin, ti * g.tile_h), r_hi = min(rec.imax, (ti + 1) * g.tile_h);
        const int c = c_lo + q * g.chunk;
        u.erow = ebase + (uint32_t)(r_lo - rec.imin);
        u.ecol = ebase + (uint32_t)(edge_cap + c - rec.jmin);
        u.shape = (uint32_t)(r_hi - r_lo) | (uint32_t)min(g.chunk, c_hi - c) << 8 | (uint32_t)(r_lo - ti * g.tile_h) << 16 |
                  (uint32_t)(c - tj * g.tile_w) << 24;
        if (fast) {
            u.src = block + (size_t)(row_slot0 + (r_lo - rec.imin) + 1) * g.slots * (size_t)box_bytes;
            u.extra = kUnitFast | (uint32_t)(col_slot0 + (c - rec.jmin) + 1);
        } else {
            u.src = table;
            u.extra = 0;
        }
        return u;
    }
};

template <int CTAS>
__global__ void __launch_bounds__(256, CTAS)
spot_bin_fused_kernel(Geo g, int64_t n, int64_t stride, const double *__restrict__ depth, const double *__restrict__ x,
                      const double *__restrict__ y, const double *__restrict__ weight,
                      const double *__restrict__ inv_scale, const int32_t *__restrict__ slot_of_key,
                      SpotRec *__restrict__ spots, int *__restrict__ tile_count,
                      unsigned long long *__restrict__ wmax_bits, int32_t *__restrict__ errors,
                      const int32_t *__restrict__ order, int *__restrict__ walk_list, unsigned *__restrict__ walk_count,
                      int edge_cap, const int64_t *__restrict__ sat, const void *__restrict__ box_table, int box_bytes,
                      Unit *__restrict__ units, ListPlan plan) {
    __shared__ int s_count[kFusedStrips];          // overlaps counted per strip of the CTA's box, then the share's offset
    __shared__ int s_base[kFusedStrips];           // start of the strip's list
    __shared__ int s_room[kFusedStrips];           // units it holds
    __shared__ int s_box[4];
    const int tid = threadIdx.x;
    const int64_t in_frame = (int64_t)blockIdx.x * blockDim.x + tid;
    const int frame = blockIdx.y;
    const int64_t s = (int64_t)frame * g.spots_per_frame + in_frame;
    const int64_t n_here = (int64_t)(frame + 1) * g.spots_per_frame < n ? (int64_t)(frame + 1) * g.spots_per_frame : n;
    SpotRec rec;
    double w_seen = 0.0;
    const bool counted = spot_record(g, s, n_here, in_frame, frame, stride, depth, x, y, weight, inv_scale, slot_of_key, order,
                                     errors, rec, w_seen);
    if (counted && rec.walk) spots[s] = rec;       // read (and completed) by spot_edges_kernel only

    if (tid == 0) { s_box[0] = INT_MAX; s_box[1] = -1; s_box[2] = INT_MAX; s_box[3] = -1; }
    __syncthreads();
    const int t0 = strip_row_of(g, rec.imin), t1 = strip_row_of(g, max(rec.imax - 1, 0));
    const int u0 = strip_col_of(g, rec.jmin), u1 = strip_col_of(g, max(rec.jmax - 1, 0));
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
    const bool local = cells <= kFusedStrips;                          // CTA uniform
    const int frame_tile0 = frame * g.nti * g.ntj;
    int *count = tile_count + frame_tile0;
    UnitMaker maker(g, rec, s, edge_cap, sat, box_table, box_bytes);
    // a unit's place: `at` in the list of `tile` (which starts at `base` and holds `room`), or the overflow list
    auto put = [&](int tile, int base, int room, int at, const Unit &u) {
        if (at < room && at < plan.capacity - base) {
            units[base + at] = u;
        } else {
            const unsigned o = atomicAdd(plan.overflow_count, 1u);
            if (o < (unsigned)kOverflowCap) {
                plan.overflow_tile[o] = tile;
                plan.overflow_units[o] = u;
            } else {
                atomicAdd(errors, 1);
            }
        }
    };
    if (local) {
        for (int c = tid; c < cells; c += kCensusThreads) s_count[c] = 0;
        __syncthreads();
        // pass 1: count; the counter's old value is the overlap's place inside the CTA's share (kept packed, 16 bits
        // each, in registers for the first kFusedKeep strips of the footprint; the rest claim their places directly)
        unsigned long long keep0 = 0, keep1 = 0, keep2 = 0;
        if (counted) {
            int k = 0;
            for (int tj = u0; tj <= u1; ++tj) {
                const int entries = overlap_entries(g, rec.jmin, rec.jmax, tj);
                for (int ti = t0; ti <= t1; ++ti, ++k) {
                    if (k < kFusedKeep) {
                        const unsigned long long at = (unsigned)atomicAdd(&s_count[(ti - bi0) * span_j + (tj - bj0)], entries);
                        if (k < 4) keep0 |= at << (16 * k);
                        else if (k < 8) keep1 |= at << (16 * (k - 4));
                        else keep2 |= at << (16 * (k - 8));
                    }
                }
            }
        }
        __syncthreads();
        for (int c = tid; c < cells; c += kCensusThreads) {
            const int mine = s_count[c];
            if (mine > 0) {
                const int ci = c / span_j;
                const int tile = (bi0 + ci) * g.ntj + bj0 + (c - ci * span_j);
                s_count[c] = atomicAdd(&count[tile], mine);
                const int b = plan.base[frame_tile0 + tile];
                s_base[c] = b;
                s_room[c] = plan.base[frame_tile0 + tile + 1] - b;
            }
        }
        __syncthreads();
        if (counted) {
            int k = 0;
            for (int tj = u0; tj <= u1; ++tj) {
                const int entries = overlap_entries(g, rec.jmin, rec.jmax, tj);
                for (int ti = t0; ti <= t1; ++ti, ++k) {
                    const int tile = ti * g.ntj + tj;
                    int base, room, at;
                    if (k < kFusedKeep) {
                        const int c = (ti - bi0) * span_j + (tj - bj0);
                        const unsigned long long word = k < 4 ? keep0 >> (16 * k) : (k < 8 ? keep1 >> (16 * (k - 4)) : keep2 >> (16 * (k - 8)));
                        base = s_base[c];
                        room = s_room[c];
                        at = s_count[c] + (int)(word & 0xffffu);
                    } else {
                        base = plan.base[frame_tile0 + tile];
                        room = plan.base[frame_tile0 + tile + 1] - base;
                        at = atomicAdd(&count[tile], entries);
                    }
                    for (int q = 0; q < entries; ++q)
                        put(frame_tile0 + tile, base, room, at + q, maker.make(g, rec, ti, tj, q, edge_cap, box_bytes));
                }
            }
        }
    } else if (counted) {
        // the CTA's strips spread too far for its counters: every overlap claims its place with its own atomic
        for (int tj = u0; tj <= u1; ++tj) {
            const int entries = overlap_entries(g, rec.jmin, rec.jmax, tj);
            for (int ti = t0; ti <= t1; ++ti) {
                const int tile = ti * g.ntj + tj;
                const int base = plan.base[frame_tile0 + tile];
                const int room = plan.base[frame_tile0 + tile + 1] - base;
                const int at = atomicAdd(&count[tile], entries);
                for (int q = 0; q < entries; ++q)
                    put(frame_tile0 + tile, base, room, at + q, maker.make(g, rec, ti, tj, q, edge_cap, box_bytes));
            }
        }
    }
    // footprints whose edges must be walked are listed (as in spot_prepare_kernel)
    {
        const unsigned walks = __ballot_sync(0xffffffffu, counted && rec.walk != 0);
        if (walks) {
            const unsigned lane = tid & 31u;
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(walk_count, (unsigned)__popc(walks));
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((walks >> lane) & 1u) walk_list[base + __popc(walks & ((1u << lane) - 1u))] = (int)s;
        }
    }
    {
        unsigned long long bits = (unsigned long long)__double_as_longlong(w_seen);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, d);
            bits = other > bits ? other : bits;
        }
        if ((tid & 31) == 0 && bits != 0) atomicMax(wmax_bits + frame, bits);
    }
}

// ---- mbarrier / TMA helpers ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_row(void *dst, const void *src, uint32_t bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// Fast unit, producer side: one bulk copy of the unit's box-table rows (contiguous in the block),
// issued by the lane that fetched the unit (it still holds source and size in registers).
__device__ __forceinline__ void unit_stage(const void *src, uint32_t bytes, void *stage, void *bar) {
    mbar_expect_tx(bar, bytes);
    tma_row(stage, src, bytes, bar);
}

// the same by shared-memory address (the barrier sits right behind its stage)
__device__ __forceinline__ void unit_stage_at(const void *src, uint32_t bytes, uint32_t stage_s, uint32_t bar_s) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(stage_s),
                 "l"(src), "r"(bytes), "r"(bar_s)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_at(uint32_t bar_s, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar_s), "r"(parity) : "memory");
}

// box * weight -> accumulator LSBs, in the precision of the box table
__device__ __forceinline__ long long to_fixed(double box, double ws) { return __double2ll_rn(__dmul_rn(box, ws)); }
__device__ __forceinline__ int to_fixed(float box, float ws) { return __float2int_rn(__fmul_rn(box, ws)); }

#include "render_reg.cuh"
#include "render_tile.cuh"

// Per-lane cp.async (LDGSTS) staging: the alternative to the TMA bulk copy.  A unit's rows are one
// contiguous run of the block (<= 2 KB), so lane l copies the 16-byte pieces l, l + 32, ... of it: no
// elected lane, no uniform-register traffic, no mbarrier -- completion is this thread's cp.async group.
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Accumulator update of four consecutive rows of a unit, shared by both paths.  Rows beyond the unit's
// last carry stale values: their products are computed and dropped (only the update is predicated), which
// keeps the body branch free.  _epifm.py:280-282 (`if photons > 0` needs no branch: adding zero changes nothing)
template <typename Acc, typename BoxT, int COLS>
__device__ __forceinline__ void rows4_add(Acc *a, int rows_left, BoxT ws, const BoxT (&box)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const Acc q = to_fixed(box[k], ws);
        if (k < rows_left) a[k * COLS] += q;
    }
}

// Build-time variant, measured and off: the accumulator update as a shared-memory reduction (red.shared.add, SASS
// ATOMS.ADD) -- 4 instead of 6 instructions per pixel row, no row or lane predicates, same bits -- runs the C4 block
// at 4.95 instead of 2.32 ms: the shared-memory pipe retires an ATOMS far slower than a load plus a store.
#ifndef SCB_ACC_RED
#define SCB_ACC_RED 0
#endif
// N rows of a unit into 32-bit accumulators by red.shared: all box values first, then multiply / convert / reduce
template <int N, int COLS>
__device__ __forceinline__ void rows_red(uint32_t a_s, const float *st, int slots, float w) {
    float box[N];
#pragma unroll
    for (int k = 0; k < N; ++k) box[k] = st[k * slots];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const int q = to_fixed(box[k], w);
        asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a_s + (uint32_t)(k * COLS * 4)), "r"(q) : "memory");
    }
}

template <typename BoxT, int ROWS, int ROUNDS>
__device__ __forceinline__ void rounds_add(typename Mode<BoxT, ROWS>::Acc *a, const BoxT *st, int slots, int rows, BoxT ws) {
    using M = Mode<BoxT, ROWS>;
    BoxT box[ROUNDS][4];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) box[r][k] = st[(r * 4 + k) * slots];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r)
        rows4_add<typename M::Acc, BoxT, M::kCols>(a + r * 4 * M::kCols, rows - r * 4, ws, box[r]);
}

// Fast unit, consumer side: lane l reads its column of the staged box rows, four rows per round
// (a unit has 1..ROWS rows; the round count is warp uniform).
template <typename BoxT, int ROWS, int SLOTS>
__device__ __forceinline__ void unit_accumulate_fast(const Unit *meta, int u, int lane,
                                                     typename Mode<BoxT, ROWS>::Acc *acc, const BoxT *stage,
                                                     int runtime_slots, double scale) {
    using M = Mode<BoxT, ROWS>;
    const int slots = SLOTS ? SLOTS : runtime_slots;
    BoxT ws;
    int n_rows, n_cols, acc_at, box_col;
    if constexpr (sizeof(BoxT) == 4) {
        // fp32 mode: the fetching lane has unpacked the unit (see the fetch below): one 16-byte load
        const uint4 packed = *reinterpret_cast<const uint4 *>(meta + u);
        ws = __uint_as_float(packed.x);
        acc_at = (int)packed.y;
        box_col = (int)packed.z;
        n_rows = (int)(packed.w & 0xffu);
        n_cols = (int)(packed.w >> 8);
    } else {
        ws = (BoxT)(meta[u].ws * scale);
        const uint32_t shape = meta[u].shape;
        n_rows = shape & 0xff;
        n_cols = (shape >> 8) & 0xff;
        acc_at = ((shape >> 16) & 0xff) * M::kCols + (shape >> 24);
        box_col = (int)(meta[u].extra & 0xffu);
    }
    if constexpr (SCB_ACC_RED && sizeof(BoxT) == 4 && ROWS == 8) {
        // Accumulator update by shared-memory reduction: one instruction and one pass through the shared-memory
        // pipe per pixel instead of load / add / store, no row predicates (the code exists once per row count,
        // warp uniform) and no lane predicate (idle lanes repeat the unit's last column with weight zero).
        const int col = min(lane, n_cols - 1);
        const float w = lane < n_cols ? ws : 0.0f;
        const uint32_t a_s = smem_addr(acc + acc_at + col);
        const BoxT *st = stage + min(box_col + col, slots - 1);
        switch (n_rows) {
        case 1: rows_red<1, M::kCols>(a_s, st, slots, w); break;
        case 2: rows_red<2, M::kCols>(a_s, st, slots, w); break;
        case 3: rows_red<3, M::kCols>(a_s, st, slots, w); break;
        case 4: rows_red<4, M::kCols>(a_s, st, slots, w); break;
        case 5: rows_red<5, M::kCols>(a_s, st, slots, w); break;
        case 6: rows_red<6, M::kCols>(a_s, st, slots, w); break;
        case 7: rows_red<7, M::kCols>(a_s, st, slots, w); break;
        default: rows_red<8, M::kCols>(a_s, st, slots, w); break;
        }
        return;
    }
    int rows = lane < n_cols ? n_rows : 0;   // idle lanes: no rows
    asm volatile("" : "+r"(rows));       // keep it one value: one compare per row below instead of two
    typename M::Acc *a = acc + acc_at + lane;
    const BoxT *st = stage + min(box_col + lane, slots - 1);   // idle lanes stay inside the row
    // rounds of four rows, fully unrolled per round count (warp uniform): all loads of a unit are in flight
    // before the first conversion.  (A third path without predicates for whole units -- 8 rows, 32 columns, two
    // units in five -- was measured: 1.6 % slower, the extra branch costs more than the eight compares.)
    const int rounds = (n_rows + 3) >> 2;
    if (rounds <= 1) {
        rounds_add<BoxT, ROWS, 1>(a, st, slots, rows, ws);
    } else if (rounds == 2 || ROWS == 8) {
        rounds_add<BoxT, ROWS, 2>(a, st, slots, rows, ws);
    } else if (rounds == 3) {
        rounds_add<BoxT, ROWS, 3>(a, st, slots, rows, ws);
    } else {
        rounds_add<BoxT, ROWS, 4>(a, st, slots, rows, ws);
    }
}

// Gather unit: per-edge table offsets from `edges`, corners straight from global memory.  Neighbouring
// lanes share a column edge (the right corners of lane l are the left corners of lane l + 1), so a lane
// fetches its left corners only and takes the right ones from its neighbour by shuffle; the last active
// lane fetches both.  The kernel is latency bound on this path (ncu at 66.39 nm pixels: 24 long-scoreboard stall
// cycles per issued instruction, DRAM a third busy), so the loads are arranged in two dependent levels per unit
// whatever its height: all edge offsets first (column edges by every lane, the unit's <= ROWS + 1 row edges by the
// first lanes), then every corner of the unit at once.
struct GatherEdges {
    uint32_t left, right, my_row;
};
// level 1 of a gather unit: its edge offsets (column edges by every lane, row edge `lane` by the first lanes)
__device__ __forceinline__ GatherEdges gather_edges(const Unit *meta, int u, int lane, const uint32_t *__restrict__ edges) {
    const uint32_t shape = meta[u].shape;
    const int n_rows = shape & 0xff, n_cols = (shape >> 8) & 0xff;
    const uint32_t c = meta[u].ecol + (uint32_t)min(lane, n_cols - 1);   // idle lanes repeat the last column
    GatherEdges e;
    e.left = __ldg(edges + c);
    e.right = __ldg(edges + c + 1);
    e.my_row = __ldg(edges + meta[u].erow + (uint32_t)min(lane, n_rows));
    return e;
}

template <typename BoxT, int ROWS>
__device__ __forceinline__ void unit_accumulate_gather(const Unit *meta, int u, int lane,
                                                       typename Mode<BoxT, ROWS>::Acc *acc, const GatherEdges e,
                                                       double scale) {
    using M = Mode<BoxT, ROWS>;
    BoxT ws;
    if constexpr (sizeof(BoxT) == 4) ws = *reinterpret_cast<const float *>(&meta[u].ws);   // scaled at fetch time
    else ws = (BoxT)(meta[u].ws * scale);
    const uint32_t shape = meta[u].shape;
    const int n_rows = shape & 0xff, n_cols = (shape >> 8) & 0xff;
    const int rows = lane < n_cols ? n_rows : 0;
    typename M::Acc *a = acc + ((shape >> 16) & 0xff) * M::kCols + (shape >> 24) + lane;
    const long long *table = static_cast<const long long *>(meta[u].src);
    const uint32_t left = e.left, right = e.right;
    const bool last_lane = lane == n_cols - 1 || lane == 31;
    // level 2: the corners on every row edge of the unit (edges past the unit's last repeat it: their rows are dropped)
    long long L[ROWS + 1], own_right[ROWS + 1];
#pragma unroll
    for (int k = 0; k <= ROWS; ++k) {
        const uint32_t row = __shfl_sync(0xffffffffu, e.my_row, k);
        L[k] = 0;
        own_right[k] = 0;
        if (k <= n_rows) {                                     // warp uniform
            if (!((row | left) & kEdgeZero)) L[k] = __ldg(table + (row + left));
            if (last_lane && !((row | right) & kEdgeZero)) own_right[k] = __ldg(table + (row + right));
        }
    }
    long long before = 0;
#pragma unroll
    for (int k = 0; k <= ROWS; ++k) {
        const long long from_neighbour = __shfl_down_sync(0xffffffffu, L[k], 1);
        const long long here = (last_lane ? own_right[k] : from_neighbour) - L[k];
        if (k > 0) {
            // >= 0: the table is non-negative, edges are monotone; rounded as the box table is
            const auto q = to_fixed((BoxT)(here - before), ws);
            if (k - 1 < rows) a[(k - 1) * M::kCols] += q;
        }
        before = here;
    }
}

static_assert(ctas_per_sm<double, 8>() * (kMaxWarps * warp_smem_bytes<double, 8>() + 1024) <= 232448 &&
                  ctas_per_sm<float, 8>() * (kMaxWarps * warp_smem_bytes<float, 8>() + 1024) <= 232448 &&
                  ctas_per_sm<float, 16>() * (kMaxWarps * warp_smem_bytes<float, 16>() + 1024) <= 232448,
              "the CTAs of one SM must fit in shared memory");

// COPY: 0 = TMA bulk copy per unit (issued by the lane that fetched the unit, completion on an mbarrier),
//       1 = per-lane cp.async pieces (completion by cp.async group).
template <typename OutT, typename BoxT, int ROWS, int SLOTS, int COPY>
__global__ void __launch_bounds__(kMaxWarps * 32, ctas_per_sm<BoxT, ROWS, COPY>())
render_strips_kernel(Geo g, const Unit *__restrict__ units, const uint32_t *__restrict__ edges,
                     const int *__restrict__ tile_start, int *__restrict__ next_tile,
                     const unsigned long long *__restrict__ wmax_bits, int64_t n_spots,
                     OutT *__restrict__ out, int accumulate, ListPlan plan) {
    using M = Mode<BoxT, ROWS>;
    using Acc = typename M::Acc;
    constexpr int kStripCols = M::kCols;
    constexpr int kStageEntries = ROWS * kFastSlots;
    constexpr size_t kAccBytes = ROWS * kStripCols * sizeof(Acc);
    constexpr int kBatch = batch_units<ROWS>();
    static_assert(kAccBytes == 4096, "accumulators of one strip");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per-warp carve: accumulators | copy ring | unit batch | mbarriers
    unsigned char *mine = smem_raw + warp * warp_smem_bytes<BoxT, ROWS, COPY>();
    Acc *acc = reinterpret_cast<Acc *>(mine);
    BoxT *ring = reinterpret_cast<BoxT *>(mine + kAccBytes);
    // a stage = its box rows, then 16 bytes holding its mbarrier: stage and barrier share one address register
    constexpr int kStageStride = kStageEntries + kStageTail / (int)sizeof(BoxT);          // in table entries
    constexpr uint32_t kStageBytes = kStageEntries * sizeof(BoxT), kStrideBytes = kStageBytes + kStageTail;
    constexpr size_t kRingBytes = COPY == kCopyNone ? 0 : kStages * (size_t)kStrideBytes;
    Unit *meta = reinterpret_cast<Unit *>(mine + kAccBytes + kRingBytes);
    auto stage_of = [&](int st) { return ring + st * kStageStride; };
    auto bar_of = [&](int st) { return reinterpret_cast<unsigned long long *>(ring + st * kStageStride + kStageEntries); };
    const uint32_t ring_s = smem_addr(ring);

    const int frame_tiles = g.nti * g.ntj, n_tiles = g.frames * frame_tiles;
    const int slots = SLOTS ? SLOTS : g.slots;
    const uint32_t row_bytes = (uint32_t)slots * (uint32_t)sizeof(BoxT);
    const int64_t frame_spots = g.frames > 1 ? g.spots_per_frame : n_spots;

    for (int i = lane; i < ROWS * kStripCols; i += 32) acc[i] = 0;
    if (COPY != kCopyNone) {
        for (int i = lane; i < kStages * kStageStride; i += 32) ring[i] = (BoxT)0;   // rows past a unit's last are read (and dropped)
        __syncwarp();                                                                // ... before the barriers inside are set up
    }
    if (COPY != 1 && COPY != kCopyNone) {
        if (lane == 0) {
            for (int st = 0; st < kStages; ++st) mbar_init(bar_of(st), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // zeroed ring visible to the async proxy
    }
    __syncwarp();
    int p_stage = 0, c_stage = 0;            // ring positions of the producer and the consumer (fast units only)
    uint32_t c_parity = 0;                   // COPY 0: every stage completes one mbarrier phase per turn of the ring
    // COPY 0 walks the ring by address: c_at = shared address of the consumer's stage (the producer's destinations
    // are worked out per lane when a batch is fetched: the ring is drained at every batch boundary, so the batch's
    // k-th box-table unit lands k stages behind c_at)
    uint32_t c_at = ring_s;
    uint32_t phases = 0;                     // COPY 2: bit s = phase of stage s's mbarrier (only its TMA units advance it)

    for (;;) {
        int tile = 0;
        if (lane == 0) tile = atomicAdd(next_tile, 1);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= n_tiles) break;
        const int frame = tile / frame_tiles, in_frame = tile - frame * frame_tiles;
        const int ti = in_frame / g.ntj, tj = in_frame - ti * g.ntj;
        const int row0 = ti * ROWS, col0 = tj * kStripCols;
        OutT *image = out + (size_t)frame * g.n_w * g.n_h;
        // the strip's list: a segment of the scanned lists, or -- planned block renders -- the filled part of the
        // region the plan gave it, followed by the units of the overflow list that belong to it
        int seg_begin, seg_end, n_units;
        if (plan.base) {
            seg_begin = plan.base[tile];
            const int room = min(plan.base[tile + 1], plan.capacity) - seg_begin;
            n_units = plan.count[tile];
            seg_end = seg_begin + max(0, min(n_units, room));
        } else {
            seg_begin = tile_start[tile * g.stripes];
            seg_end = tile_start[(tile + 1) * g.stripes];
            n_units = seg_end - seg_begin;
        }
        // LSB per strip (32-bit accumulators) or per frame (64-bit), from the frame's own largest weight: a frame's
        // bits do not depend on the frames it shares the launch with
        const unsigned long long wmax = wmax_bits[frame];
        const int shift = sizeof(Acc) == 4 ? strip_shift(n_units, wmax, g.box_peak)
                                           : accumulator_shift(wmax, frame_spots);
        const double scale = scalbn(1.0, shift), lsb = scalbn(1.0, -shift);

        const Unit *list = units;
        unsigned overflow_at = 0;                  // overflow entries looked at so far
        int missing = n_units - (seg_end - seg_begin);     // units of this strip that sit in the overflow list
        for (;;) {
        for (int base = seg_begin; base < seg_end; base += kBatch) {
            const int nb = min(kBatch, seg_end - base);
            __syncwarp();
            // lane u fetches unit u of the batch, publishes it in shared memory and keeps what its
            // TMA copy needs
            const void *my_src = nullptr;
            uint32_t my_bytes = 0;
            bool my_fast = false;
            if (lane < nb) {
                const uint4 *src = reinterpret_cast<const uint4 *>(list + base + lane);
                uint4 *dst = reinterpret_cast<uint4 *>(meta + lane);
                uint4 head = __ldg(src);                                    // {ws, src}
                uint4 tail = __ldg(src + 1);                                // {erow, ecol, shape, extra}
                my_src = reinterpret_cast<const void *>(((unsigned long long)head.w << 32) | head.z);
                my_bytes = (tail.z & 0xffu) * row_bytes;
                my_fast = (tail.w & kUnitFast) != 0;
                // (an L2 prefetch of the unit's rows here, up to a batch before its copy is issued, was measured: 4.5 %
                // slower -- the kernel has no issue slots to spare and does not wait for memory)
                if constexpr (sizeof(BoxT) == 4) {
                    // fp32 mode: the fetching lane forms the unit's weight in accumulator LSBs once (the same
                    // double product and conversion the 32 consumer lanes would each repeat) and, for a
                    // box-table unit, unpacks shape and column into the words the consumers use directly:
                    // {weight, accumulator offset, box-table column, rows | columns << 8}, source pointer behind
                    const double ws = __longlong_as_double(((long long)head.y << 32) | head.x);
                    head.x = __float_as_uint((float)(ws * scale));
                    if (my_fast) {
                        tail.x = head.z;                                    // source pointer (edge indices are unused)
                        tail.y = head.w;
                        head.y = ((tail.z >> 16) & 0xffu) * M::kCols + (tail.z >> 24);
                        head.z = tail.w & 0xffu;
                        head.w = tail.z & 0xffffu;
                    }
                }
                dst[0] = head;
                dst[1] = tail;
            }
            const uint32_t fast_mask = COPY == kCopyNone ? 0u : __ballot_sync(0xffffffffu, my_fast);
            __syncwarp();
            uint32_t my_dst = 0;             // COPY 0: shared address of the stage this lane's unit is copied into
            if constexpr (COPY == 0) {
                uint32_t turn = (uint32_t)__popc(fast_mask & ((1u << lane) - 1u)) + (c_at - ring_s) / kStrideBytes;
                turn -= (turn * 43u >> 7) * 3u;                            // mod 3 (turn < 35)
                static_assert(kStages == 3, "ring arithmetic");
                my_dst = ring_s + turn * kStrideBytes;
            }

            // COPY 2 splits the units between the two engines -- even units of a batch by TMA bulk copy, odd ones by
            // per-lane cp.async -- because neither is free: an SM retires one bulk copy per ~40-50 cycles whatever its
            // size (tools/probes/tma_rate_probe.cu) and its issue costs ~37 instructions, while cp.async pieces cost
            // shared-memory wavefronts the accumulators need too.
            auto by_lanes = [&](int u) { return COPY == 1 || (COPY == 2 && (u & 1)); };
            // The batch loop exists twice: a batch of box-table units only (the rule: every unit of a C4 strip) does
            // not test each unit's kind, twice per unit
            auto run_batch = [&](auto all_fast) {
            auto is_fast = [&](int u) { return decltype(all_fast)::value || ((fast_mask >> u) & 1u) != 0; };
            auto stage_unit = [&](int u) {
                if (is_fast(u)) {                          // warp uniform
                    if (!by_lanes(u)) {
                        if constexpr (COPY == 0) {
                            if (lane == u) unit_stage_at(my_src, my_bytes, my_dst, my_dst + kStageBytes);
                        } else if (lane == u) {
                            // the stage was last written by another lane's cp.async pieces (generic proxy)
                            if (COPY == 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            unit_stage(my_src, my_bytes, stage_of(p_stage), bar_of(p_stage));
                        }
                    } else {
                        const char *src;
                        uint32_t bytes;
                        if constexpr (sizeof(BoxT) == 4) {     // the fetch moved the pointer behind the unpacked words
                            const uint2 where = *reinterpret_cast<const uint2 *>(&meta[u].erow);
                            src = reinterpret_cast<const char *>(((unsigned long long)where.y << 32) | where.x);
                        } else {
                            src = static_cast<const char *>(meta[u].src);
                        }
                        bytes = (meta[u].shape & 0xffu) * row_bytes;
                        char *dst = reinterpret_cast<char *>(stage_of(p_stage));
                        for (uint32_t off = lane * 16u; off < bytes; off += 512u) cp_async16(dst + off, src + off);
                    }
                    if (COPY != 0 && ++p_stage == kStages) p_stage = 0;
                }
                if (COPY != 0 && COPY != kCopyNone) cp_async_commit();          // one group per unit, empty for TMA and gather units
            };
            for (int u = 0; u < min(nb, kStages - 1); ++u) stage_unit(u);
            for (int u = 0; u < nb; ++u) {
                // the previous unit's accumulator updates (other lanes, other columns) are visible, and every
                // lane has finished reading the ring stage the next copy overwrites (the one unit u - 1 used)
                __syncwarp();
                if (u + kStages - 1 < nb) stage_unit(u + kStages - 1);
                if (is_fast(u)) {
                    if (!by_lanes(u)) {
                        if (COPY == 2) {
                            mbar_wait(bar_of(c_stage), (phases >> c_stage) & 1u);
                            phases ^= 1u << c_stage;
                        } else {
                            mbar_wait_at(c_at + kStageBytes, c_parity);
                        }
                    } else {
                        // groups committed after unit u's: min(nb - 1 - u, 2)
                        const int ahead = min(nb - 1 - u, kStages - 1);
                        if (ahead >= 2) cp_async_wait<2>();
                        else if (ahead == 1) cp_async_wait<1>();
                        else cp_async_wait<0>();
                        __syncwarp();                      // the other lanes' pieces have landed too
                    }
                    if constexpr (COPY == 0) {
                        unit_accumulate_fast<BoxT, ROWS, SLOTS>(meta, u, lane, acc,
                                                                ring + (c_at - ring_s) / (uint32_t)sizeof(BoxT), slots, scale);
                        c_at += kStrideBytes;
                        if (c_at == ring_s + kStages * kStrideBytes) { c_at = ring_s; c_parity ^= 1u; }
                    } else {
                        unit_accumulate_fast<BoxT, ROWS, SLOTS>(meta, u, lane, acc, stage_of(c_stage), slots, scale);
                        if (++c_stage == kStages) { c_stage = 0; c_parity ^= 1u; }
                    }
                } else {
                    // (requesting the next gather unit's edge offsets before this unit's corners was measured: no
                    // gain at 66.39 nm, and the registers it holds cost the box-table path 2 %)
                    unit_accumulate_gather<BoxT, ROWS>(meta, u, lane, acc, gather_edges(meta, u, lane, edges), scale);
                }
            }
            };
            const uint32_t whole = nb >= 32 ? 0xffffffffu : (1u << nb) - 1u;
            if (COPY != kCopyNone && fast_mask == whole) run_batch(std::true_type());
            else run_batch(std::false_type());
            if (COPY != 0 && COPY != kCopyNone) cp_async_wait<0>();     // TMA and gather units at the end of a batch leave empty groups behind
        }
        // the next unit of the overflow list that belongs to this strip, if any (the list is empty unless the plan
        // fell short somewhere)
        if (!plan.base || missing <= 0) break;     // (a strip whose plan held looks at nothing: only the strips that
        --missing;                                 // overflowed search the list, and stop at their last unit)
        const unsigned n_over = min(*plan.overflow_count, (unsigned)kOverflowCap);
        int found = -1;
        while (overflow_at < n_over && found < 0) {
            const unsigned i = overflow_at + lane;
            const unsigned hits = __ballot_sync(0xffffffffu, i < n_over && plan.overflow_tile[i] == tile);
            if (hits) found = (int)overflow_at + __ffs(hits) - 1;
            overflow_at = hits ? (unsigned)found + 1u : overflow_at + 32u;
        }
        if (found < 0) break;
        list = plan.overflow_units;
        seg_begin = found;
        seg_end = found + 1;
        }

        // ---- write the strip (coalesced rows) and clear the accumulators for the next one
        __syncwarp();
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int i = row0 + r;
#pragma unroll
            for (int q = 0; q < kStripCols / 32; ++q) {
                const int cc = q * 32 + lane, j = col0 + cc;
                const double v = (double)acc[r * kStripCols + cc] * lsb;
                acc[r * kStripCols + cc] = 0;
                if (i < g.n_w && j < g.n_h) {
                    const size_t o = (size_t)i * g.n_h + j;
                    if (accumulate) image[o] = (OutT)((double)image[o] + v);
                    else image[o] = (OutT)v;
                }
            }
        }
    }
}

// ---- measurement hook -----------------------------------------------------------
struct ProfilePool {
    std::mutex mu;         // begin / end / the launches that record may come from different threads
    cudaEvent_t *start = nullptr, *stop = nullptr;
    int capacity = 0, used = 0;
    bool enabled = false;
} g_profile;

}  // namespace

extern "C" int scb_profile_begin(int max_launches) {
    SCB_REQUIRE(max_launches > 0 && max_launches <= (1 << 20), SCB_E_INVALID, "scb_profile_begin: max_launches=%d", max_launches);
    std::lock_guard<std::mutex> lock(g_profile.mu);
    if (g_profile.capacity < max_launches) {
        for (int i = 0; i < g_profile.capacity; ++i) { cudaEventDestroy(g_profile.start[i]); cudaEventDestroy(g_profile.stop[i]); }
        free(g_profile.start); free(g_profile.stop);
        g_profile.start = (cudaEvent_t *)malloc(sizeof(cudaEvent_t) * max_launches);
        g_profile.stop = (cudaEvent_t *)malloc(sizeof(cudaEvent_t) * max_launches);
        for (int i = 0; i < max_launches; ++i) { SCB_CUDA(cudaEventCreate(&g_profile.start[i])); SCB_CUDA(cudaEventCreate(&g_profile.stop[i])); }
        g_profile.capacity = max_launches;
    }
    g_profile.used = 0;
    g_profile.enabled = true;
    return 0;
}

extern "C" int scb_profile_end(double *total_ms, int64_t *launches) {
    SCB_REQUIRE(total_ms && launches, SCB_E_NULL, "scb_profile_end: NULL pointer");
    std::lock_guard<std::mutex> lock(g_profile.mu);
    g_profile.enabled = false;
    double sum = 0.0;
    for (int i = 0; i < g_profile.used; ++i) {
        float ms = 0.f;
        SCB_CUDA(cudaEventSynchronize(g_profile.stop[i]));
        SCB_CUDA(cudaEventElapsedTime(&ms, g_profile.start[i], g_profile.stop[i]));
        sum += ms;
    }
    *total_ms = sum;
    *launches = g_profile.used;
    g_profile.used = 0;
    return 0;
}

// Measured variants of the fp32 render, chosen once per process from the environment:
//   SCB_RENDER_ROWS = 8 | 16   strip shape 8 x 128 or 16 x 64 pixels
//   SCB_RENDER_COPY = tma | ldgsts   one TMA bulk copy per unit, or per-lane cp.async pieces
//   SCB_RENDER_PATH = smem | tensor | ldg
//       smem (default): render_strips_kernel, accumulators in shared memory, one TMA bulk copy per unit;
//       tensor / ldg: the register-accumulator kernels of render_reg.cuh (one tensor-map TMA copy per unit / plain
//       loads into registers).  All three give the same bits; round 2 measured 1.26 / 1.33 / 1.50 ms per 16-frame
//       C4 block (profiles/render_variants_r2.md).
struct RenderVariant {
    int rows, copy, reg;
};
static RenderVariant render_variant() {      // read at every call (a few getenv): tests flip the variables between calls
    RenderVariant r = {kDefaultRows, kDefaultCopy, 0};
    if (const char *e = getenv("SCB_RENDER_ROWS")) r.rows = atoi(e) == 16 ? 16 : (atoi(e) == 8 ? 8 : r.rows);
    if (const char *e = getenv("SCB_RENDER_COPY")) r.copy = e[0] == 'l' ? 1 : (e[0] == 'm' ? 2 : (e[0] == 't' ? 0 : r.copy));
    if (const char *e = getenv("SCB_RENDER_PATH"))
        r.reg = !strcmp(e, "tile") ? 3 : (e[0] == 't' ? 2 : (e[0] == 'l' ? 1 : 0));
    if (r.rows != 8 || r.copy != 0) r.reg = 0;        // strip shape and copy engine belong to the shared-memory kernel
    if (r.rows != 8 && r.copy == 2) r.copy = 0;
    return r;
}

static int forced_gather() {
    const char *force = getenv("SCB_RENDER_FORCE_GATHER");
    return force ? (force[0] == '1' ? 1 : (force[0] == '2' ? 2 : 0)) : 0;
}

// The register kernel takes every fp32 box table whose rows a tensor map can describe (strides in multiples of 16 bytes)
static bool reg_path_for(bool have_box, int box_bytes, int slots) {
    if (!(have_box && box_bytes == 4 && slots % 4 == 0 && slots <= 1000)) return false;
    // asked for, or the only kernels that can: block rows wider than the shared-memory kernel's ring (small pixels,
    // e.g. 50 nm: 48 slots) would otherwise send every footprint down the SAT-corner gather
    return render_variant().reg != 0 || slots > kFastSlots;
}

static Geo strip_geo(const scb_geometry *geom, bool have_box, int box_bytes, int frames = 1, int64_t spots_per_frame = 0) {
    const int rows = box_bytes == 4 ? render_variant().rows : 8;
    const int cols = box_bytes == 4 ? (rows == 16 ? Mode<float, 16>::kCols : Mode<float, 8>::kCols) : Mode<double, 8>::kCols;
    Geo g = make_geo(geom, rows, cols, kUnitCols, frames);
    g.spots_per_frame = spots_per_frame;
    g.box_peak = geom->box_peak > 0.0 && geom->box_peak <= 1.0 ? geom->box_peak : 1.0;
    g.special_edges = 1;
    // with a box table whose block rows the TMA ring can hold (and copy: multiples of 16 bytes),
    // evenly spaced footprints never read their edges
    g.quick_runs = reg_path_for(have_box, box_bytes, g.slots) ||
                   (have_box && g.slots <= kFastSlots && (g.slots * box_bytes) % 16 == 0);
    // test hook: keep the table precision and accumulators but send every footprint down the SAT-corner path
    // ('1': of the shared-memory kernel, '2': of the kernel the tables would select)
    if (forced_gather()) g.quick_runs = 0;
    return g;
}

extern "C" size_t scb_render_workspace_bytes(const scb_geometry *geom, int64_t n_spots) {
    if (check_geometry(geom) != 0 || n_spots < 0) return 0;
    const size_t exact = carve(strip_geo(geom, false, 8), n_spots, nullptr, sizeof(Unit), true).bytes;
    const size_t fp32 = carve(strip_geo(geom, false, 4), n_spots, nullptr, sizeof(Unit), true).bytes;
    return exact > fp32 ? exact : fp32;
}

extern "C" size_t scb_render_frames_workspace_bytes(const scb_geometry *geom, int64_t n_per_frame, int n_frames) {
    if (check_geometry(geom) != 0 || n_per_frame < 0 || n_frames < 1) return 0;
    const size_t exact = carve(strip_geo(geom, false, 8, n_frames, n_per_frame), n_per_frame * n_frames, nullptr,
                               sizeof(Unit), true).bytes;
    const size_t fp32 = carve(strip_geo(geom, false, 4, n_frames, n_per_frame), n_per_frame * n_frames, nullptr,
                              sizeof(Unit), true).bytes;
    return exact > fp32 ? exact : fp32;
}

template <typename OutT, typename BoxT, int ROWS, int SLOTS, int COPY>
static int launch_render_as(const Geo &g, const Workspace &w, int64_t n_spots, OutT *out, int accumulate,
                            cudaStream_t s, const ListPlan &plan) {
    const int n_tiles = g.frames * g.nti * g.ntj;
    // persistent grid: a few CTAs per SM, each warp pulls strips from a queue; small images get
    // narrower CTAs so that the strips still spread over all SMs
    const int slots = ctas_per_sm<BoxT, ROWS, COPY>() * SCB_SM_COUNT;
    int warps = (n_tiles + slots - 1) / slots;
    warps = warps < 1 ? 1 : (warps > kMaxWarps ? kMaxWarps : warps);
    int ctas = (n_tiles + warps - 1) / warps;
    if (ctas > slots) ctas = slots;
    const size_t smem = (size_t)warps * warp_smem_bytes<BoxT, ROWS, COPY>();
    // the attribute is per device: remembered per template instance and device ordinal
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    SCB_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        SCB_CUDA(cudaFuncSetAttribute(render_strips_kernel<OutT, BoxT, ROWS, SLOTS, COPY>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(kMaxWarps * warp_smem_bytes<BoxT, ROWS, COPY>())));
        configured.fetch_or(bit, std::memory_order_release);
    }
    render_strips_kernel<OutT, BoxT, ROWS, SLOTS, COPY><<<ctas, warps * 32, smem, s>>>(
        g, (const Unit *)w.pair_spot, w.edges, w.tile_start, w.next_tile, w.wmax_bits, n_spots, out, accumulate, plan);
    return 0;
}

template <typename OutT, typename BoxT, int ROWS, int COPY>
static int launch_render_slots(const Geo &g, const Workspace &w, int64_t n_spots, OutT *out, int accumulate,
                               cudaStream_t s, const ListPlan &plan) {
    return g.slots == 32 ? launch_render_as<OutT, BoxT, ROWS, 32, COPY>(g, w, n_spots, out, accumulate, s, plan)
                         : launch_render_as<OutT, BoxT, ROWS, 0, COPY>(g, w, n_spots, out, accumulate, s, plan);
}

template <typename OutT, int WARPS, int CTAS, int STAGES>
static int launch_render_reg_as(const CUtensorMap &box_map, const Geo &g, const Workspace &w, const int64_t *sat, OutT *out,
                                int accumulate, cudaStream_t s) {
    const int n_tiles = g.frames * g.nti * g.ntj;
    // persistent grid: CTAS CTAs per SM, each warp pulls strips from a queue; small images get narrower CTAs
    const int slots = CTAS * SCB_SM_COUNT;
    int warps = (n_tiles + slots - 1) / slots;
    warps = warps < 1 ? 1 : (warps > WARPS ? WARPS : warps);
    int ctas = (n_tiles + warps - 1) / warps;
    if (ctas > slots) ctas = slots;
    const size_t smem = (size_t)warps * reg_warp_smem<STAGES>();
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    SCB_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        SCB_CUDA(cudaFuncSetAttribute(render_strips_reg_kernel<OutT, WARPS, CTAS, STAGES>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(WARPS * reg_warp_smem<STAGES>())));
        configured.fetch_or(bit, std::memory_order_release);
    }
    render_strips_reg_kernel<OutT, WARPS, CTAS, STAGES><<<ctas, warps * 32, smem, s>>>(
        box_map, g, (const RUnit *)w.pair_spot, w.spots, w.edges, w.edge_cap, sat, w.tile_start, w.next_tile, w.wmax_bits,
        out, accumulate);
    return 0;
}

template <typename OutT, int WARPS, int CTAS, int SLOTS, int DIST>
static int launch_render_ldg_as(const Geo &g, const Workspace &w, const int64_t *sat, const float *box, OutT *out,
                                int accumulate, cudaStream_t s) {
    const int n_tiles = g.frames * g.nti * g.ntj;
    const int slots = CTAS * SCB_SM_COUNT;
    int warps = (n_tiles + slots - 1) / slots;
    warps = warps < 1 ? 1 : (warps > WARPS ? WARPS : warps);
    int ctas = (n_tiles + warps - 1) / warps;
    if (ctas > slots) ctas = slots;
    render_strips_ldg_kernel<OutT, WARPS, CTAS, SLOTS, DIST><<<ctas, warps * 32, (size_t)warps * kLdgWarpSmem, s>>>(
        g, box, (const RUnit *)w.pair_spot, w.spots, w.edges, w.edge_cap, sat, w.tile_start, w.next_tile, w.wmax_bits, out,
        accumulate);
    return 0;
}

template <typename OutT, int SLOTS, int CTAS>
static int launch_render_tiles_as(const Geo &g, const Workspace &w, const int64_t *sat, const float *box, OutT *out,
                                  int accumulate, cudaStream_t s) {
    const int n_tiles = g.frames * g.nti * g.ntj;
    int ctas = CTAS * SCB_SM_COUNT;
    if (ctas > n_tiles) ctas = n_tiles;
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    SCB_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        SCB_CUDA(cudaFuncSetAttribute(render_tiles_kernel<OutT, SLOTS, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kTileSmem));
        configured.fetch_or(bit, std::memory_order_release);
    }
    render_tiles_kernel<OutT, SLOTS, CTAS><<<ctas, kTileWarps * 32, kTileSmem, s>>>(
        g, box, (const TUnit *)w.pair_spot, w.spots, w.edges, w.edge_cap, sat, w.tile_start, w.next_tile, w.wmax_bits, out,
        accumulate);
    return 0;
}

template <typename OutT>
static int launch_render_tiles(const Geo &g, const Workspace &w, const int64_t *sat, const float *box, OutT *out,
                               int accumulate, cudaStream_t s) {
    if (g.slots == 32) return launch_render_tiles_as<OutT, 32, 6>(g, w, sat, box, out, accumulate, s);
    return launch_render_tiles_as<OutT, 0, 6>(g, w, sat, box, out, accumulate, s);
}

template <typename OutT>
static int launch_render_reg(const CUtensorMap &box_map, const Geo &g, const Workspace &w, const int64_t *sat,
                             const float *box, OutT *out, int accumulate, cudaStream_t s) {
    if (render_variant().reg == 2) return launch_render_reg_as<OutT, 8, 3, 4>(box_map, g, w, sat, out, accumulate, s);
    if (g.slots != 32) return launch_render_ldg_as<OutT, 8, 3, 0, 1>(g, w, sat, box, out, accumulate, s);
    return launch_render_ldg_as<OutT, 8, 3, 32, 1>(g, w, sat, box, out, accumulate, s);
}

template <typename OutT>
static int launch_render(const Geo &g, const Workspace &w, int64_t n_spots, OutT *out, int accumulate, int box_type,
                         cudaStream_t s, const ListPlan &plan) {
    const RenderVariant v = render_variant();
    // no footprint can take the box-table path (no table, or the test hook): the instance without a copy ring
    if (!g.quick_runs && g.tile_h == 8) {
        if (box_type == SCB_F32) return launch_render_as<OutT, float, 8, 0, kCopyNone>(g, w, n_spots, out, accumulate, s, plan);
        return launch_render_as<OutT, double, 8, 0, kCopyNone>(g, w, n_spots, out, accumulate, s, plan);
    }
    if (box_type == SCB_F32) {
        if (g.tile_h == 16)
            return v.copy ? launch_render_slots<OutT, float, 16, 1>(g, w, n_spots, out, accumulate, s, plan)
                          : launch_render_slots<OutT, float, 16, 0>(g, w, n_spots, out, accumulate, s, plan);
        if (v.copy == 2) return launch_render_slots<OutT, float, 8, 2>(g, w, n_spots, out, accumulate, s, plan);
        return v.copy ? launch_render_slots<OutT, float, 8, 1>(g, w, n_spots, out, accumulate, s, plan)
                      : launch_render_slots<OutT, float, 8, 0>(g, w, n_spots, out, accumulate, s, plan);
    }
    return v.copy == 1 ? launch_render_slots<OutT, double, 8, 1>(g, w, n_spots, out, accumulate, s, plan)
                       : launch_render_slots<OutT, double, 8, 0>(g, w, n_spots, out, accumulate, s, plan);
}

static int render_expected_strided(const scb_geometry *geom, int64_t n_spots, int frames, int64_t stride,
                                   const int32_t *d_order, const double *d_depth,
                                   const double *d_x, const double *d_y, const double *d_weight,
                                   const int64_t *d_sat, const void *d_box, int box_type, const double *d_inv_scale,
                                   const int32_t *d_slot_of_key, void *d_out, int out_type,
                                   int accumulate, void *d_workspace, size_t workspace_bytes,
                                   int32_t *d_errors, void *stream, int plan_mode = 0) {
    int rc = check_geometry(geom);
    if (rc) return rc;
    SCB_REQUIRE(n_spots >= 0 && n_spots < (int64_t)1 << 30, SCB_E_INVALID, "n_spots=%lld", (long long)n_spots);
    SCB_REQUIRE(d_out && d_workspace && d_errors, SCB_E_NULL, "scb_render_expected: NULL out/workspace/errors");
    SCB_REQUIRE(n_spots == 0 || (d_depth && d_x && d_y && d_weight && d_sat && d_inv_scale && d_slot_of_key),
                SCB_E_NULL, "scb_render_expected: NULL spot/table pointer");
    SCB_REQUIRE(out_type == SCB_F32 || out_type == SCB_F64, SCB_E_INVALID, "out_type=%d", out_type);
    SCB_REQUIRE(box_type == SCB_F32 || box_type == SCB_F64, SCB_E_INVALID, "box_type=%d", box_type);
    if (!d_box) box_type = SCB_F64;       // no box table: every footprint is summed exactly from the SAT (64-bit accumulators)
    const int box_bytes = box_type == SCB_F32 ? 4 : 8;
    SCB_REQUIRE(frames >= 1 && frames <= 4096 && n_spots % frames == 0, SCB_E_INVALID,
                "scb_render_expected: %lld spots do not split into %d frames", (long long)n_spots, frames);
    Geo g = strip_geo(geom, d_box != nullptr, box_bytes, frames, n_spots / frames);
    if (plan_mode != 0 && frames > 1) g.stripes = 1;      // list plans index one counter per strip
    // the CTA-tile kernel (render_tile.cuh): 32 x 128 tiles, one list entry per (spot, tile, run of 32 columns)
    const bool tile_path = render_variant().reg == 3 && forced_gather() != 1 && d_box && box_bytes == 4 &&
                           g.slots <= 32 && g.slots % 4 == 0;
    if (tile_path) {
        const int quick = forced_gather() ? 0 : 1;
        const double box_peak = g.box_peak;
        g = make_geo(geom, kTileRows, kTileCols, kUnitCols, frames);
        g.spots_per_frame = n_spots / frames;
        g.box_peak = box_peak;
        g.special_edges = 1;
        g.quick_runs = quick;
    }
    Workspace w = carve(g, n_spots, d_workspace, sizeof(Unit), true);
    SCB_REQUIRE(workspace_bytes >= w.bytes, SCB_E_WORKSPACE, "scb_render_expected: workspace %zu < %zu",
                workspace_bytes, w.bytes);
    SCB_REQUIRE(g.tile_w / g.chunk < 8, SCB_E_UNSUPPORTED, "scb_render_expected: strips of %d columns in chunks of %d",
                g.tile_w, g.chunk);     // the census votes on three bits of a (spot, strip) overlap's entry count
    SCB_REQUIRE((double)n_spots * 2.0 * w.edge_cap < 4294967296.0, SCB_E_UNSUPPORTED,
                "scb_render_expected: %lld spots x %d pixel edges exceed the 32-bit edge index",
                (long long)n_spots, 2 * w.edge_cap);
    cudaStream_t s = (cudaStream_t)stream;
    const int n_tiles = g.frames * g.nti * g.ntj;
    // census, cursors, weight maximum and the strip queue are adjacent 256-aligned blocks: clear them all
    SCB_CUDA(cudaMemsetAsync(w.tile_count, 0, (size_t)((char *)w.tile_start - (char *)w.tile_count), s));
    const bool reg_path = !tile_path && forced_gather() != 1 && reg_path_for(d_box != nullptr, box_bytes, g.slots);
    // A planned block render (plan_mode bit 0: the workspace holds the list plan the previous call left behind for the
    // same geometry, spots and frames) counts, places and writes the units in one pass; bit 1 leaves a plan behind.
    // Only the shared-memory kernel's units are written that way.
    const bool can_plan = !tile_path && !reg_path && frames > 1 && g.stripes == 1 && w.plan_base != nullptr && n_spots > 0;
    const bool planned = can_plan && (plan_mode & 1);
    ListPlan plan = {nullptr, w.tile_count, (int)(w.pair_capacity < INT_MAX ? w.pair_capacity : INT_MAX), w.overflow_count,
                     w.overflow_tile, (Unit *)w.overflow_units};
    if (planned) plan.base = w.plan_base;
    if (planned) {
        // 4 CTAs per SM (64 registers); measured: 3 (72 registers) -1.5 %, 5 (48 registers, spills) -2 % on the C4 step
        spot_bin_fused_kernel<4><<<dim3(scb_grid_for(n_spots / frames, 256), frames), 256, 0, s>>>(
            g, n_spots, stride, d_depth, d_x, d_y, d_weight, d_inv_scale, d_slot_of_key, w.spots, w.tile_count, w.wmax_bits,
            d_errors, d_order, w.walk_list, w.walk_count, w.edge_cap, d_sat, d_box, box_bytes, (Unit *)w.pair_spot, plan);
        dim3 egrid, eblock;
        edges_launch_shape(w.edge_cap, n_spots, egrid, eblock);
        const unsigned cap = (unsigned)(SCB_SM_COUNT * 16);
        if (egrid.x > cap) egrid.x = cap;
        spot_edges_kernel<<<egrid, eblock, 0, s>>>(g, n_spots, w.spots, w.edges, w.edge_cap, w.walk_list, w.walk_count);
    } else {
    if (n_spots > 0) {
        // SCB_PREPARE_CENSUS=global: the census by global atomics only (the measured alternative)
        const char *census = getenv("SCB_PREPARE_CENSUS");
        auto prepare = census && census[0] == 'g' ? spot_prepare_kernel<6, false>
                                                  : spot_prepare_kernel<SCB_PREPARE_CTAS, true>;
        prepare<<<dim3(scb_grid_for(n_spots / frames, 256), frames), 256, 0, s>>>(
            g, n_spots, stride, d_depth, d_x, d_y, d_weight, d_inv_scale, d_slot_of_key, w.spots, w.tile_count,
            w.wmax_bits, d_errors, w.ranks, w.rank_cap, d_order, w.walk_list, w.walk_count);
        // the edge kernel walks the listed footprints only: a grid-stride loop over the list, sized for the device
        dim3 egrid, eblock;
        edges_launch_shape(w.edge_cap, n_spots, egrid, eblock);
        const unsigned cap = (unsigned)(SCB_SM_COUNT * 16);
        if (egrid.x > cap) egrid.x = cap;
        spot_edges_kernel<<<egrid, eblock, 0, s>>>(g, n_spots, w.spots, w.edges, w.edge_cap, w.walk_list, w.walk_count);
    }
    tile_scan_kernel<false><<<kScanCtas, 1024, 0, s>>>(n_tiles, g.stripes, w.tile_count, w.tile_start);
    }
    CUtensorMap box_map;
    if (reg_path && render_variant().reg == 2) {
        rc = make_box_map(&box_map, d_box, g.slots, (long long)(g.n_depth_keys + 1) * g.modulus * g.modulus);
        if (rc) return rc;
    }
    if (n_spots > 0 && !planned) {
        if (tile_path)
            tile_fill_units_kernel<<<scb_grid_for(n_spots, 256), 256, 0, s>>>(g, n_spots, w.spots, w.ranks, w.rank_cap,
                                                                             w.tile_start, (TUnit *)w.pair_spot);
        else if (reg_path)
            strip_fill_reg_kernel<<<scb_grid_for(n_spots, 256), 256, 0, s>>>(g, n_spots, w.spots, w.ranks, w.rank_cap,
                                                                            w.tile_start, (RUnit *)w.pair_spot);
        else {
            // SCB_FILL_VARIANT=old: list positions loaded strip by strip (the measured alternative: 119 vs 112 us per
            // 16-frame C4 block; with 4 / 2 CTAs per SM the up-front requests take 124 / 135 us)
            const char *fv = getenv("SCB_FILL_VARIANT");
            auto fill = fv && fv[0] == 'o' ? strip_fill_kernel<4, 0> : strip_fill_kernel<SCB_FILL_CTAS, 10>;
            fill<<<scb_grid_for(n_spots, 256), 256, 0, s>>>(g, n_spots, w.spots, w.edge_cap, w.ranks, w.rank_cap, d_sat, d_box,
                                                            box_bytes, w.tile_start, w.wmax_bits, (Unit *)w.pair_spot);
        }
    }
    int timed = -1;        // slot of this launch in the measurement hook's event pool
    if (g_profile.enabled) {
        std::lock_guard<std::mutex> lock(g_profile.mu);
        if (g_profile.enabled && g_profile.used < g_profile.capacity) timed = g_profile.used++;
    }
    if (timed >= 0) cudaEventRecord(g_profile.start[timed], s);
    if (tile_path) {
        if (out_type == SCB_F32) rc = launch_render_tiles<float>(g, w, d_sat, (const float *)d_box, (float *)d_out, accumulate, s);
        else rc = launch_render_tiles<double>(g, w, d_sat, (const float *)d_box, (double *)d_out, accumulate, s);
    } else if (reg_path) {
        if (out_type == SCB_F32) rc = launch_render_reg<float>(box_map, g, w, d_sat, (const float *)d_box, (float *)d_out, accumulate, s);
        else rc = launch_render_reg<double>(box_map, g, w, d_sat, (const float *)d_box, (double *)d_out, accumulate, s);
    } else if (out_type == SCB_F32) rc = launch_render<float>(g, w, n_spots, (float *)d_out, accumulate, box_type, s, plan);
    else rc = launch_render<double>(g, w, n_spots, (double *)d_out, accumulate, box_type, s, plan);
    if (timed >= 0) cudaEventRecord(g_profile.stop[timed], s);
    if (rc) return rc;
    // the plan for the next block of the same shape: every list's room from this block's census
    if (can_plan && (plan_mode & 2)) {
        const char *tight = getenv("SCB_PLAN_TIGHT");       // test hook: plans that are too small
        tile_scan_kernel<true><<<kScanCtas, 1024, 0, s>>>(n_tiles, 1, w.tile_count, w.plan_base, tight && tight[0] == '1');
    }
    SCB_CUDA_LAUNCH_CHECK("scb_render_expected");
    return 0;
}

extern "C" int scb_render_expected(const scb_geometry *geom, int64_t n_spots, const double *d_depth,
                                   const double *d_x, const double *d_y, const double *d_weight,
                                   const int64_t *d_sat, const void *d_box, int box_type, const double *d_inv_scale,
                                   const int32_t *d_slot_of_key, void *d_out, int out_type,
                                   int accumulate, void *d_workspace, size_t workspace_bytes,
                                   int32_t *d_errors, void *stream) {
    return render_expected_strided(geom, n_spots, 1, 1, nullptr, d_depth, d_x, d_y, d_weight, d_sat, d_box, box_type, d_inv_scale,
                                   d_slot_of_key, d_out, out_type, accumulate, d_workspace, workspace_bytes, d_errors,
                                   stream);
}

// ---- particle rows as the host API holds them: (n, 5) float64 rows (depth, x, y, molecule id,
// p_state), the output of EPIFMSimulator.__format_data (base.py:61-110), read in place
extern "C" int scb_emit_bleach_rows(uint64_t budget_seed, int64_t n, const double *d_rows,
                                    const int32_t *d_mol_slot, double unit_time, double focal_depth,
                                    const scb_photophysics *phys, double *d_budget, double *d_weight,
                                    double *d_true_data, void *stream) {
    SCB_REQUIRE(n == 0 || d_rows, SCB_E_NULL, "scb_emit_bleach_rows: NULL rows");
    return scb_emit_bleach_strided(budget_seed, n, d_rows, d_rows + 1, d_rows + 2, d_rows + 4, d_mol_slot, nullptr,
                                   d_rows + 3, 5, unit_time, focal_depth, phys, d_budget, d_weight, d_true_data,
                                   stream);
}

extern "C" int scb_render_expected_rows(const scb_geometry *geom, int64_t n, const double *d_rows,
                                        const double *d_weight, const int64_t *d_sat, const void *d_box, int box_type,
                                        const double *d_inv_scale, const int32_t *d_slot_of_key, void *d_out,
                                        int out_type, int accumulate, void *d_workspace, size_t workspace_bytes,
                                        int32_t *d_errors, void *stream) {
    SCB_REQUIRE(n == 0 || d_rows, SCB_E_NULL, "scb_render_expected_rows: NULL rows");
    return render_expected_strided(geom, n, 1, 5, nullptr, d_rows, d_rows + 1, d_rows + 2, d_weight, d_sat, d_box, box_type, d_inv_scale,
                                   d_slot_of_key, d_out, out_type, accumulate, d_workspace, workspace_bytes, d_errors,
                                   stream);
}

// A block of movie frames in one call: frame f takes spots [f * n_per_frame, (f + 1) * n_per_frame)
// of the arrays and image f of d_out.  Every kernel of the pipeline runs once for the whole block
// (the binning kernels are latency bound, so a block costs little more than a frame).
extern "C" int scb_render_expected_frames_ordered(const scb_geometry *geom, int64_t n_per_frame, int n_frames,
                                                  const int32_t *d_order, const double *d_depth, const double *d_x,
                                                  const double *d_y, const double *d_weight, const int64_t *d_sat,
                                                  const void *d_box, int box_type, const double *d_inv_scale,
                                                  const int32_t *d_slot_of_key, void *d_out, int out_type,
                                                  void *d_workspace, size_t workspace_bytes, int32_t *d_errors,
                                                  void *stream) {
    SCB_REQUIRE(n_per_frame >= 0 && n_frames >= 1, SCB_E_INVALID, "scb_render_expected_frames: n=%lld frames=%d",
                (long long)n_per_frame, n_frames);
    SCB_REQUIRE(n_per_frame < ((int64_t)1 << 31), SCB_E_INVALID, "scb_render_expected_frames: n=%lld", (long long)n_per_frame);
    return render_expected_strided(geom, n_per_frame * n_frames, n_frames, 1, d_order, d_depth, d_x, d_y, d_weight, d_sat,
                                   d_box, box_type, d_inv_scale, d_slot_of_key, d_out, out_type, 0, d_workspace,
                                   workspace_bytes, d_errors, stream);
}

extern "C" int scb_render_expected_frames_planned(const scb_geometry *geom, int64_t n_per_frame, int n_frames,
                                                  const int32_t *d_order, const double *d_depth, const double *d_x,
                                                  const double *d_y, const double *d_weight, const int64_t *d_sat,
                                                  const void *d_box, int box_type, const double *d_inv_scale,
                                                  const int32_t *d_slot_of_key, void *d_out, int out_type,
                                                  void *d_workspace, size_t workspace_bytes, int32_t *d_errors,
                                                  int plan_mode, void *stream) {
    SCB_REQUIRE(n_per_frame >= 0 && n_frames >= 1, SCB_E_INVALID, "scb_render_expected_frames_planned: n=%lld frames=%d",
                (long long)n_per_frame, n_frames);
    SCB_REQUIRE(n_per_frame < ((int64_t)1 << 31), SCB_E_INVALID, "scb_render_expected_frames_planned: n=%lld", (long long)n_per_frame);
    SCB_REQUIRE(plan_mode >= 0 && plan_mode <= 3, SCB_E_INVALID, "scb_render_expected_frames_planned: plan_mode=%d", plan_mode);
    return render_expected_strided(geom, n_per_frame * n_frames, n_frames, 1, d_order, d_depth, d_x, d_y, d_weight, d_sat,
                                   d_box, box_type, d_inv_scale, d_slot_of_key, d_out, out_type, 0, d_workspace,
                                   workspace_bytes, d_errors, stream, plan_mode);
}

extern "C" int scb_render_expected_frames(const scb_geometry *geom, int64_t n_per_frame, int n_frames,
                                          const double *d_depth, const double *d_x, const double *d_y,
                                          const double *d_weight, const int64_t *d_sat, const void *d_box,
                                          int box_type, const double *d_inv_scale, const int32_t *d_slot_of_key,
                                          void *d_out, int out_type, void *d_workspace, size_t workspace_bytes,
                                          int32_t *d_errors, void *stream) {
    return scb_render_expected_frames_ordered(geom, n_per_frame, n_frames, nullptr, d_depth, d_x, d_y, d_weight, d_sat,
                                              d_box, box_type, d_inv_scale, d_slot_of_key, d_out, out_type, d_workspace,
                                              workspace_bytes, d_errors, stream);
}

// A block of frames given as particle ROWS: frame f is the (n_per_frame, 5) float64 rows at
// d_rows + f * n_per_frame * 5 (the facade's snapshots of consecutive frames, uploaded back to back)
extern "C" int scb_render_expected_rows_frames(const scb_geometry *geom, int64_t n_per_frame, int n_frames,
                                               const double *d_rows, const double *d_weight, const int64_t *d_sat,
                                               const void *d_box, int box_type, const double *d_inv_scale,
                                               const int32_t *d_slot_of_key, void *d_out, int out_type,
                                               void *d_workspace, size_t workspace_bytes, int32_t *d_errors,
                                               void *stream) {
    SCB_REQUIRE(n_per_frame >= 0 && n_frames >= 1, SCB_E_INVALID, "scb_render_expected_rows_frames: n=%lld frames=%d",
                (long long)n_per_frame, n_frames);
    SCB_REQUIRE(n_per_frame < ((int64_t)1 << 31), SCB_E_INVALID, "scb_render_expected_rows_frames: n=%lld", (long long)n_per_frame);
    SCB_REQUIRE(n_per_frame == 0 || d_rows, SCB_E_NULL, "scb_render_expected_rows_frames: NULL rows");
    return render_expected_strided(geom, n_per_frame * n_frames, n_frames, 5, nullptr, d_rows, d_rows + 1, d_rows + 2, d_weight,
                                   d_sat, d_box, box_type, d_inv_scale, d_slot_of_key, d_out, out_type, 0, d_workspace,
                                   workspace_bytes, d_errors, stream);
}
