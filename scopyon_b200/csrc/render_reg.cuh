// fp32 render with register accumulators and tensor-map TMA (included by render.cu, inside its
// anonymous namespace, after the mbarrier helpers and strip_shift).
//
// Reference: PointSpreadingFunction.overlay_signal_ (/root/reference/src/scopyon/_epifm.py:224-282), the
// same box sums and the same arithmetic as render_strips_kernel<float, float, 8, ..> -- the two kernels
// produce bit-identical images (tests/test_gpu_render.py) -- with the work laid out differently:
//
//   * one warp still owns one 8 x 128-pixel strip, but its 1024 accumulators are 32 REGISTERS per lane:
//     lane l owns the strip columns l, l + 32, l + 64, l + 96 of all eight rows.  The accumulators never
//     touch shared memory (the older kernel spends two of its three shared-memory wavefronts per pixel
//     row on them and runs at the LSU's wavefront rate);
//   * a unit is the overlap of one spot with the strip and one run of 32 columns starting at strip column
//     c0.  Lane l takes the column congruent to l: stage column (l - c0) mod 32 -- a rotation, hence bank
//     conflict free -- which belongs to column group c0 / 32 for the lanes at or above c0 mod 32 and to
//     the next group for the others (dropped when that group lies beyond the strip: the neighbouring
//     strip has a unit of its own);
//   * the box values arrive by ONE tensor-map TMA copy per unit (cp.async.bulk.tensor.3d, SASS UTMALDG)
//     of a fixed 8-row x 36-column box out of the box table seen as a 3-D tensor [block][row][column].
//     Stage row k is strip row k whatever rows of the strip the footprint covers: the copy starts at a
//     NEGATIVE table row when the footprint begins below the strip's first row, and the TMA unit fills
//     what lies outside the block -- rows above the footprint, columns to its right -- with zeros
//     without reading memory (rows and columns past the footprint but inside the block hold zeros in the
//     table itself).  The consumer therefore needs no row or column predicate and no per-unit shape: eight
//     shared-memory loads at fixed offsets, eight multiplies and conversions, sixteen predicated adds.
//   * irregular footprints (SAT-corner gather) form their eight box values per lane from global memory and
//     join the same accumulation.
#pragma once

struct __align__(16) RUnit {
    double ws;          // weight * res^2 / table scale (scaled to accumulator LSBs by the fetching lane)
    int32_t where;      // fast: block of the box table (table * M * M + row phase * M + column phase); gather: spot
    uint32_t packed;    // fast:   1 << 31 | c0 << 20 | (first table row + 8) << 10 | first table column
                        // gather:           c0 << 20 | first strip row << 17 | rows << 13 | columns << 7
};
constexpr uint32_t kRUnitFast = 0x80000000u;
constexpr int kRegBatch = 32;                       // units fetched per round: one per lane
// a stage holds the copy box: 8 rows of 36 columns.  The innermost start coordinate of a tensor-map copy must be a
// multiple of 16 bytes (tools/probes/tma_box_probe.cu: anything else is an illegal instruction), so the copy starts at
// the unit's first table column rounded down to a multiple of four and is four columns wider; the lanes skip the
// 0..3 leading columns.  9 x 128 bytes per stage: stages stay 128-byte aligned, rows rotate through the banks.
constexpr int kStageCols = 36;
constexpr int kStageBytes = 8 * kStageCols * 4;
// ring depth STAGES: copies run STAGES - 1 units ahead of the arithmetic; TMA boxes land on 128-byte boundaries
template <int STAGES> constexpr size_t reg_warp_smem() { return (STAGES * kStageBytes + kRegBatch * 8 + 64 + 127) & ~(size_t)127; }

__global__ void __launch_bounds__(256)
strip_fill_reg_kernel(Geo g, int64_t n, const SpotRec *__restrict__ spots, const int *__restrict__ ranks, int rank_cap,
                      const int *__restrict__ tile_start, RUnit *__restrict__ units) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const SpotRec rec = spots[s];
    if (rec.slot < 0) return;
    RUnit u;
    u.ws = rec.w;
    const bool fast = g.quick_runs && rec.row_run >= 0 && rec.col_run >= 0;
    // edge e of a regular axis sits at slot slot0 + e (>= -1); the pixel between edges e and e + 1 is box-table
    // row / column slot0 + e + 1
    const int row_slot0 = (rec.row_run >> 16) - 1, col_slot0 = (rec.col_run >> 16) - 1;
    const int block = (rec.slot * g.modulus + (rec.row_run & 0xffff)) * g.modulus + (rec.col_run & 0xffff);
    const int stripe = stripe_of(g, s, rec.frame);
    const int frame_tile0 = rec.frame * g.nti * g.ntj;
    const int *my_rank = ranks + (size_t)s * rank_cap;
    int visited = 0;
    const int t0 = rec.imin / g.tile_h, t1 = (rec.imax - 1) / g.tile_h;
    const int u0 = rec.jmin / g.tile_w, u1 = (rec.jmax - 1) / g.tile_w;
    for (int tj = u0; tj <= u1; ++tj) {
        const int c_lo = max(rec.jmin, tj * g.tile_w), c_hi = min(rec.jmax, (tj + 1) * g.tile_w);
        const int entries = (c_hi - c_lo + g.chunk - 1) / g.chunk;
        for (int ti = t0; ti <= t1; ++ti) {
            const int r_lo = max(rec.imin, ti * g.tile_h), r_hi = min(rec.imax, (ti + 1) * g.tile_h);
            const int tile = frame_tile0 + ti * g.ntj + tj;
            RUnit *dst = units + tile_start[tile * g.stripes + stripe] + __ldg(my_rank + visited);
            ++visited;
            for (int q = 0; q < entries; ++q) {
                const int c = c_lo + q * g.chunk;
                const uint32_t c0 = (uint32_t)(c - tj * g.tile_w);
                if (fast) {
                    u.where = block;
                    const int first_row = row_slot0 + 1 + (ti * g.tile_h - rec.imin);     // table row of strip row 0 (>= -7)
                    const int first_col = col_slot0 + 1 + (c - rec.jmin);
                    u.packed = kRUnitFast | c0 << 20 | (uint32_t)(first_row + 8) << 10 | (uint32_t)first_col;
                } else {
                    u.where = (int32_t)s;
                    u.packed = c0 << 20 | (uint32_t)(r_lo - ti * g.tile_h) << 17 | (uint32_t)(r_hi - r_lo) << 13 |
                               (uint32_t)min(g.chunk, c_hi - c) << 7;
                }
                dst[q] = u;
            }
        }
    }
}

__device__ __forceinline__ void tma_box(void *dst, const CUtensorMap *map, int col, int row, int block, void *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_addr(dst)),
        "l"(map), "r"(col), "r"(row), "r"(block), "r"(smem_addr(bar))
        : "memory");
}

// Box values of a gather unit for this lane's column: rows outside the unit and idle lanes get zero.  The four
// SAT corners of a pixel come through per-edge table offsets (`edges`); a lane fetches the corners on its left
// edge only and takes the right ones from the lane holding the next column (the last column fetches both).
// Returned by value (registers): the caller's fast path must not see an array in local memory.
struct Box8 {
    float v[8];
};
__device__ __noinline__ Box8 gather_unit(uint32_t packed, int spot, int lane, const SpotRec *__restrict__ spots,
                                         const uint32_t *__restrict__ edges, int edge_cap,
                                         const int64_t *__restrict__ sat, long long table_entries, int row0, int col0) {
    const int c0 = (packed >> 20) & 127, r0 = (packed >> 17) & 7, n_rows = (packed >> 13) & 15, n_cols = (packed >> 7) & 63;
    const int imin = spots[spot].imin, jmin = spots[spot].jmin, slot = spots[spot].slot;
    const long long *table = reinterpret_cast<const long long *>(sat) + (size_t)slot * table_entries;
    const int s = (lane - c0) & 31;
    const bool active = s < n_cols;
    const uint32_t ebase = (uint32_t)spot * 2u * (uint32_t)edge_cap;
    const uint32_t erow = ebase + (uint32_t)(row0 + r0 - imin);
    const uint32_t c = ebase + (uint32_t)(edge_cap + col0 + c0 - jmin) + (uint32_t)min(s, n_cols - 1);
    const uint32_t left = __ldg(edges + c), right = __ldg(edges + c + 1);
    const bool last_lane = s == n_cols - 1 || s == 31;
    const int next_lane = (lane + 1) & 31;
    long long before = 0;
    Box8 out;
#pragma unroll
    for (int k = 0; k <= 8; ++k) {               // row edge k of the strip: edge clamp(k - r0, 0, n_rows) of the unit
        const uint32_t row = __ldg(edges + erow + (uint32_t)min(max(k - r0, 0), n_rows));
        long long L = 0, own_right = 0;
        if (!((row | left) & kEdgeZero)) L = __ldg(table + (row + left));
        if (last_lane && !((row | right) & kEdgeZero)) own_right = __ldg(table + (row + right));
        const long long from_neighbour = __shfl_sync(0xffffffffu, L, next_lane);
        const long long here = (last_lane ? own_right : from_neighbour) - L;
        // >= 0: the table is non-negative, edges are monotone; rounded as the box table is.  Rows outside the
        // unit see the same edge twice: zero.
        if (k > 0) out.v[k - 1] = active ? __ll2float_rn(here - before) : 0.0f;
        before = here;
    }
    return out;
}

// acc += d on the lanes where `on` holds: one predicated add (left to itself the compiler selects between d and
// zero first: two instructions)
__device__ __forceinline__ void add_if(int &acc, int d, int on) {
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q add.s32 %0, %0, %1;\n\t}" : "+r"(acc) : "r"(d), "r"(on));
}

template <int Q>
__device__ __forceinline__ void group_add(int (&acc)[8][4], const int (&d)[8], bool in_first) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        add_if(acc[k][Q], d[k], in_first);
        if (Q < 3) add_if(acc[k][Q + 1], d[k], !in_first);
    }
}

template <typename OutT, int WARPS, int CTAS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32, CTAS)
render_strips_reg_kernel(const __grid_constant__ CUtensorMap box_map, Geo g, const RUnit *__restrict__ units,
                         const SpotRec *__restrict__ spots, const uint32_t *__restrict__ edges, int edge_cap,
                         const int64_t *__restrict__ sat, const int *__restrict__ tile_start,
                         int *__restrict__ next_tile, const unsigned long long *__restrict__ wmax_bits,
                         OutT *__restrict__ out, int accumulate) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *mine = smem_raw + warp * reg_warp_smem<STAGES>();
    float *ring = reinterpret_cast<float *>(mine);
    uint2 *meta = reinterpret_cast<uint2 *>(mine + STAGES * kStageBytes);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(meta + kRegBatch);

    const int frame_tiles = g.nti * g.ntj, n_tiles = g.frames * frame_tiles;
    const long long table_entries = (long long)g.modulus * g.modulus * g.slots * g.slots;

    if (lane == 0) {
        for (int st = 0; st < STAGES; ++st) mbar_init(&bars[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    int p_stage = 0, c_stage = 0;
    uint32_t c_parity = 0;
    int acc[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[k][q] = 0;

    for (;;) {
        int tile = 0;
        if (lane == 0) tile = atomicAdd(next_tile, 1);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= n_tiles) break;
        const int frame = tile / frame_tiles, in_frame = tile - frame * frame_tiles;
        const int ti = in_frame / g.ntj, tj = in_frame - ti * g.ntj;
        const int row0 = ti * 8, col0 = tj * 128;
        const int seg_begin = tile_start[tile * g.stripes], seg_end = tile_start[(tile + 1) * g.stripes];
        const int shift = strip_shift(seg_end - seg_begin, wmax_bits[frame], g.box_peak);
        const double scale = scalbn(1.0, shift), lsb = scalbn(1.0, -shift);

        for (int base = seg_begin; base < seg_end; base += kRegBatch) {
            const int nb = min(kRegBatch, seg_end - base);
            __syncwarp();                                  // the previous batch's units have been read by every lane
            // lane u fetches unit u of the batch, publishes {weight in accumulator LSBs, packed word} and keeps
            // what its own TMA copy needs
            int my_where = 0;
            uint32_t my_packed = 0;
            if (lane < nb) {
                const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(units + base + lane));
                const double ws = __longlong_as_double(((long long)raw.y << 32) | raw.x);
                my_where = (int)raw.z;
                my_packed = raw.w;
                meta[lane] = make_uint2(__float_as_uint((float)(ws * scale)), raw.w);
            }
            const uint32_t fast_mask = __ballot_sync(0xffffffffu, (my_packed & kRUnitFast) != 0);
            __syncwarp();

            auto stage_unit = [&](int u) {
                if ((fast_mask >> u) & 1u) {               // warp uniform
                    if (lane == u) {
                        void *bar = &bars[p_stage];
                        mbar_expect_tx(bar, kStageBytes);
                        tma_box(ring + p_stage * (kStageBytes / 4), &box_map, (int)(my_packed & 1020u),
                                (int)((my_packed >> 10) & 1023u) - 8, my_where, bar);
                    }
                    if (++p_stage == STAGES) p_stage = 0;
                }
            };
            for (int u = 0; u < min(nb, STAGES - 1); ++u) stage_unit(u);
            for (int u = 0; u < nb; ++u) {
                // every lane has finished reading the ring stage the next copy overwrites (the one unit u - 1 used)
                __syncwarp();
                if (u + STAGES - 1 < nb) stage_unit(u + STAGES - 1);
                const uint2 m = meta[u];
                const float ws = __uint_as_float(m.x);
                const int c0 = (int)((m.y >> 20) & 127u);
                Box8 box;
                float (&v)[8] = box.v;
                if ((fast_mask >> u) & 1u) {
                    mbar_wait(&bars[c_stage], c_parity);
                    const float *st = ring + c_stage * (kStageBytes / 4) + ((lane - c0) & 31) + (int)(m.y & 3u);
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = st[k * kStageCols];
                    if (++c_stage == STAGES) { c_stage = 0; c_parity ^= 1u; }
                } else {
                    box = gather_unit(m.y, __shfl_sync(0xffffffffu, my_where, u), lane, spots, edges, edge_cap, sat,
                                      table_entries, row0, col0);
                }
                int d[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) d[k] = to_fixed(v[k], ws);     // _epifm.py:280-282: adding zero changes nothing
                const bool in_first = lane >= (c0 & 31);
                switch (c0 >> 5) {                                         // warp uniform
                case 0: group_add<0>(acc, d, in_first); break;
                case 1: group_add<1>(acc, d, in_first); break;
                case 2: group_add<2>(acc, d, in_first); break;
                default: group_add<3>(acc, d, in_first); break;
                }
            }
        }

        // ---- write the strip (coalesced 128-byte rows) and clear the accumulators for the next one
        OutT *image = out + (size_t)frame * g.n_w * g.n_h;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = row0 + k;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = col0 + q * 32 + lane;
                const double val = (double)acc[k][q] * lsb;
                acc[k][q] = 0;
                if (i < g.n_w && j < g.n_h) {
                    const size_t o = (size_t)i * g.n_h + j;
                    if (accumulate) image[o] = (OutT)((double)image[o] + val);
                    else image[o] = (OutT)val;
                }
            }
        }
    }
}

// ---- the same strip ownership and accumulation with the box values loaded STRAIGHT INTO REGISTERS -----------------
// A unit's eight rows are eight coalesced 128-byte global loads (lane l reads its own column of every row), issued
// DIST units ahead of the arithmetic into a rotating set of register arrays: no copy engine, no shared-memory stage,
// no mbarrier.  Rows outside the block (a footprint that begins below the strip's first row, or ends above its last)
// and columns past the block's end are not read and count as zero -- the predicates replace the TMA unit's fill.
// tools/probes/random_chunk_probe.cu: random 1 KB reads by plain loads reach 6.7 TB/s on this part, so the access
// pattern itself is no obstacle.
constexpr size_t kLdgWarpSmem = kRegBatch * 16;     // the published batch: {weight, packed word, block, -} per unit

template <int SLOTS>
__device__ __forceinline__ void load_rows(float (&v)[8], const float *__restrict__ box, int runtime_slots, uint4 m, int lane) {
    const int slots = SLOTS ? SLOTS : runtime_slots;
    const int c0 = (int)((m.y >> 20) & 127u);
    const int col = (int)(m.y & 1023u) + ((lane - c0) & 31);
    const int row = (int)((m.y >> 10) & 1023u) - 8;
    const bool col_ok = col < slots;
    const float *p = box + ((size_t)(int)m.z * (size_t)slots + (size_t)(long long)row) * (size_t)slots + (size_t)col;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k] = 0.0f;
        if (col_ok && (unsigned)(row + k) < (unsigned)slots) v[k] = __ldg(p + k * slots);
    }
}

template <typename OutT, int WARPS, int CTAS, int SLOTS, int DIST>
__global__ void __launch_bounds__(WARPS * 32, CTAS)
render_strips_ldg_kernel(Geo g, const float *__restrict__ box, const RUnit *__restrict__ units,
                         const SpotRec *__restrict__ spots, const uint32_t *__restrict__ edges, int edge_cap,
                         const int64_t *__restrict__ sat, const int *__restrict__ tile_start,
                         int *__restrict__ next_tile, const unsigned long long *__restrict__ wmax_bits,
                         OutT *__restrict__ out, int accumulate) {
    constexpr int kRing = DIST + 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *meta = reinterpret_cast<uint4 *>(smem_raw + warp * kLdgWarpSmem);

    const int frame_tiles = g.nti * g.ntj, n_tiles = g.frames * frame_tiles;
    const long long table_entries = (long long)g.modulus * g.modulus * g.slots * g.slots;
    int acc[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[k][q] = 0;

    for (;;) {
        int tile = 0;
        if (lane == 0) tile = atomicAdd(next_tile, 1);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= n_tiles) break;
        const int frame = tile / frame_tiles, in_frame = tile - frame * frame_tiles;
        const int ti = in_frame / g.ntj, tj = in_frame - ti * g.ntj;
        const int row0 = ti * 8, col0 = tj * 128;
        const int seg_begin = tile_start[tile * g.stripes], seg_end = tile_start[(tile + 1) * g.stripes];
        const int shift = strip_shift(seg_end - seg_begin, wmax_bits[frame], g.box_peak);
        const double scale = scalbn(1.0, shift), lsb = scalbn(1.0, -shift);

        for (int base = seg_begin; base < seg_end; base += kRegBatch) {
            const int nb = min(kRegBatch, seg_end - base);
            __syncwarp();                                  // the previous batch has been read by every lane
            if (lane < nb) {                               // lane u fetches unit u and publishes it
                const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(units + base + lane));
                const double ws = __longlong_as_double(((long long)raw.y << 32) | raw.x);
                meta[lane] = make_uint4(__float_as_uint((float)(ws * scale)), raw.w, raw.z, 0u);
                if (raw.w & kRUnitFast) {
                    // ... and asks the L2 for its rows now: by the time the unit's turn comes (up to 31 units later)
                    // the loads below find them there, ~300 instead of ~1 500 cycles away
                    const int slots = SLOTS ? SLOTS : g.slots;
                    const int row = (int)((raw.w >> 10) & 1023u) - 8, col = (int)(raw.w & 1020u);
                    const float *p = box + ((size_t)(int)raw.z * (size_t)slots + (size_t)(long long)row) * (size_t)slots + (size_t)col;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if ((unsigned)(row + k) < (unsigned)slots)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(p + k * slots));
                }
            }
            __syncwarp();

            float v[kRing][8];
            auto fetch = [&](int u, float (&dst)[8]) {
                const uint4 m = meta[u];
                if (m.y & kRUnitFast) {                    // warp uniform
                    load_rows<SLOTS>(dst, box, g.slots, m, lane);
                } else {
                    const Box8 b = gather_unit(m.y, (int)m.z, lane, spots, edges, edge_cap, sat, table_entries, row0, col0);
#pragma unroll
                    for (int k = 0; k < 8; ++k) dst[k] = b.v[k];
                }
            };
            auto consume = [&](int u, const float (&src)[8]) {
                const uint2 m = *reinterpret_cast<const uint2 *>(meta + u);
                const float ws = __uint_as_float(m.x);
                const int c0 = (int)((m.y >> 20) & 127u);
                int d[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) d[k] = to_fixed(src[k], ws);     // _epifm.py:280-282: adding zero changes nothing
                const bool in_first = lane >= (c0 & 31);
                switch (c0 >> 5) {                                           // warp uniform
                case 0: group_add<0>(acc, d, in_first); break;
                case 1: group_add<1>(acc, d, in_first); break;
                case 2: group_add<2>(acc, d, in_first); break;
                default: group_add<3>(acc, d, in_first); break;
                }
            };
#pragma unroll
            for (int r = 0; r < DIST; ++r)
                if (r < nb) fetch(r, v[r]);
            for (int u0 = 0; u0 < nb; u0 += kRing) {
#pragma unroll
                for (int r = 0; r < kRing; ++r) {
                    const int u = u0 + r;
                    if (u < nb) {                           // warp uniform
                        if (u + DIST < nb) fetch(u + DIST, v[(r + DIST) % kRing]);
                        consume(u, v[r]);
                    }
                }
            }
        }

        // ---- write the strip (coalesced 128-byte rows) and clear the accumulators for the next one
        OutT *image = out + (size_t)frame * g.n_w * g.n_h;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = row0 + k;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = col0 + q * 32 + lane;
                const double val = (double)acc[k][q] * lsb;
                acc[k][q] = 0;
                if (i < g.n_w && j < g.n_h) {
                    const size_t o = (size_t)i * g.n_h + j;
                    if (accumulate) image[o] = (OutT)((double)image[o] + val);
                    else image[o] = (OutT)val;
                }
            }
        }
    }
}

// The box table as a 3-D tensor [blocks][slots][slots] of floats; the copy box is 1 x 8 x 36.  The driver entry
// point is looked up at run time, so the library itself does not link against libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_box_map(CUtensorMap *map, const void *d_box, int slots, long long blocks) {
    static std::atomic<EncodeTiledFn> cached{nullptr};
    EncodeTiledFn encode = cached.load(std::memory_order_acquire);
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult found;
        SCB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &found));
        SCB_REQUIRE(fn && found == cudaDriverEntryPointSuccess, SCB_E_UNSUPPORTED,
                    "cuTensorMapEncodeTiled is not available from this driver");
        encode = (EncodeTiledFn)fn;
        cached.store(encode, std::memory_order_release);
    }
    const cuuint64_t dims[3] = {(cuuint64_t)slots, (cuuint64_t)slots, (cuuint64_t)blocks};
    const cuuint64_t strides[2] = {(cuuint64_t)slots * 4u, (cuuint64_t)slots * slots * 4u};
    const cuuint32_t box[3] = {(cuuint32_t)kStageCols, 8u, 1u}, elem[3] = {1u, 1u, 1u};
    const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(d_box), dims, strides, box, elem,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SCB_REQUIRE(rc == CUDA_SUCCESS, SCB_E_UNSUPPORTED, "cuTensorMapEncodeTiled failed (%d) for %d slots, %lld blocks", (int)rc,
                slots, blocks);
    return 0;
}
