// fp32 render, CTA-tile form: one CTA of four warps owns a 32-row x 128-column tile of the image and shares ONE TMA
// bulk copy per (spot, tile) (included by render.cu after render_reg.cuh, inside its anonymous namespace).
//
// Reference: PointSpreadingFunction.overlay_signal_ (/root/reference/src/scopyon/_epifm.py:224-282); same box sums,
// same per-pixel arithmetic (box * weight -> fixed point -> integer add) as render_strips_kernel.
//
// Why: tools/probes/tma_rate_probe.cu -- an SM retires one bulk copy per ~50 cycles whatever its size, and the
// strip kernels issue one per (spot, 8-row strip): 5.9 per spot.  Here a unit is the overlap of a spot with a whole
// tile: 2.5 copies per spot, each of all the footprint's rows inside the tile (<= 32 rows, <= 4 KB, contiguous in the
// box table).  The binning handles 2.5 instead of 5.9 list entries per spot as well.
//   * warp w of the CTA owns tile rows 8w .. 8w + 7 and keeps its 8 x 128 pixels in 32 registers per lane (lane l:
//     columns l, l + 32, l + 64, l + 96), as the register strip kernels do; a unit's 32 columns map to lanes by
//     rotation, (lane - c0) mod 32;
//   * the four warps walk the tile's list in step: every warp fetches the same batches of unit records, warp (u mod 4)
//     issues the copy of unit u + STAGES - 1 into the CTA's ring once all four have released that stage (an "empty"
//     mbarrier with four arrivals per phase), all four wait for the "full" mbarrier of unit u, each adds the rows that
//     fall into its band (none: it only releases the stage; all eight: a branch-free body; some: a predicated one);
//   * the 32-bit accumulators take their LSB per TILE (list length x largest weight x largest box value), so this kernel
//     agrees with the strip kernels to that LSB (a few 1e-7 of a strip's brightest pixel), not bit for bit.
#pragma once

struct __align__(16) TUnit {
    double ws;          // weight * res^2 / table scale
    int32_t where;      // fast: block of the box table; gather: spot
    uint32_t rows;      // first tile row | rows << 8 | first box-table row << 16 | fast flag << 31
    uint32_t cols;      // first tile column c0 | columns << 8 | first box-table column << 16
    uint32_t pad;
};
static_assert(sizeof(TUnit) == 32, "list entries are sized like the strip kernel's units");
constexpr uint32_t kTUnitFast = 0x80000000u;
constexpr int kTileRows = 32, kTileCols = 128, kTileWarps = 4;
constexpr int kTileStages = 8;
constexpr int kTileStageBytes = 32 * 32 * 4;        // 32 rows of <= 32 slots
// per CTA: ring | full + empty mbarriers | next-tile slot | per-warp batch of published units
constexpr size_t kTileMetaBytes = 32 * 16;
constexpr size_t kTileSmem = kTileStages * kTileStageBytes + 128 + kTileWarps * kTileMetaBytes;

__global__ void __launch_bounds__(256)
tile_fill_units_kernel(Geo g, int64_t n, const SpotRec *__restrict__ spots, const int *__restrict__ ranks, int rank_cap,
                       const int *__restrict__ tile_start, TUnit *__restrict__ units) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const SpotRec rec = spots[s];
    if (rec.slot < 0) return;
    TUnit u;
    u.ws = rec.w;
    u.pad = 0;
    const bool fast = g.quick_runs && rec.row_run >= 0 && rec.col_run >= 0;
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
            TUnit *dst = units + tile_start[tile * g.stripes + stripe] + __ldg(my_rank + visited);
            ++visited;
            const uint32_t row_word = (uint32_t)(r_lo - ti * g.tile_h) | (uint32_t)(r_hi - r_lo) << 8;
            for (int q = 0; q < entries; ++q) {
                const int c = c_lo + q * g.chunk;
                const uint32_t col_word = (uint32_t)(c - tj * g.tile_w) | (uint32_t)min(g.chunk, c_hi - c) << 8;
                if (fast) {
                    u.where = block;
                    u.rows = kTUnitFast | row_word | (uint32_t)(row_slot0 + (r_lo - rec.imin) + 1) << 16;
                    u.cols = col_word | (uint32_t)(col_slot0 + (c - rec.jmin) + 1) << 16;
                } else {
                    u.where = (int32_t)s;
                    u.rows = row_word;
                    u.cols = col_word;
                }
                dst[q] = u;
            }
        }
    }
}

__device__ __forceinline__ bool mbar_try(void *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(void *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

template <typename OutT, int SLOTS, int CTAS>
__global__ void __launch_bounds__(kTileWarps * 32, CTAS)
render_tiles_kernel(Geo g, const float *__restrict__ box, const TUnit *__restrict__ units,
                    const SpotRec *__restrict__ spots, const uint32_t *__restrict__ edges, int edge_cap,
                    const int64_t *__restrict__ sat, const int *__restrict__ tile_start, int *__restrict__ next_tile,
                    const unsigned long long *__restrict__ wmax_bits, OutT *__restrict__ out, int accumulate) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *ring = reinterpret_cast<float *>(smem_raw);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + kTileStages * kTileStageBytes);
    unsigned long long *empty = full + kTileStages;
    int *tile_slot = reinterpret_cast<int *>(empty + kTileStages);
    uint4 *meta = reinterpret_cast<uint4 *>(smem_raw + kTileStages * kTileStageBytes + 128 + warp * kTileMetaBytes);

    const int slots = SLOTS ? SLOTS : g.slots;
    const uint32_t row_bytes = (uint32_t)slots * 4u;
    const int frame_tiles = g.nti * g.ntj, n_tiles = g.frames * frame_tiles;
    const long long table_entries = (long long)g.modulus * g.modulus * g.slots * g.slots;
    if (threadIdx.x == 0) {
        for (int st = 0; st < kTileStages; ++st) {
            mbar_init(&full[st], 1);
            mbar_init(&empty[st], kTileWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tile_slot[0] = atomicAdd(next_tile, 1);
    }
    __syncthreads();
    int acc[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[k][q] = 0;
    uint32_t count = 0;          // units this CTA has gone through: stage = count % STAGES, phase = (count / STAGES) & 1
    int turn = 0;                // which of the two tile slots holds the current tile

    for (;;) {
        const int tile = tile_slot[turn];
        if (tile >= n_tiles) break;
        int upcoming = 0;
        if (threadIdx.x == 0) upcoming = atomicAdd(next_tile, 1);      // in flight while this tile is rendered
        const int frame = tile / frame_tiles, in_frame = tile - frame * frame_tiles;
        const int ti = in_frame / g.ntj, tj = in_frame - ti * g.ntj;
        const int row0 = ti * kTileRows + warp * 8, col0 = tj * kTileCols;       // this warp's band
        const int seg_begin = tile_start[tile * g.stripes], seg_end = tile_start[(tile + 1) * g.stripes];
        const int shift = strip_shift(seg_end - seg_begin, wmax_bits[frame], g.box_peak);
        const double scale = scalbn(1.0, shift), lsb = scalbn(1.0, -shift);

        for (int base = seg_begin; base < seg_end; base += 32) {
            const int nb = min(32, seg_end - base);
            __syncwarp();                                  // this warp has read the previous batch
            // lane u fetches unit u (all four warps fetch the same batch: the others' loads hit the L1), publishes it
            // for its own warp and keeps what the copy of that unit needs
            uint32_t my_rows = 0, my_cols = 0;
            int my_where = 0;
            if (lane < nb) {
                const uint4 *src = reinterpret_cast<const uint4 *>(units + base + lane);
                const uint4 head = __ldg(src);
                const double ws = __longlong_as_double(((long long)head.y << 32) | head.x);
                my_where = (int)head.z;
                my_rows = head.w;
                my_cols = __ldg(reinterpret_cast<const uint32_t *>(src + 1));
                meta[lane] = make_uint4(__float_as_uint((float)(ws * scale)), my_rows, my_cols, (uint32_t)my_where);
            }
            const uint32_t fast_mask = __ballot_sync(0xffffffffu, (my_rows & kTUnitFast) != 0);
            __syncwarp();

            // the copy of unit u of this batch, issued by warp (count + u) % 4 once the stage's previous unit has been
            // released by all four warps.  Gather units take a turn of the ring too (nothing is copied), which keeps
            // stage and phase a function of the unit count alone.
            auto stage_unit = [&](int u) {
                const uint32_t c = count + (uint32_t)u;
                if ((c & (kTileWarps - 1)) != (uint32_t)warp) return;
                const int st = (int)(c % kTileStages);
                const uint32_t round = c / kTileStages;
                if (lane == u) {
                    if (round > 0) while (!mbar_try(&empty[st], (round - 1) & 1u)) {}
                    if ((fast_mask >> u) & 1u) {
                        const uint32_t n_rows = (my_rows >> 8) & 0xffu, first = (my_rows >> 16) & 0x7fffu;
                        const uint32_t bytes = n_rows * row_bytes;
                        const char *src = reinterpret_cast<const char *>(box) +
                                          (((size_t)my_where * (size_t)slots + first) * (size_t)slots) * 4u;
                        mbar_expect_tx(&full[st], bytes);
                        tma_row(ring + st * (kTileStageBytes / 4), src, bytes, &full[st]);
                    } else {
                        mbar_arrive(&full[st]);            // nothing to copy: the stage is "full" at once
                    }
                }
            };
            for (int u = 0; u < min(nb, kTileStages - 1); ++u) stage_unit(u);
            for (int u = 0; u < nb; ++u) {
                if (u + kTileStages - 1 < nb) stage_unit(u + kTileStages - 1);
                const uint32_t c = count + (uint32_t)u;
                const int st = (int)(c % kTileStages);
                const uint4 m = meta[u];                   // {weight, rows word, columns word, where}
                const int r0 = (int)(m.y & 0xffu), n_rows = (int)((m.y >> 8) & 0xffu);
                // rows of the unit inside this warp's band: [a, b) of 0 .. 8
                const int a = max(r0 - warp * 8, 0), b = min(r0 + n_rows - warp * 8, 8);
                // Every warp waits for every unit's "full" phase, rows in its band or not: its release below must be
                // counted in THIS unit's "empty" phase, and this wait is what keeps a warp without work from running
                // more than a ring ahead (the copy of unit u is not issued before unit u - STAGES has been released
                // by all four).
                while (!mbar_try(&full[st], (c / kTileStages) & 1u)) {}
                if (b > a) {                               // warp uniform
                    const int c0 = (int)(m.z & 0xffu), n_cols = (int)((m.z >> 8) & 0xffu);
                    const int rot = (lane - c0) & 31;
                    const float ws = rot < n_cols ? __uint_as_float(m.x) : 0.0f;     // idle lanes add nothing
                    float v[8];
                    if (m.y & kTUnitFast) {
                        const int box_col = (int)((m.z >> 16) & 0xffu);
                        // stage row j is unit row j = tile row r0 + j: band row k is stage row 8 * warp + k - r0
                        const float *base_row = ring + st * (kTileStageBytes / 4) + (warp * 8 - r0) * slots +
                                                min(box_col + rot, slots - 1);
                        if (a == 0 && b == 8) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] = base_row[k * slots];
                        } else {
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] = (k >= a && k < b) ? base_row[k * slots] : 0.0f;
                        }
                    } else {
                        const uint32_t packed = (uint32_t)c0 << 20 | (uint32_t)a << 17 | (uint32_t)(b - a) << 13 |
                                                (uint32_t)n_cols << 7;
                        const Box8 g8 = gather_unit(packed, (int)m.w, lane, spots, edges, edge_cap, sat, table_entries,
                                                    row0, col0);
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] = g8.v[k];
                    }
                    int d[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) d[k] = to_fixed(v[k], ws);
                    const bool in_first = lane >= (c0 & 31);
                    switch (c0 >> 5) {                     // warp uniform
                    case 0: group_add<0>(acc, d, in_first); break;
                    case 1: group_add<1>(acc, d, in_first); break;
                    case 2: group_add<2>(acc, d, in_first); break;
                    default: group_add<3>(acc, d, in_first); break;
                    }
                }
                __syncwarp();                              // every lane of this warp is done with the stage
                if (lane == 0) mbar_arrive(&empty[st]);
            }
            count += (uint32_t)nb;
        }

        // ---- write this warp's band (coalesced 128-byte rows) and clear the accumulators for the next tile
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
        // hand the next tile to the other warps: the slot written here was last read two tiles ago
        if (threadIdx.x == 0) tile_slot[turn ^ 1] = upcoming;
        __syncthreads();
        turn ^= 1;
    }
}
