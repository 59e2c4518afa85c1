// Expected-image rendering: spot -> screen-tile binning (counting sort) and
// tile-per-CTA accumulation from summed-area-table corners.
//
// Reference: _EPIFMSimulator.get_molecule_plane + PointSpreadingFunction.overlay_signal_
// (/root/reference/src/scopyon/_epifm.py:1262-1264, 224-282).  The reference walks, per
// spot, the ~31x31 pixels under the 1999x1999-sample PSF table and sums a ~66x66 slice of
// the table for each.  Here
//   * the pixel-edge -> table-index arithmetic is evaluated with the same IEEE fp64
//     operations in the same order (explicit __d*_rn intrinsics: never contracted to FMA);
//   * each slice sum is S[i1][j1] - S[i0][j1] - S[i1][j0] + S[i0][j0] on the int64 SAT
//     built by scb_psf_sat_build -- exact, so results do not depend on tile shape;
//   * one CTA owns one 16x16-pixel tile: a thread owns a pixel and accumulates in a
//     register, in ascending spot order -> no atomics on the image, bitwise reproducible.
#include "scb_common.cuh"

namespace {

constexpr int kTile = 16;             // pixels per tile edge
constexpr int kEdge = kTile + 1;      // pixel edges per tile edge
constexpr int kBatch = 8;             // spots staged per round = warps per CTA
constexpr int kThreads = kTile * kTile;
constexpr int kSortCap = 1024;        // spots per tile ordered in shared memory per chunk

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
                    SpotRec *__restrict__ spots, uint16_t *__restrict__ edges, int edge_cap,
                    int *__restrict__ tile_count, int32_t *__restrict__ errors) {
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
    const double dz = fabs(__dsub_rn(depth[s], g.f0));
    if (w > 0.0 && isfinite(xi) && isfinite(yi) && isfinite(dz)) {   // _epifm.py:217-218
        // depth key, _epifm.py:76-84
        int key;
        if (dz < __dadd_rn(g.depth_cutoff, g.res)) {
            key = clamp_to_int(__ddiv_rn(dz, g.res), 0, g.n_depth_keys - 1);
        } else {
            key = g.n_depth_keys;  // frozen at the cutoff ("key -1")
        }
        int slot = slot_of_key[key];
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
                rec.w = w * (g.res * g.res) * inv_scale[slot];
                const int t0 = rec.imin / kTile, t1 = (rec.imax - 1) / kTile;
                const int u0 = rec.jmin / kTile, u1 = (rec.jmax - 1) / kTile;
                for (int ti = t0; ti <= t1; ++ti)
                    for (int tj = u0; tj <= u1; ++tj) atomicAdd(&tile_count[ti * g.ntj + tj], 1);
                // table sample index of every pixel edge the footprint touches (_epifm.py:236-253):
                // rows first, then columns; the render kernel only looks them up
                uint16_t *e = edges + (size_t)s * 2 * edge_cap;
                for (int i = rec.imin; i <= rec.imax; ++i)
                    e[i - rec.imin] = (uint16_t)edge_index(i, rec.imin, rec.imax, rec.ox, g);
                for (int j = rec.jmin; j <= rec.jmax; ++j) {   // columns: stored as SAT storage offsets
                    const int b = edge_index(j, rec.jmin, rec.jmax, rec.oy, g);
                    e[edge_cap + j - rec.jmin] = (uint16_t)((b % g.modulus) * g.blocks + b / g.modulus);
                }
            }
        }
    }
    spots[s] = rec;
}

// Exclusive scan of the tile census (one CTA; n_tiles is at most a few 10^5).
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int n_tiles, const int *__restrict__ tile_count, int *__restrict__ tile_start) {
    __shared__ int warp_tot[32];
    const int per = (n_tiles + 1023) / 1024;
    const int b0 = threadIdx.x * per;
    int run = 0;
    for (int i = 0; i < per; ++i)
        if (b0 + i < n_tiles) run += tile_count[b0 + i];
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
    for (int i = 0; i < per; ++i) {
        if (b0 + i < n_tiles) {
            tile_start[b0 + i] = base;
            base += tile_count[b0 + i];
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
    const int t0 = imin / kTile, t1 = (imax - 1) / kTile;
    const int u0 = jmin / kTile, u1 = (jmax - 1) / kTile;
    for (int ti = t0; ti <= t1; ++ti)
        for (int tj = u0; tj <= u1; ++tj) {
            const int tile = ti * g.ntj + tj;
            pair_spot[tile_start[tile] + atomicAdd(&tile_cursor[tile], 1)] = (int)s;
        }
}

struct __align__(16) StageMeta {
    int r0, nrow, c0, ncol;   // footprint rectangle inside the tile (pixels)
    double w;
    double pad;
};

// 8-byte asynchronous global -> shared copy (LDGSTS): the SAT corner gathers of a whole
// round are in flight at once and need no registers.
__device__ __forceinline__ void cp_async_8(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <typename OutT>
__global__ void __launch_bounds__(kThreads)
render_tiles_kernel(Geo g, const SpotRec *__restrict__ spots, const uint16_t *__restrict__ edges,
                    int edge_cap, const int *__restrict__ tile_start, const int *__restrict__ pair_spot,
                    const int64_t *__restrict__ sat, OutT *__restrict__ out, int accumulate) {
    __shared__ long long corners[2][kBatch][kEdge * kEdge];
    __shared__ StageMeta meta[2][kBatch];
    __shared__ int ids_raw[kSortCap], ids[kSortCap];

    const int tile = blockIdx.x;
    const int ti = tile / g.ntj, tj = tile - ti * g.ntj;
    const int row0 = ti * kTile, col0 = tj * kTile;
    const int py = threadIdx.x / kTile, px = threadIdx.x % kTile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int seg_begin = tile_start[tile], seg_end = tile_start[tile + 1];
    const size_t pitch = (size_t)g.pitch, table = (size_t)(g.side + 1) * g.pitch;

    double acc = 0.0;

    for (int chunk = seg_begin; chunk < seg_end; chunk += kSortCap) {
        const int n_chunk = min(kSortCap, seg_end - chunk);
        // ---- order the chunk by spot index (rank sort; indices are distinct)
        __syncthreads();
        for (int t = threadIdx.x; t < n_chunk; t += kThreads) ids_raw[t] = pair_spot[chunk + t];
        __syncthreads();
        for (int t = threadIdx.x; t < n_chunk; t += kThreads) {
            const int mine = ids_raw[t];
            int rank = 0;
            for (int u = 0; u < n_chunk; ++u) rank += (ids_raw[u] < mine);
            ids[rank] = mine;
        }
        __syncthreads();

        // Warp b stages spot base+b of a round: lane l owns column edge l, the row edge of
        // each footprint row is broadcast, and lane l copies corner (row k, col l) straight
        // into shared memory with cp.async.
        auto stage = [&](int buf, int base) {
            StageMeta m;
            m.r0 = 0; m.nrow = 0; m.c0 = 0; m.ncol = 0; m.w = 0.0; m.pad = 0.0;
            if (base + warp < n_chunk) {
                const int sid = ids[base + warp];
                const SpotRec rec = spots[sid];
                const int r_lo = max(rec.imin, row0), r_hi = min(rec.imax, row0 + kTile);
                const int c_lo = max(rec.jmin, col0), c_hi = min(rec.jmax, col0 + kTile);
                const int nrow = r_hi - r_lo, ncol = c_hi - c_lo;
                const uint16_t *e = edges + (size_t)sid * 2 * edge_cap;
                const int my_row = (lane <= nrow) ? (int)e[r_lo - rec.imin + lane] : 0;
                const int my_col = (lane <= ncol) ? (int)e[edge_cap + c_lo - rec.jmin + lane] : 0;
                const int64_t *S = sat + (size_t)rec.slot * table + my_col;
                long long *dst = &corners[buf][warp][lane];
                for (int k = 0; k <= nrow; ++k) {
                    const int a = __shfl_sync(0xffffffffu, my_row, k);
                    if (lane <= ncol) cp_async_8(dst + k * kEdge, S + (size_t)a * pitch);
                }
                m.r0 = r_lo - row0; m.nrow = nrow; m.c0 = c_lo - col0; m.ncol = ncol;
                m.w = rec.w;
            }
            cp_async_commit();
            if (lane == 0) meta[buf][warp] = m;
        };

        stage(0, 0);
        int cur = 0;
        for (int base = 0; base < n_chunk; base += kBatch, cur ^= 1) {
            const bool more = base + kBatch < n_chunk;
            if (more) {
                stage(cur ^ 1, base + kBatch);   // next round's gathers fly during this round's math
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            // ---- accumulate: thread (py, px) owns one pixel
#pragma unroll
            for (int q = 0; q < kBatch; ++q) {
                const StageMeta m = meta[cur][q];
                const int rk = py - m.r0, rl = px - m.c0;
                if ((unsigned)rk < (unsigned)m.nrow && (unsigned)rl < (unsigned)m.ncol) {
                    const long long *c = &corners[cur][q][rk * kEdge + rl];
                    const long long box = c[kEdge + 1] - c[kEdge] - c[1] + c[0];
                    if (box > 0) acc = __dadd_rn(acc, __dmul_rn((double)box, m.w));   // _epifm.py:280-282
                }
            }
            __syncthreads();
        }
    }

    const int i = row0 + py, j = col0 + px;
    if (i < g.n_w && j < g.n_h) {
        const size_t o = (size_t)i * g.n_h + j;
        if (accumulate) out[o] = (OutT)((double)out[o] + acc);
        else out[o] = (OutT)acc;
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

Geo make_geo(const scb_geometry *geom) {
    Geo g;
    g.n_w = geom->n_w; g.n_h = geom->n_h;
    g.nti = (geom->n_w + kTile - 1) / kTile;
    g.ntj = (geom->n_h + kTile - 1) / kTile;
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
    int64_t per_axis = (int64_t)((rows + kTile - 2) / kTile) + 1;
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

// ---- measurement hook -----------------------------------------------------------
struct ProfilePool {
    cudaEvent_t *start = nullptr, *stop = nullptr;
    int capacity = 0, used = 0;
    bool enabled = false;
} g_profile;

}  // namespace

extern "C" int scb_profile_begin(int max_launches) {
    SCB_REQUIRE(max_launches > 0 && max_launches <= (1 << 20), SCB_E_INVALID, "scb_profile_begin: max_launches=%d", max_launches);
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

extern "C" size_t scb_render_workspace_bytes(const scb_geometry *geom, int64_t n_spots) {
    if (check_geometry(geom) != 0 || n_spots < 0) return 0;
    Geo g = make_geo(geom);
    return carve(g, n_spots, nullptr).bytes;
}

extern "C" int scb_render_expected(const scb_geometry *geom, int64_t n_spots, const double *d_depth,
                                   const double *d_x, const double *d_y, const double *d_weight,
                                   const int64_t *d_sat, const double *d_inv_scale,
                                   const int32_t *d_slot_of_key, void *d_out, int out_type,
                                   int accumulate, void *d_workspace, size_t workspace_bytes,
                                   int32_t *d_errors, void *stream) {
    int rc = check_geometry(geom);
    if (rc) return rc;
    SCB_REQUIRE(n_spots >= 0 && n_spots < (int64_t)1 << 31, SCB_E_INVALID, "n_spots=%lld", (long long)n_spots);
    SCB_REQUIRE(d_out && d_workspace && d_errors, SCB_E_NULL, "scb_render_expected: NULL out/workspace/errors");
    SCB_REQUIRE(n_spots == 0 || (d_depth && d_x && d_y && d_weight && d_sat && d_inv_scale && d_slot_of_key),
                SCB_E_NULL, "scb_render_expected: NULL spot/table pointer");
    SCB_REQUIRE(out_type == SCB_F32 || out_type == SCB_F64, SCB_E_INVALID, "out_type=%d", out_type);
    Geo g = make_geo(geom);
    Workspace w = carve(g, n_spots, d_workspace);
    SCB_REQUIRE(workspace_bytes >= w.bytes, SCB_E_WORKSPACE, "scb_render_expected: workspace %zu < %zu",
                workspace_bytes, w.bytes);
    cudaStream_t s = (cudaStream_t)stream;
    const int n_tiles = g.nti * g.ntj;
    {
        // The corner gathers use 8 bytes of every sector they touch; ask the L2 to fetch 32-byte
        // sectors from HBM instead of 128-byte lines (measured 4 sectors per miss by default).
        // Purely a performance hint, set once.
        static bool hinted = false;
        if (!hinted) {
            cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
            cudaGetLastError();
            hinted = true;
        }
    }
    // tile_count and tile_cursor are adjacent 256-aligned blocks: clear both
    SCB_CUDA(cudaMemsetAsync(w.tile_count, 0, (size_t)((char *)w.tile_start - (char *)w.tile_count), s));
    if (n_spots > 0) {
        spot_prepare_kernel<<<scb_grid_for(n_spots, 256), 256, 0, s>>>(
            g, n_spots, d_depth, d_x, d_y, d_weight, d_inv_scale, d_slot_of_key, w.spots, w.edges, w.edge_cap,
            w.tile_count, d_errors);
    }
    tile_scan_kernel<<<1, 1024, 0, s>>>(n_tiles, w.tile_count, w.tile_start);
    if (n_spots > 0) {
        tile_fill_kernel<<<scb_grid_for(n_spots, 256), 256, 0, s>>>(g, n_spots, w.spots, w.tile_start,
                                                                   w.tile_cursor, w.pair_spot);
    }
    const bool timed = g_profile.enabled && g_profile.used < g_profile.capacity;
    if (timed) cudaEventRecord(g_profile.start[g_profile.used], s);
    if (out_type == SCB_F32)
        render_tiles_kernel<float><<<n_tiles, kThreads, 0, s>>>(g, w.spots, w.edges, w.edge_cap, w.tile_start,
                                                               w.pair_spot, d_sat, (float *)d_out, accumulate);
    else
        render_tiles_kernel<double><<<n_tiles, kThreads, 0, s>>>(g, w.spots, w.edges, w.edge_cap, w.tile_start,
                                                                w.pair_spot, d_sat, (double *)d_out, accumulate);
    if (timed) cudaEventRecord(g_profile.stop[g_profile.used++], s);
    SCB_CUDA_LAUNCH_CHECK("scb_render_expected");
    return 0;
}
