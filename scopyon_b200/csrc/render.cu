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
#include "binning.cuh"

namespace {

constexpr int kTile = 16;             // pixels per tile edge
constexpr int kEdge = kTile + 1;      // pixel edges per tile edge
constexpr int kBatch = 8;             // spots staged per round = warps per CTA
constexpr int kThreads = kTile * kTile;
constexpr int kSortCap = 1024;        // spots per tile ordered in shared memory per chunk

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
                // a warp covers tile rows 2*warp and 2*warp+1: skip spots that miss both (warp uniform)
                if (m.r0 > 2 * warp + 1 || m.r0 + m.nrow <= 2 * warp) continue;
                const int rk = py - m.r0, rl = px - m.c0;
                if ((unsigned)rk < (unsigned)m.nrow && (unsigned)rl < (unsigned)m.ncol) {
                    const long long *c = &corners[cur][q][rk * kEdge + rl];
                    const long long box = c[kEdge + 1] - c[kEdge] - c[1] + c[0];
                    // box >= 0 (the table is non-negative, edges are monotone), and adding a zero
                    // leaves acc unchanged, so the reference's `if photons > 0` needs no branch
                    acc = __dadd_rn(acc, __dmul_rn((double)box, m.w));   // _epifm.py:280-282
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
    Geo g = make_geo(geom, kTile);
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
    Geo g = make_geo(geom, kTile);
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
            g, n_spots, d_depth, d_x, d_y, d_weight, d_inv_scale, d_slot_of_key, w.spots, w.tile_count, d_errors);
        spot_edges_kernel<<<scb_grid_for(n_spots * 2 * w.edge_cap, 256), 256, 0, s>>>(g, n_spots, w.spots, w.edges,
                                                                                     w.edge_cap);
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
