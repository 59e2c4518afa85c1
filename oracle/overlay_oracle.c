/*
 * CPU ORACLE (C part) -- TEST INFRASTRUCTURE ONLY, never linked into the product.
 *
 * Plain-C restatement of the per-spot PSF overlay of the reference
 * (/root/reference/src/scopyon/_epifm.py:98-126 radial_to_cartesian, :224-282
 * overlay_signal_) and of the per-pixel detector loops (:329-433, :1472-1484).
 * Two evaluations of the same pixel box sums are provided:
 *   orc_render_bruteforce  sums the table samples of every pixel directly, i.e. the
 *                          reference's algorithm (`signal[i0:i1, j0:j1].sum()`), used
 *                          as the compiled CPU baseline ("port") in bench.py and to
 *                          validate the summed-area form;
 *   orc_render_sat         the int64 summed-area-table form the CUDA path uses, with the
 *                          same table quantisation and the same accumulation order
 *                          (ascending spot index per pixel), so the GPU result must
 *                          match it BIT FOR BIT.
 * Pinned: tests/test_oracle_c.py checks both against the numpy oracle
 * (oracle/epifm_oracle.py), which itself equals the live reference bit for bit.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- table ------------------------------------------------------------------- */

/* T[a][b] = lerp(prof, min(sqrt(da^2+db^2), c)) in the sample-index domain; the same
 * expression, operation by operation, as table_sample() in scopyon_b200/csrc/psf.cu. */
static double table_sample(const double *prof, int n_radial, int da, int db) {
    const int c = n_radial - 1;
    double R = sqrt((double)(da * da + db * db));
    if (R > (double)c) R = (double)c;
    int k = (int)R;
    if (k > c - 1) k = c - 1;
    double t = R - (double)k;
    double lo = prof[k], hi = prof[k + 1];
    return lo + (hi - lo) * t;
}

void orc_table_from_radial(const double *prof, int n_radial, double *T) {
    const int c = n_radial - 1, side = 2 * c + 1;
    for (int a = 0; a < side; ++a)
        for (int b = 0; b < side; ++b) T[(size_t)a * side + b] = table_sample(prof, n_radial, a - c, b - c);
}

/* power-of-two scale: 2^(60 - ilogb(sum T)) */
double orc_table_scale(const double *T, int side) {
    double total = 0.0;
    for (size_t i = 0; i < (size_t)side * side; ++i) total += T[i];
    int e = (total > 0.0 && isfinite(total)) ? ilogb(total) : 0;
    return ldexp(1.0, 60 - e);
}

/* S[a][b] = sum_{a'<a, b'<b} llrint(T[a'][b'] * scale), (side+1)^2 entries */
void orc_sat_int64(const double *T, int side, double scale, int64_t *S) {
    const int pitch = side + 1;
    for (int b = 0; b < pitch; ++b) S[b] = 0;
    for (int a = 0; a < side; ++a) {
        int64_t run = 0;
        S[(size_t)(a + 1) * pitch] = 0;
        for (int b = 0; b < side; ++b) {
            run += llrint(T[(size_t)a * side + b] * scale);
            S[(size_t)(a + 1) * pitch + b + 1] = S[(size_t)a * pitch + b + 1] + run;
        }
    }
}

/* ---- pixel edges, _epifm.py:228-253 ---------------------------------------------- */

typedef struct {
    int n_w, n_h, n_radial, n_depth_keys;
    double pixel_length, resolution, depth_cutoff;
    double focal[3];
} orc_geometry;

/* returns number of edges written (pixels = edges - 1), first pixel index in *first */
static int edges_of(double xi, int n_pixel, double pl, int side, double res, int *first, int *edges, int cap) {
    const double sw = res * (double)(side - 1);
    const double ew = (double)n_pixel * pl;
    const double o = ew * 0.5 + xi - sw * 0.5;
    double fmin_ = floor(o / pl), fmax_ = ceil((ew * 0.5 + xi + sw * 0.5) / pl);
    if (!(fmin_ > -1.0)) fmin_ = -1.0;
    if (fmin_ > n_pixel + 1) fmin_ = n_pixel + 1;
    if (!(fmax_ > -1.0)) fmax_ = -1.0;
    if (fmax_ > n_pixel + 1) fmax_ = n_pixel + 1;
    int imin = (int)fmin_, imax = (int)fmax_;
    int lo = imin > 0 ? imin : 0, hi = imax < n_pixel ? imax : n_pixel;
    int n = hi - lo + 1;
    if (n <= 0) return 0;
    if (n > cap) n = cap;
    for (int k = 0; k < n; ++k) {
        double v = ceil(((double)(lo + k) * pl - o) / res);
        if (!(v > -1.0)) v = -1.0;
        if (v > side + 1) v = side + 1;
        int e = (int)v;
        if (k == 0 && e < 0) e = 0;
        if (k == n - 1 && e > side) e = side;
        if (e < 0) e = 0;
        if (e > side) e = side;
        edges[k] = e;
    }
    *first = lo;
    return n;
}

int orc_edges(double xi, int n_pixel, double pl, int side, double res, int *first, int *edges, int cap) {
    return edges_of(xi, n_pixel, pl, side, res, first, edges, cap);
}

static int depth_slot(const orc_geometry *g, double dz, const int32_t *slot_of_key) {
    dz = fabs(dz);
    int key;
    if (dz < g->depth_cutoff + g->resolution) {
        double q = dz / g->resolution;
        key = (int)q;
        if (key > g->n_depth_keys - 1) key = g->n_depth_keys - 1;
    } else {
        key = g->n_depth_keys;
    }
    return slot_of_key[key];
}

#define ORC_MAX_EDGES 4096

/* ---- summed-area form: what the GPU must equal bit for bit ------------------------ */
int orc_render_sat(const orc_geometry *g, int64_t n_spots, const double *depth, const double *x,
                   const double *y, const double *weight, const int64_t *sat, const double *inv_scale,
                   const int32_t *slot_of_key, double *expected) {
    const int side = 2 * (g->n_radial - 1) + 1, pitch = side + 1;
    int *left = (int *)malloc(sizeof(int) * ORC_MAX_EDGES * 2), *top = left + ORC_MAX_EDGES;
    int missing = 0;
    for (int64_t s = 0; s < n_spots; ++s) {
        const double w = weight[s];
        if (!(w > 0.0)) continue;
        const double xi = x[s] - g->focal[1], yi = y[s] - g->focal[2], dz = depth[s] - g->focal[0];
        if (!isfinite(xi) || !isfinite(yi) || !isfinite(dz)) continue;
        const int slot = depth_slot(g, dz, slot_of_key);
        if (slot < 0) { ++missing; continue; }
        int i0, j0;
        const int ni = edges_of(xi, g->n_w, g->pixel_length, side, g->resolution, &i0, left, ORC_MAX_EDGES);
        const int nj = edges_of(yi, g->n_h, g->pixel_length, side, g->resolution, &j0, top, ORC_MAX_EDGES);
        if (ni < 2 || nj < 2) continue;
        const int64_t *S = sat + (size_t)slot * pitch * pitch;
        const double ws = w * (g->resolution * g->resolution) * inv_scale[slot];
        for (int a = 0; a + 1 < ni; ++a)
            for (int b = 0; b + 1 < nj; ++b) {
                const int64_t box = S[(size_t)left[a + 1] * pitch + top[b + 1]] - S[(size_t)left[a + 1] * pitch + top[b]]
                                  - S[(size_t)left[a] * pitch + top[b + 1]] + S[(size_t)left[a] * pitch + top[b]];
                if (box > 0) {
                    double *px = &expected[(size_t)(i0 + a) * g->n_h + (j0 + b)];
                    *px = *px + (double)box * ws;
                }
            }
    }
    free(left);
    return missing;
}

/* ---- reference algorithm: direct slice sums over the fp64 table ------------------- */
/* tables: [n_slots][side][side] fp64.  Threads take spots dynamically (the reference's
 * Pool.map over array_split(particles), _epifm.py:1250-1260, in spirit); each spot's
 * footprint is summed into a private buffer and then added to the shared image. */
int orc_render_bruteforce(const orc_geometry *g, int64_t n_spots, const double *depth, const double *x,
                          const double *y, const double *weight, const double *tables,
                          const int32_t *slot_of_key, double *expected, int n_threads) {
    const int side = 2 * (g->n_radial - 1) + 1;
    const double unit_area = g->resolution * g->resolution;
    int missing = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel reduction(+ : missing)
    {
        int *left = (int *)malloc(sizeof(int) * ORC_MAX_EDGES * 2), *top = left + ORC_MAX_EDGES;
        size_t cap = 64 * 64;
        double *foot = (double *)malloc(sizeof(double) * cap);
#pragma omp for schedule(dynamic, 1)
        for (int64_t s = 0; s < n_spots; ++s) {
            const double w = weight[s];
            if (!(w > 0.0)) continue;
            const double xi = x[s] - g->focal[1], yi = y[s] - g->focal[2], dz = depth[s] - g->focal[0];
            const int slot = depth_slot(g, dz, slot_of_key);
            if (slot < 0) { ++missing; continue; }
            int i0, j0;
            const int ni = edges_of(xi, g->n_w, g->pixel_length, side, g->resolution, &i0, left, ORC_MAX_EDGES);
            const int nj = edges_of(yi, g->n_h, g->pixel_length, side, g->resolution, &j0, top, ORC_MAX_EDGES);
            if (ni < 2 || nj < 2) continue;
            const size_t need = (size_t)(ni - 1) * (nj - 1);
            if (need > cap) { cap = need; foot = (double *)realloc(foot, sizeof(double) * cap); }
            const double *T = tables + (size_t)slot * side * side;
            for (int a = 0; a + 1 < ni; ++a)
                for (int b = 0; b + 1 < nj; ++b) {
                    double sum = 0.0;
                    for (int p = left[a]; p < left[a + 1]; ++p) {
                        const double *row = T + (size_t)p * side;
                        double rs = 0.0;
                        for (int q = top[b]; q < top[b + 1]; ++q) rs += row[q];
                        sum += rs;
                    }
                    const double photons = sum * unit_area;
                    foot[(size_t)a * (nj - 1) + b] = photons > 0 ? photons * w : 0.0;
                }
#pragma omp critical
            for (int a = 0; a + 1 < ni; ++a)
                for (int b = 0; b + 1 < nj; ++b)
                    expected[(size_t)(i0 + a) * g->n_h + (j0 + b)] += foot[(size_t)a * (nj - 1) + b];
        }
        free(left);
        free(foot);
    }
    return missing;
}

/* ---- detector loops for the CPU baseline (statistically equivalent port) ----------- */
static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t *s) { return ((double)(splitmix64(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

static double poisson_draw(double lam, uint64_t *s) {
    if (!(lam > 0.0)) return 0.0;
    if (lam < 10.0) { /* numpy's product-of-uniforms method for small lam */
        const double enlam = exp(-lam);
        double prod = u01(s);
        long k = 0;
        while (prod > enlam) { prod *= u01(s); ++k; }
        return (double)k;
    }
    const double slam = sqrt(lam), loglam = log(lam), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2);
    for (;;) {
        const double U = u01(s) - 0.5, V = u01(s), us = 0.5 - fabs(U);
        const double k = floor((2 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1)) return k;
    }
}

/* CMOS / CCD frame: Poisson signal + (categorical | Gaussian) readout + ADC, per pixel
 * like _epifm.py:332-353, 418-433, 1472-1484.  cdf: cumulative read-noise table. */
void orc_detector_frame(int64_t n_pix, const double *photons, double qe, double background, int is_cmos,
                        const double *rn_values, const double *rn_cdf, int n_rn, double readout_sigma,
                        double fullwell, double adc0, int bit, uint64_t seed, double *adc, int n_threads) {
    const double adc_max = ldexp(1.0, bit) - 1.0, gain = fullwell / (ldexp(1.0, bit) - adc0);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_pix; ++i) {
        uint64_t s = seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(i + 1));
        const double E = qe * (photons[i] + background);
        const double signal = poisson_draw(E, &s);
        double noise;
        if (is_cmos) {
            const double u = u01(&s);
            int lo = 0, hi = n_rn - 1;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (rn_cdf[mid] > u) hi = mid; else lo = mid + 1; }
            noise = rn_values[lo];
        } else {
            const double r = sqrt(-2.0 * log(u01(&s)));
            noise = readout_sigma * r * cos(6.283185307179586 * u01(&s));
        }
        double v = signal + noise;
        if (v > fullwell) v = fullwell;
        v = v / gain + adc0;
        if (v > adc_max) v = adc_max;
        if (v < 0) v = 0;
        adc[i] = v;
    }
}

/* Brownian step for the CPU baseline, sampling.py:116-118 (one normal per coordinate). */
void orc_move_points(int64_t n, int ndim, double *coords /* [n][ndim] */, const double *sigma, uint64_t seed) {
    for (int64_t i = 0; i < n; ++i) {
        uint64_t s = seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(i + 1));
        for (int d = 0; d < ndim; ++d) {
            const double r = sqrt(-2.0 * log(u01(&s)));
            coords[i * ndim + d] += sigma[d] * r * cos(6.283185307179586 * u01(&s));
        }
    }
}
