// Spot detection on the device: the step downstream of image formation.
//
// Reference: /root/reference/src/scopyon/analysis/spot_detection.py
//   blob_detection (:14-47)  = skimage.feature.blob_log: a Laplacian-of-Gaussian scale space
//       (scipy.ndimage.gaussian_laplace per sigma, times -sigma^2), 3x3x3 local maxima above a
//       threshold, then pruning of overlapping blobs;
//   spot_detection (:110-174) = per blob: crop a ROI, fit a background plane to the ROI border
//       (:60-76), subtract it, fit a1*exp(-((x-a2)^2+(y-a3)^2)/a4) by least squares (:95-103).
//
// Three kernels groups:
//   log_axis0_kernel / log_axis1_kernel  separable scale space, one CTA tile per (scale, tile), the
//       tile and its reflected halo staged in shared memory.  The sums are formed exactly as
//       scipy's correlate1d forms them (centre tap first, then symmetric pairs from the far end
//       inwards, separate multiply and add) so the cube matches the CPU library bit for bit and
//       the peak set cannot differ by rounding.
//   log_peaks_kernel   thread per voxel: above threshold and no greater 3x3x3 neighbour (edges
//       replicated); hits are appended through one atomic counter.
//   spot_fit_kernel    warp per blob: ROI in shared memory, normal equations of the border plane,
//       then the reference's optimiser itself -- scipy's trust-region least squares, restated
//       step for step on warp-reduced J^T J -- so the fit stops where the reference's stops.
#include "scb_common.cuh"

namespace {

// scipy.ndimage 'reflect' boundary: d c b a | a b c d | d c b a
__device__ __forceinline__ int reflect_index(int i, int n) {
    if ((unsigned)i < (unsigned)n) return i;
    const int period = 2 * n;
    int m = i % period;
    if (m < 0) m += period;
    return m < n ? m : period - 1 - m;
}

constexpr int kACols = 32, kARows = 64;    // axis-0 pass: outputs per CTA
constexpr int kBCols = 128, kBRows = 8;    // axis-1 pass

__host__ __device__ inline size_t axis0_smem(int R) { return sizeof(double) * (2 * (size_t)(R + 1) + (size_t)(kARows + 2 * R) * kACols); }
__host__ __device__ inline size_t axis1_smem(int R) { return sizeof(double) * (2 * (size_t)(R + 1) + 2 * (size_t)kBRows * (kBCols + 2 * R)); }

// Along axis 0 (stride n_h): t0 = G (*) image, t2 = G'' (*) image for every scale.
__global__ void __launch_bounds__(256)
log_axis0_kernel(const double *__restrict__ image, int H, int W, const int32_t *__restrict__ radius,
                 const double *__restrict__ weights, int pitch, double *__restrict__ t2, double *__restrict__ t0) {
    extern __shared__ double sm[];
    const int s = blockIdx.z, R = radius[s];
    double *w0 = sm, *w2 = sm + (R + 1), *tile = sm + 2 * (R + 1);
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int k = tid; k <= R; k += 256) {
        w0[k] = weights[(size_t)(2 * s) * pitch + k];
        w2[k] = weights[(size_t)(2 * s + 1) * pitch + k];
    }
    const int r0 = blockIdx.y * kARows, col = blockIdx.x * kACols + threadIdx.x;
    const int rows = kARows + 2 * R;
    for (int r = threadIdx.y; r < rows; r += 8) {
        const int src = reflect_index(r0 - R + r, H);
        tile[r * kACols + threadIdx.x] = col < W ? image[(size_t)src * W + col] : 0.0;
    }
    __syncthreads();
    if (col >= W) return;
    const size_t plane = (size_t)s * H * W;
    for (int o = threadIdx.y; o < kARows && r0 + o < H; o += 8) {
        const double *c = tile + (o + R) * kACols + threadIdx.x;
        double a0 = __dmul_rn(c[0], w0[0]), a2 = __dmul_rn(c[0], w2[0]);
        for (int k = R; k >= 1; --k) {
            const double pair = __dadd_rn(c[-k * kACols], c[k * kACols]);
            a0 = __dadd_rn(a0, __dmul_rn(pair, w0[k]));
            a2 = __dadd_rn(a2, __dmul_rn(pair, w2[k]));
        }
        const size_t at = plane + (size_t)(r0 + o) * W + col;
        t0[at] = a0;
        t2[at] = a2;
    }
}

// Along axis 1 (contiguous): cube = -(G (*) t2 + G'' (*) t0) * sigma^2.
__global__ void __launch_bounds__(256)
log_axis1_kernel(const double *__restrict__ t2, const double *__restrict__ t0, int H, int W,
                 const int32_t *__restrict__ radius, const double *__restrict__ weights, int pitch,
                 const double *__restrict__ sigma2, double *__restrict__ cube) {
    extern __shared__ double sm[];
    const int s = blockIdx.z, R = radius[s];
    const int span = kBCols + 2 * R;
    double *w0 = sm, *w2 = sm + (R + 1), *a = sm + 2 * (R + 1), *b = a + kBRows * span;
    const int tid = threadIdx.x;
    for (int k = tid; k <= R; k += 256) {
        w0[k] = weights[(size_t)(2 * s) * pitch + k];
        w2[k] = weights[(size_t)(2 * s + 1) * pitch + k];
    }
    const int r0 = blockIdx.y * kBRows, c0 = blockIdx.x * kBCols;
    const size_t plane = (size_t)s * H * W;
    for (int idx = tid; idx < kBRows * span; idx += 256) {
        const int r = idx / span, c = idx - r * span;
        if (r0 + r < H) {
            const size_t at = plane + (size_t)(r0 + r) * W + reflect_index(c0 - R + c, W);
            a[idx] = t2[at];
            b[idx] = t0[at];
        }
    }
    __syncthreads();
    const int tx = tid & (kBCols - 1), col = c0 + tx;
    if (col >= W) return;
    const double scale = sigma2[s];
    for (int r = tid >> 7; r < kBRows && r0 + r < H; r += 2) {
        const double *pa = a + r * span + tx + R, *pb = b + r * span + tx + R;
        double u = __dmul_rn(pa[0], w0[0]), v = __dmul_rn(pb[0], w2[0]);
        for (int k = R; k >= 1; --k) {
            u = __dadd_rn(u, __dmul_rn(__dadd_rn(pa[-k], pa[k]), w0[k]));
            v = __dadd_rn(v, __dmul_rn(__dadd_rn(pb[-k], pb[k]), w2[k]));
        }
        cube[plane + (size_t)(r0 + r) * W + col] = __dmul_rn(-__dadd_rn(u, v), scale);
    }
}

// Voxels above the threshold that no 3x3x3 neighbour exceeds (skimage peak_local_max with
// footprint ones((3,3,3)), exclude_border=False; maximum_filter mode 'nearest').
__global__ void __launch_bounds__(256)
log_peaks_kernel(const double *__restrict__ cube, int H, int W, int S, double threshold, int32_t *__restrict__ peaks,
                 double *__restrict__ values, int64_t capacity, unsigned long long *__restrict__ count) {
    const int64_t total = (int64_t)S * H * W;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const double v = cube[idx];
        if (!(v > threshold)) continue;
        const int j = (int)(idx % W);
        const int i = (int)((idx / W) % H);
        const int s = (int)(idx / ((int64_t)W * H));
        bool peak = true;
        for (int ds = -1; ds <= 1 && peak; ++ds) {
            const int ss = min(max(s + ds, 0), S - 1);
            for (int di = -1; di <= 1 && peak; ++di) {
                const int ii = min(max(i + di, 0), H - 1);
                const double *row = cube + ((size_t)ss * H + ii) * W;
                const double l = row[max(j - 1, 0)], c = row[j], r = row[min(j + 1, W - 1)];
                peak = !(l > v) && !(c > v) && !(r > v);
            }
        }
        if (!peak) continue;
        const unsigned long long pos = atomicAdd(count, 1ull);
        if ((int64_t)pos < capacity) {
            peaks[3 * pos + 0] = i;
            peaks[3 * pos + 1] = j;
            peaks[3 * pos + 2] = s;
            values[pos] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
constexpr int kRoiMax = 32;     // ROI side limit: roi_size <= 15
constexpr int kFitWarps = 4;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// Solve the symmetric positive semi-definite 3x3 system A p = b; a direction the data do not
// determine (one-row ROI: every i is 0) keeps p = 0, where least_squares leaves it as well.
__device__ void solve_plane(double A[3][3], double b[3], double p[3]) {
    bool used[3] = {false, false, false};
    double scale = fmax(fmax(A[0][0], A[1][1]), A[2][2]);
    p[0] = p[1] = p[2] = 0.0;
    int order[3], n = 0;
    for (int step = 0; step < 3; ++step) {
        int piv = -1;
        double best = scale * 1e-13;
        for (int k = 0; k < 3; ++k)
            if (!used[k] && A[k][k] > best) { best = A[k][k]; piv = k; }
        if (piv < 0) break;
        used[piv] = true;
        order[n++] = piv;
        const double inv = 1.0 / A[piv][piv];
        for (int r = 0; r < 3; ++r) {
            if (r == piv || used[r]) continue;
            const double f = A[r][piv] * inv;
            for (int c = 0; c < 3; ++c) A[r][c] -= f * A[piv][c];
            b[r] -= f * b[piv];
        }
    }
    for (int t = n - 1; t >= 0; --t) {      // back substitution over the pivots, last first
        const int k = order[t];
        double v = b[k];
        for (int u = t + 1; u < n; ++u) v -= A[k][order[u]] * p[order[u]];
        p[k] = v / A[k][k];
    }
}

// Eigen-decomposition of a symmetric 4x4 matrix by cyclic Jacobi rotations: A = V diag(w) V^T,
// eigenvalues in decreasing order.  With A = J^T J these are the squared singular values and the
// right singular vectors of J that scipy's trust-region solver takes from an SVD.
__device__ void eig4(const double A[4][4], double w[4], double V[4][4]) {
    double a[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[i][j] = A[i][j]; V[i][j] = i == j ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 16; ++sweep) {
        double off = 0, diag = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            diag += a[i][i] * a[i][i];
#pragma unroll
            for (int j = i + 1; j < 4; ++j) off += a[i][j] * a[i][j];
        }
        if (!(off > 1e-34 * diag)) break;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                const double apq = a[p][q];
                if (apq == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
                const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double kp = a[k][p], kq = a[k][q];
                    a[k][p] = c * kp - sn * kq;
                    a[k][q] = sn * kp + c * kq;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double pk = a[p][k], qk = a[q][k];
                    a[p][k] = c * pk - sn * qk;
                    a[q][k] = sn * pk + c * qk;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double kp = V[k][p], kq = V[k][q];
                    V[k][p] = c * kp - sn * kq;
                    V[k][q] = sn * kp + c * kq;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = a[i][i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {          // selection sort, decreasing
#pragma unroll
        for (int j = i + 1; j < 4; ++j) {
            if (w[j] > w[i]) {
                const double tw = w[i]; w[i] = w[j]; w[j] = tw;
#pragma unroll
                for (int k = 0; k < 4; ++k) { const double tv = V[k][i]; V[k][i] = V[k][j]; V[k][j] = tv; }
            }
        }
    }
}

__device__ __forceinline__ double norm4(const double v[4]) {
    return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
}

// The trust-region subproblem min |J p + f| subject to |p| <= Delta, solved the way
// scipy.optimize._lsq.common.solve_lsq_trust_region does (a Newton iteration on the secular
// equation, at most ten steps, 1 % tolerance on |p| = Delta), from s^2 = eigenvalues of J^T J,
// V and suf = V^T g (= s * U^T f).  alpha is the Levenberg-Marquardt parameter carried between
// calls.
__device__ void trust_region_step(const double s2[4], const double V[4][4], const double suf[4], int n_residuals,
                                  double Delta, double &alpha, double p[4]) {
    const double eps = 2.220446049250313e-16;
    const bool full_rank = sqrt(fmax(s2[3], 0.0)) > eps * n_residuals * sqrt(fmax(s2[0], 0.0));
    double y[4];
    if (full_rank) {
#pragma unroll
        for (int k = 0; k < 4; ++k) y[k] = suf[k] / s2[k];
        if (norm4(y) <= Delta) {
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = -(V[i][0] * y[0] + V[i][1] * y[1] + V[i][2] * y[2] + V[i][3] * y[3]);
            alpha = 0.0;
            return;
        }
    }
    auto phi = [&](double a, double &slope) {
        double sum2 = 0, sum3 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double den = s2[k] + a, r = suf[k] / den;
            sum2 += r * r;
            sum3 += suf[k] * suf[k] / (den * den * den);
        }
        const double pn = sqrt(sum2);
        slope = -sum3 / pn;
        return pn - Delta;
    };
    double upper = norm4(suf) / Delta, lower = 0.0, slope;
    if (full_rank) {
        const double at0 = phi(0.0, slope);
        lower = -at0 / slope;
    }
    double a = (!full_rank && alpha == 0.0) ? fmax(0.001 * upper, sqrt(lower * upper)) : alpha;
    for (int it = 0; it < 10; ++it) {
        if (a < lower || a > upper) a = fmax(0.001 * upper, sqrt(lower * upper));
        const double value = phi(a, slope);
        if (value < 0) upper = a;
        const double ratio = value / slope;
        lower = fmax(lower, a - ratio);
        a -= (value + Delta) * ratio / Delta;
        if (fabs(value) < 0.01 * Delta) break;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) y[k] = suf[k] / (s2[k] + a);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = -(V[i][0] * y[0] + V[i][1] * y[1] + V[i][2] * y[2] + V[i][3] * y[3]);
    const double scale = Delta / norm4(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] *= scale;
    alpha = a;
}

struct FitSums {
    double cost;        // sum r^2
    double A[4][4];     // J^T J
    double g[4];        // J^T r
    double total;       // sum of the model (the reference's "intensity")
};

// residuals r = a1 exp(-((i-a2)^2 + (j-a3)^2)/a4) - data over the m x n ROI (spot_detection.py:84-86, 99)
__device__ void fit_eval(const double *data, int n, int cnt, const double p[4], int lane, FitSums &out) {
    double cost = 0, total = 0, A[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, g[4] = {0, 0, 0, 0};
    const double inv = 1.0 / p[3];
    for (int t = lane; t < cnt; t += 32) {
        const int i = t / n, j = t - i * n;
        const double dx = (double)i - p[1], dy = (double)j - p[2];
        const double d2 = dx * dx + dy * dy;
        const double e = exp(-d2 * inv);
        const double f = p[0] * e;
        const double r = f - data[t];
        const double J[4] = {e, 2.0 * f * dx * inv, 2.0 * f * dy * inv, f * d2 * inv * inv};
        cost += r * r;
        total += f;
        int q = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            g[a] += J[a] * r;
#pragma unroll
            for (int b = 0; b <= a; ++b) A[q++] += J[a] * J[b];
        }
    }
    out.cost = warp_sum(cost);
    out.total = warp_sum(total);
    int q = 0;
    for (int a = 0; a < 4; ++a) {
        out.g[a] = warp_sum(g[a]);
        for (int b = 0; b <= a; ++b) {
            const double v = warp_sum(A[q++]);
            out.A[a][b] = v;
            out.A[b][a] = v;
        }
    }
}

enum { FIT_OK = 0, FIT_LOW_SIGNAL = 1, FIT_NO_BACKGROUND = 2, FIT_NOT_CONVERGED = 3, FIT_OUTSIDE = 4 };

__global__ void __launch_bounds__(kFitWarps * 32)
spot_fit_kernel(const double *__restrict__ image, int H, int W, const double *__restrict__ blobs, int blob_stride,
                int n_blobs, double roi_size, int max_iter, double *__restrict__ spots, int32_t *__restrict__ status) {
    __shared__ double s_roi[kFitWarps][kRoiMax * kRoiMax];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blob = blockIdx.x * kFitWarps + warp;
    if (blob >= n_blobs) return;
    double *roi = s_roi[warp];
    double *out = spots + (size_t)blob * 6;
    const double x = blobs[(size_t)blob * blob_stride], y = blobs[(size_t)blob * blob_stride + 1];
    // spot_detection.py:113-117: int() truncates towards zero, like the cast
    int x0 = (int)(x - roi_size), x1 = (int)(x + roi_size) + 1;
    int y0 = (int)(y - roi_size), y1 = (int)(y + roi_size) + 1;
    x0 = max(0, x0); x1 = min(H, x1);
    y0 = max(0, y0); y1 = min(W, y1);
    const int m = x1 - x0, n = y1 - y0;
    int state = FIT_OK;
    if (m <= 0 || n <= 0 || m > kRoiMax || n > kRoiMax) state = FIT_LOW_SIGNAL;      // empty crop: sum 0
    const int cnt = state == FIT_OK ? m * n : 0;
    double sum = 0;
    for (int t = lane; t < cnt; t += 32) {
        const int i = t / n, j = t - i * n;
        const double v = image[(size_t)(x0 + i) * W + (y0 + j)];
        roi[t] = v;
        sum += v;
    }
    sum = warp_sum(sum);
    __syncwarp();
    if (state == FIT_OK && !(sum > 0.0)) state = FIT_LOW_SIGNAL;                      // :119-120
    if (state != FIT_OK) {
        if (lane == 0) status[blob] = state;
        return;
    }

    // background plane through the four border lines (corners counted twice), :60-76
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const int border = 2 * n + 2 * m;
    for (int t = lane; t < border; t += 32) {
        int i, j;
        if (t < n) { i = 0; j = t; }
        else if (t < 2 * n) { i = m - 1; j = t - n; }
        else if (t < 2 * n + m) { i = t - 2 * n; j = 0; }
        else { i = t - 2 * n - m; j = n - 1; }
        const double v = roi[i * n + j], di = i, dj = j;
        acc[0] += di * di; acc[1] += di * dj; acc[2] += di; acc[3] += dj * dj; acc[4] += dj; acc[5] += 1.0;
        acc[6] += di * v; acc[7] += dj * v; acc[8] += v;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = warp_sum(acc[k]);
    double A3[3][3] = {{acc[0], acc[1], acc[2]}, {acc[1], acc[3], acc[4]}, {acc[2], acc[4], acc[5]}};
    double b3[3] = {acc[6], acc[7], acc[8]}, plane[3];
    solve_plane(A3, b3, plane);
    if (!isfinite(plane[0]) || !isfinite(plane[1]) || !isfinite(plane[2])) {
        if (lane == 0) status[blob] = FIT_NO_BACKGROUND;
        return;
    }

    // subtract the plane (:78-82), centre of mass of what is left (:88-93)
    double bg = 0, total = 0, sx = 0, sy = 0;
    for (int t = lane; t < cnt; t += 32) {
        const int i = t / n, j = t - i * n;
        const double level = __dadd_rn(__dadd_rn(__dmul_rn((double)i, plane[0]), __dmul_rn((double)j, plane[1])), plane[2]);
        const double v = roi[t] - level;
        roi[t] = v;
        bg += level;
        total += v;
        sx += i * v;
        sy += j * v;
    }
    bg = warp_sum(bg);
    total = warp_sum(total);
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    __syncwarp();

    // scipy.optimize.least_squares(..., method='trf') without bounds, the reference's call (:99-100),
    // step for step: trust radius |x0| to start with, exact subproblem solution, the same radius
    // update and the same ftol = xtol = gtol = 1e-8 stopping rules, at most 100 * n evaluations.
    // The Jacobian is analytic where scipy takes forward differences.  Start (:96-98):
    // (255, centre of mass, roi_size / 2).
    double p[4] = {255.0, sx / total, sy / total, roi_size / 2};
    FitSums cur, trial;
    fit_eval(roi, n, cnt, p, lane, cur);
    const double tol = 1e-8;
    double Delta = norm4(p), alpha = 0.0;
    if (Delta == 0.0) Delta = 1.0;
    int nfev = 1, verdict = 0;          // 0 = running; scipy's status 1..4 = converged; -1 = failed
    if (!isfinite(cur.cost)) verdict = -1;      // least_squares raises: "Residuals are not finite in the initial point"
    while (verdict == 0) {
        if (fmax(fmax(fabs(cur.g[0]), fabs(cur.g[1])), fmax(fabs(cur.g[2]), fabs(cur.g[3]))) < tol) { verdict = 1; break; }
        if (nfev >= max_iter) { verdict = -1; break; }
        double s2[4], V[4][4], suf[4];
        eig4(cur.A, s2, V);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s2[k] = fmax(s2[k], 0.0);
            suf[k] = V[0][k] * cur.g[0] + V[1][k] * cur.g[1] + V[2][k] * cur.g[2] + V[3][k] * cur.g[3];
        }
        double actual = -1.0, q[4];
        while (actual <= 0.0 && nfev < max_iter) {
            double h[4];
            trust_region_step(s2, V, suf, cnt, Delta, alpha, h);
            double quad = 0, lin = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                lin += cur.g[a] * h[a];
#pragma unroll
                for (int b = 0; b < 4; ++b) quad += h[a] * cur.A[a][b] * h[b];
                q[a] = p[a] + h[a];
            }
            const double predicted = -(0.5 * quad + lin);
            const double step = norm4(h);
            fit_eval(roi, n, cnt, q, lane, trial);
            ++nfev;
            if (!isfinite(trial.cost)) {
                Delta = 0.25 * step;
                continue;
            }
            actual = 0.5 * (cur.cost - trial.cost);
            double ratio;
            if (predicted > 0) ratio = actual / predicted;
            else ratio = (predicted == 0.0 && actual == 0.0) ? 1.0 : 0.0;
            double next = Delta;
            if (ratio < 0.25) next = 0.25 * step;
            else if (ratio > 0.75 && step > 0.95 * Delta) next = 2.0 * Delta;
            const bool f_ok = actual < tol * (0.5 * cur.cost) && ratio > 0.25;
            const bool x_ok = step < tol * (tol + norm4(p));
            if (f_ok && x_ok) verdict = 4;
            else if (f_ok) verdict = 2;
            else if (x_ok) verdict = 3;
            if (verdict) break;
            alpha *= Delta / next;
            Delta = next;
        }
        if (actual > 0.0) {             // the step is taken even when it is the one that ends the search
#pragma unroll
            for (int a = 0; a < 4; ++a) p[a] = q[a];
            cur = trial;
        }
    }
    if (verdict < 0) {
        if (lane == 0) status[blob] = FIT_NOT_CONVERGED;
        return;
    }
    if (!(p[1] >= 0.0 && p[1] < (double)m && p[2] >= 0.0 && p[2] < (double)n)) {      // :129-131
        if (lane == 0) status[blob] = FIT_OUTSIDE;
        return;
    }
    if (lane == 0) {
        out[0] = p[1] + x0;       // :133-136
        out[1] = p[2] + y0;
        out[2] = cur.total;
        out[3] = bg;
        out[4] = p[0];
        out[5] = p[3];
        status[blob] = FIT_OK;
    }
}

}  // namespace

extern "C" size_t scb_log_workspace_bytes(int n_w, int n_h, int n_sigma) {
    return 2 * sizeof(double) * (size_t)n_sigma * n_w * n_h;
}

extern "C" int scb_log_max_radius(void) {
    int R = 0;
    while (axis0_smem(R + 1) <= 200 * 1024 && axis1_smem(R + 1) <= 200 * 1024) ++R;
    return R;
}

extern "C" int scb_log_scale_space(int n_w, int n_h, int n_sigma, const double *d_image, const int32_t *d_radius,
                                   int max_radius, const double *d_weights, int weight_pitch, const double *d_sigma2,
                                   double *d_cube, void *d_workspace, size_t workspace_bytes, void *stream) {
    SCB_REQUIRE(d_image && d_radius && d_weights && d_sigma2 && d_cube && d_workspace, SCB_E_NULL,
                "scb_log_scale_space: NULL pointer");
    SCB_REQUIRE(n_w > 0 && n_h > 0 && n_sigma > 0 && n_sigma <= 65535, SCB_E_INVALID,
                "scb_log_scale_space: image %dx%d, %d scales", n_w, n_h, n_sigma);
    SCB_REQUIRE(max_radius >= 0 && max_radius < weight_pitch, SCB_E_INVALID, "max_radius=%d, weight_pitch=%d", max_radius,
                weight_pitch);
    SCB_REQUIRE(max_radius <= scb_log_max_radius(), SCB_E_INVALID,
                "scb_log_scale_space: kernel radius %d exceeds %d (sigma too large for the shared-memory tile)", max_radius,
                scb_log_max_radius());
    SCB_REQUIRE(workspace_bytes >= scb_log_workspace_bytes(n_w, n_h, n_sigma), SCB_E_WORKSPACE,
                "scb_log_scale_space: workspace %zu < %zu", workspace_bytes, scb_log_workspace_bytes(n_w, n_h, n_sigma));
    cudaStream_t s = (cudaStream_t)stream;
    double *t2 = (double *)d_workspace, *t0 = t2 + (size_t)n_sigma * n_w * n_h;
    const size_t smem0 = axis0_smem(max_radius), smem1 = axis1_smem(max_radius);
    SCB_CUDA(cudaFuncSetAttribute(log_axis0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
    SCB_CUDA(cudaFuncSetAttribute(log_axis1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    const dim3 grid0((n_h + kACols - 1) / kACols, (n_w + kARows - 1) / kARows, n_sigma);
    log_axis0_kernel<<<grid0, dim3(32, 8), smem0, s>>>(d_image, n_w, n_h, d_radius, d_weights, weight_pitch, t2, t0);
    const dim3 grid1((n_h + kBCols - 1) / kBCols, (n_w + kBRows - 1) / kBRows, n_sigma);
    log_axis1_kernel<<<grid1, 256, smem1, s>>>(t2, t0, n_w, n_h, d_radius, d_weights, weight_pitch, d_sigma2, d_cube);
    SCB_CUDA_LAUNCH_CHECK("scb_log_scale_space");
    return 0;
}

extern "C" int scb_log_peaks(int n_w, int n_h, int n_sigma, const double *d_cube, double threshold, int32_t *d_peaks,
                             double *d_values, int64_t capacity, unsigned long long *d_count, void *stream) {
    SCB_REQUIRE(d_cube && d_peaks && d_values && d_count, SCB_E_NULL, "scb_log_peaks: NULL pointer");
    SCB_REQUIRE(n_w > 0 && n_h > 0 && n_sigma > 0 && capacity > 0, SCB_E_INVALID, "scb_log_peaks: %dx%dx%d, capacity %lld",
                n_w, n_h, n_sigma, (long long)capacity);
    cudaStream_t s = (cudaStream_t)stream;
    SCB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), s));
    const int64_t total = (int64_t)n_sigma * n_w * n_h;
    const unsigned int want = scb_grid_for(total, 256, 4);
    const unsigned int grid = want < (unsigned int)SCB_SM_COUNT * 16 ? want : (unsigned int)SCB_SM_COUNT * 16;
    log_peaks_kernel<<<grid, 256, 0, s>>>(d_cube, n_w, n_h, n_sigma, threshold, d_peaks, d_values, capacity, d_count);
    SCB_CUDA_LAUNCH_CHECK("scb_log_peaks");
    return 0;
}

extern "C" int scb_spot_fit(int n_w, int n_h, const double *d_image, int64_t n_blobs, const double *d_blobs,
                            int blob_stride, double roi_size, int max_iterations, double *d_spots, int32_t *d_status,
                            void *stream) {
    SCB_REQUIRE(d_image && d_blobs && d_spots && d_status, SCB_E_NULL, "scb_spot_fit: NULL pointer");
    SCB_REQUIRE(n_w > 0 && n_h > 0 && n_blobs > 0 && blob_stride >= 2, SCB_E_INVALID,
                "scb_spot_fit: image %dx%d, %lld blobs, stride %d", n_w, n_h, (long long)n_blobs, blob_stride);
    SCB_REQUIRE(roi_size >= 0 && (int)(2 * roi_size) + 2 <= kRoiMax, SCB_E_INVALID,
                "scb_spot_fit: roi_size=%g: the ROI must fit %d x %d pixels", roi_size, kRoiMax, kRoiMax);
    SCB_REQUIRE(max_iterations > 0, SCB_E_INVALID, "max_iterations=%d", max_iterations);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned int grid = (unsigned int)((n_blobs + kFitWarps - 1) / kFitWarps);
    spot_fit_kernel<<<grid, kFitWarps * 32, 0, s>>>(d_image, n_w, n_h, d_blobs, blob_stride, (int)n_blobs, roi_size,
                                                    max_iterations, d_spots, d_status);
    SCB_CUDA_LAUNCH_CHECK("scb_spot_fit");
    return 0;
}
