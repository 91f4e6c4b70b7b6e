// smooth_kernels.cu - map_img(interpolation='smooth'): PCHIP oversampling of the image on a
// regular grid followed by linear interpolation.
//
// Replaces, for every wavelength plane at once, BodyXY._do_smooth_interpolation
// (planetmapper/body_xy.py:1704-1790) and BodyXY._pchip_grid_interp2d (:1792-1853), i.e.
// scipy.interpolate.PchipInterpolator applied row by row and then column by column over the
// finite pixels, and scipy.interpolate.RegularGridInterpolator(method='linear',
// bounds_error=False, fill_value=nan) on the oversampled image.  The host computes the
// oversampled grids (get_xy_pchip, :1723-1741) from the NaN-min / max of the x / y maps.
//
// None of this is hot (the oversampled cube is ~25x the input, the gather is one load per
// corner): the kernels are straightforward one-thread-per-line / per-cell code whose job is
// to reproduce scipy's arithmetic (Fritsch-Butland derivatives with scipy's end rule, PPoly
// Horner evaluation, searchsorted interval choice) so that NaN masks are identical.
#include "pm_device.cuh"
#include "pm_kernels.h"

namespace pm {

// ---- NaN-aware min / max of a device array (np.nanmin / np.nanmax) -----------------------
__global__ void __launch_bounds__(1024) nan_minmax_kernel(const double *__restrict__ x, int64_t n,
                                                          double *__restrict__ out2) {
    double lo = INFINITY, hi = -INFINITY;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = x[i];
        if (v == v) {
            lo = fmin(lo, v);
            hi = fmax(hi, v);
        }
    }
    __shared__ double s_lo[1024], s_hi[1024];
    s_lo[threadIdx.x] = lo;
    s_hi[threadIdx.x] = hi;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s_lo[threadIdx.x] = fmin(s_lo[threadIdx.x], s_lo[threadIdx.x + o]);
            s_hi[threadIdx.x] = fmax(s_hi[threadIdx.x], s_hi[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // all-NaN input: NaN like numpy (with its RuntimeWarning)
        out2[0] = (s_lo[0] <= s_hi[0]) ? s_lo[0] : NAN;
        out2[1] = (s_lo[0] <= s_hi[0]) ? s_hi[0] : NAN;
    }
}

// ---- scipy PCHIP -----------------------------------------------------------------------------
__device__ __forceinline__ int sgn(double v) { return (v > 0.0) - (v < 0.0); }
// PchipInterpolator._edge_case
__device__ __forceinline__ double pchip_edge(double h0, double h1, double m0, double m1) {
    double d = ((2.0 * h0 + h1) * m0 - h0 * m1) / (h0 + h1);
    if (sgn(d) != sgn(m0)) {
        d = 0.0;
    } else if (sgn(m0) != sgn(m1) && fabs(d) > 3.0 * fabs(m0)) {
        d = 3.0 * m0;
    }
    return d;
}
// PchipInterpolator._find_derivatives, interior points
__device__ __forceinline__ double pchip_interior(double h0, double h1, double m0, double m1) {
    if (sgn(m0) != sgn(m1) || m0 == 0.0 || m1 == 0.0) return 0.0;
    const double w1 = 2.0 * h1 + h0, w2 = h1 + 2.0 * h0;
    const double whmean = (w1 / m0 + w2 / m1) / (w1 + w2);
    return 1.0 / whmean;
}
// np.linspace(start, stop, num)[k]
__device__ __forceinline__ double lin_at(double start, double stop, int num, int k) {
    if (num == 1) return start;
    if (k == num - 1) return stop;
    return (double)k * ((stop - start) / (double)(num - 1)) + start;
}

// PchipInterpolator(idx[mask], v[mask], extrapolate=False)(linspace(o_start, o_stop, n_out)) for
// one line: samples v[i * stride] at integer abscissae i in [lo, hi], mask = finite.  `d` is
// scratch with the same indexing as v.  Fewer than two finite samples -> all NaN (the
// reference skips the line, body_xy.py:1824-1825).
__device__ void pchip_line(const double *__restrict__ v, int64_t stride, int lo, int hi, double *__restrict__ d,
                           double o_start, double o_stop, int n_out, double *__restrict__ out, int64_t ostride) {
    // ---- pass A: derivatives at the finite samples
    int ia = -1, ib = -1, first = -1, last = -1, count = 0;
    double ya = 0.0, yb = 0.0;
    for (int i = lo; i <= hi; i++) {
        const double yc = v[i * stride];
        if (!(fabs(yc) < INFINITY)) continue;
        count++;
        if (first < 0) first = i;
        last = i;
        if (ib >= 0 && ia >= 0) {
            const double h0 = (double)(ib - ia), h1 = (double)(i - ib);
            const double m0 = (yb - ya) / h0, m1 = (yc - yb) / h1;
            if (ia == first) d[ia * stride] = pchip_edge(h0, h1, m0, m1);
            d[ib * stride] = pchip_interior(h0, h1, m0, m1);
        }
        ia = ib;
        ya = yb;
        ib = i;
        yb = yc;
    }
    if (count < 2) {
        for (int j = 0; j < n_out; j++) out[j * ostride] = NAN;
        return;
    }
    if (count == 2) {
        const double m = (yb - ya) / (double)(ib - ia);
        d[ia * stride] = m;
        d[ib * stride] = m;
    } else {
        // last point: edge rule with the last two intervals reversed; the sample before `ia`
        // was overwritten in the loop, so walk back to find it
        int ip = ia - 1;
        while (!(fabs(v[ip * stride]) < INFINITY)) ip--;
        const double yp = v[ip * stride];
        const double h0 = (double)(ib - ia), h1 = (double)(ia - ip);
        d[ib * stride] = pchip_edge(h0, h1, (yb - ya) / h0, (ya - yp) / h1);
    }
    // ---- pass B: evaluate on the output grid (both abscissa sets are increasing)
    int k0 = first, k1 = first + 1;
    while (!(fabs(v[k1 * stride]) < INFINITY)) k1++;
    for (int j = 0; j < n_out; j++) {
        const double x = lin_at(o_start, o_stop, n_out, j);
        double r = NAN;
        if (x >= (double)first && x <= (double)last) {
            while (x >= (double)k1 && k1 < last) {  // PPoly: x_i <= x < x_{i+1}; the last break uses the last piece
                k0 = k1;
                k1++;
                while (!(fabs(v[k1 * stride]) < INFINITY)) k1++;
            }
            const double y0 = v[k0 * stride], y1 = v[k1 * stride], d0 = d[k0 * stride], d1 = d[k1 * stride];
            const double h = (double)(k1 - k0);
            // CubicHermiteSpline -> PPoly coefficients, Horner in (x - x0)
            const double slope = (y1 - y0) / h;
            const double t = (d0 + d1 - 2.0 * slope) / h;
            const double c0 = t / h, c1 = (slope - d0) / h - t;
            const double s = x - (double)k0;
            r = ((c0 * s + c1) * s + d0) * s + y0;
        }
        out[j * ostride] = r;
    }
}

// rows: cube [n_planes][ny][nx] -> inter [n_planes][ny][n_xs]; rows outside [j0, j1] stay NaN
__global__ void __launch_bounds__(128) pchip_rows_kernel(const double *__restrict__ cube, int n_planes, int ny, int nx,
                                                         int i0, int i1, int j0, int j1, double x_start,
                                                         double x_stop, int n_xs, double *__restrict__ dscr,
                                                         double *__restrict__ inter) {
    const int64_t lines = (int64_t)n_planes * ny;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < lines; t += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(t % ny);
        double *out = inter + t * n_xs;
        if (r < j0 || r > j1) {
            for (int j = 0; j < n_xs; j++) out[j] = NAN;
            continue;
        }
        pchip_line(cube + t * nx, 1, i0, i1, dscr + t * nx, x_start, x_stop, n_xs, out, 1);
    }
}
// columns: inter [n_planes][ny][n_xs] -> fine [n_planes][n_ys][n_xs]
__global__ void __launch_bounds__(128) pchip_cols_kernel(const double *__restrict__ inter, int n_planes, int ny,
                                                         int n_xs, int j0, int j1, double y_start, double y_stop,
                                                         int n_ys, double *__restrict__ dscr,
                                                         double *__restrict__ fine) {
    const int64_t lines = (int64_t)n_planes * n_xs;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < lines; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = t / n_xs, c = t - l * n_xs;
        pchip_line(inter + l * (int64_t)ny * n_xs + c, n_xs, j0, j1, dscr + l * (int64_t)ny * n_xs + c, y_start, y_stop,
                   n_ys, fine + l * (int64_t)n_ys * n_xs + c, n_xs);
    }
}

// RegularGridInterpolator._find_indices for one axis of a linspace grid: searchsorted(grid, x) - 1
// clipped to [0, n - 2], and the normalised distance inside that interval
__device__ __forceinline__ void rgi_axis(double x, double start, double stop, int n, int &i_out, double &t_out) {
    // interval with grid[i] < x <= grid[i + 1] (searchsorted side='left'), i = 0 for x <= grid[0];
    // the arithmetic guess is at most one interval off
    const double step = (stop - start) / (double)(n - 1);
    int i = (int)floor((x - start) / step);
    i = max(0, min(i, n - 2));
    if (i > 0 && lin_at(start, stop, n, i) >= x) {
        i -= 1;
    } else if (i < n - 2 && lin_at(start, stop, n, i + 1) < x) {
        i += 1;
    }
    const double g0 = lin_at(start, stop, n, i), g1 = lin_at(start, stop, n, i + 1);
    i_out = i;
    t_out = (x - g0) / (g1 - g0);
}

// linear interpolation on the oversampled planes + the reference's NaN rules
__global__ void __launch_bounds__(256) gather_grid_linear_kernel(
    const double *__restrict__ fine, int n_planes, int n_ys, int n_xs, double x_start, double x_stop, double y_start,
    double y_stop, const double *__restrict__ cube, int ny, int nx, const double *__restrict__ xmap,
    const double *__restrict__ ymap, int64_t n_cells, uint32_t flags, double *__restrict__ out, int planes_per_group) {
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cells) return;
    const int l0 = blockIdx.y * planes_per_group, l1 = min(l0 + planes_per_group, n_planes);
    const double x = __ldg(xmap + cell), y = __ldg(ymap + cell);
    bool valid = !isnan(x);  // y is never NaN when x is not (body_xy.py:1778-1783)
    const bool propagate = (flags & PM_FLAG_PROPAGATE_NAN) != 0;
    int64_t nb[4] = {0, 0, 0, 0};
    if (valid && propagate) {
        // BodyXY._should_propagate_nan_to_map (body_xy.py:1855-1866)
        if (x < 0.0 || y < 0.0 || x > nx - 1 || y > ny - 1) valid = false;
        const int x0 = max((int)floor(x), 0), x1 = min((int)ceil(x), nx - 1);
        const int y0 = max((int)floor(y), 0), y1 = min((int)ceil(y), ny - 1);
        nb[0] = (int64_t)y0 * nx + x0;
        nb[1] = (int64_t)y0 * nx + x1;
        nb[2] = (int64_t)y1 * nx + x0;
        nb[3] = (int64_t)y1 * nx + x1;
    }
    // fill_value = nan outside the oversampled grid
    const bool inside = x >= x_start && x <= x_stop && y >= y_start && y <= y_stop;
    int ix = 0, iy = 0;
    double tx = 0.0, ty = 0.0;
    if (valid && inside && n_xs >= 2 && n_ys >= 2) {
        rgi_axis(x, x_start, x_stop, n_xs, ix, tx);
        rgi_axis(y, y_start, y_stop, n_ys, iy, ty);
    } else {
        valid = false;
    }
    const int64_t plane_px = (int64_t)ny * nx, fine_px = (int64_t)n_ys * n_xs;
    const int64_t o00 = (int64_t)iy * n_xs + ix;
    for (int l = l0; l < l1; l++) {
        double v = NAN;
        bool ok = valid;
        if (ok && propagate) {
            const double *im = cube + (int64_t)l * plane_px;
            ok = !(isnan(__ldg(im + nb[0])) || isnan(__ldg(im + nb[1])) || isnan(__ldg(im + nb[2])) ||
                   isnan(__ldg(im + nb[3])));
        }
        if (ok) {
            const double *f = fine + (int64_t)l * fine_px + o00;
            // RegularGridInterpolator._evaluate_linear: vertices (0,0), (0,1), (1,0), (1,1) of (y, x)
            v = 0.0;
            v += __ldg(f) * ((1.0 - ty) * (1.0 - tx));
            v += __ldg(f + 1) * ((1.0 - ty) * tx);
            v += __ldg(f + n_xs) * (ty * (1.0 - tx));
            v += __ldg(f + n_xs + 1) * (ty * tx);
        }
        __stcs(out + (int64_t)l * n_cells + cell, v);
    }
}

// ---------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------
cudaError_t launch_nan_minmax(const double *x, int64_t n, double *out2, cudaStream_t st) {
    nan_minmax_kernel<<<1, 1024, 0, st>>>(x, n, out2);
    count_launches(1);
    return cudaGetLastError();
}

static inline int64_t align256s(int64_t v) { return (v + 255) / 256 * 256; }
int64_t pchip_work_bytes(int n_planes, int ny, int nx, int n_xs) {
    return align256s((int64_t)n_planes * ny * nx * 8) + 2 * align256s((int64_t)n_planes * ny * n_xs * 8);
}
cudaError_t launch_pchip_resample(const double *cube, int n_planes, int ny, int nx, int i0, int i1, int j0, int j1,
                                  double x_start, double x_stop, int n_xs, double y_start, double y_stop, int n_ys,
                                  double *fine, void *work, int sm_count, cudaStream_t st) {
    if (n_planes == 0) return cudaSuccess;
    char *w = static_cast<char *>(work);
    double *d_rows = reinterpret_cast<double *>(w);
    w += align256s((int64_t)n_planes * ny * nx * 8);
    double *inter = reinterpret_cast<double *>(w);
    w += align256s((int64_t)n_planes * ny * n_xs * 8);
    double *d_cols = reinterpret_cast<double *>(w);
    const int64_t rows = (int64_t)n_planes * ny, cols = (int64_t)n_planes * n_xs;
    const int rb = (int)std::min<int64_t>((rows + 127) / 128, (int64_t)sm_count * 16);
    const int cb = (int)std::min<int64_t>((cols + 127) / 128, (int64_t)sm_count * 16);
    pchip_rows_kernel<<<rb, 128, 0, st>>>(cube, n_planes, ny, nx, i0, i1, j0, j1, x_start, x_stop, n_xs, d_rows, inter);
    pchip_cols_kernel<<<cb, 128, 0, st>>>(inter, n_planes, ny, n_xs, j0, j1, y_start, y_stop, n_ys, d_cols, fine);
    count_launches(2);
    return cudaGetLastError();
}
cudaError_t launch_gather_grid_linear(const double *fine, int n_planes, int n_ys, int n_xs, double x_start,
                                      double x_stop, double y_start, double y_stop, const double *cube, int ny, int nx,
                                      const double *xmap, const double *ymap, int64_t n_cells, uint32_t flags,
                                      double *out, cudaStream_t st) {
    if (n_cells == 0 || n_planes == 0) return cudaSuccess;
    int ppg = n_planes < 128 ? n_planes : 128;
    dim3 grid((unsigned)((n_cells + 255) / 256), (unsigned)((n_planes + ppg - 1) / ppg));
    gather_grid_linear_kernel<<<grid, 256, 0, st>>>(fine, n_planes, n_ys, n_xs, x_start, x_stop, y_start, y_stop, cube,
                                                    ny, nx, xmap, ymap, n_cells, flags, out, ppg);
    count_launches(1);
    return cudaGetLastError();
}

}  // namespace pm
