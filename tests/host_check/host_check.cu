// host_check.cu - TEST INFRASTRUCTURE ONLY.
//
// Instantiates the __host__ __device__ per-pixel / per-cell routines of
// planetmapper_b200/csrc/pm_device.cuh on the CPU so the non-GPU test suite can
// compare the very code the kernels run against the oracle (logic, conventions,
// conditioning).  It is built by tests/conftest.py into tests/host_check/_build/ and is
// never linked into or loaded by libpm_b200.so / planetmapper_b200 (no CPU fallback);
// the MUFU seeds are emulated (pm_math.cuh), so last-bit results differ from the GPU.
#include <cstdint>
#include <cstring>

#include "../../planetmapper_b200/csrc/pm_device.cuh"

namespace {
struct ArraySink {
    double *v;
    void put(int k, double val) { v[k] = val; }
};
}  // namespace

extern "C" {

// out: [26][n] (all plane ids, unrequested planes NaN)
int hc_backplanes_img(const PMFrame *frame, int nx, int ny, uint64_t mask, double *out) {
    pm::FrameD fs;
    pm::load_frame_host(fs, frame);
    const int64_t n = (int64_t)nx * ny;
    for (int64_t idx = 0; idx < n; idx++) {
        double v[PM_N_PLANES];
        for (int k = 0; k < PM_N_PLANES; k++) v[k] = NAN;
        ArraySink sink{v};
        if (mask & pm::kSkyMask)
            pm::image_pixel<true>(fs, (double)(idx % nx), (double)(idx / nx), mask, sink);
        else
            pm::image_pixel<false>(fs, (double)(idx % nx), (double)(idx / nx), mask, sink);
        for (int k = 0; k < PM_N_PLANES; k++) out[(int64_t)k * n + idx] = v[k];
    }
    return 0;
}

int hc_backplanes_map(const PMFrame *frame, const double *lon, const double *lat, int64_t n, uint64_t mask,
                      double *out) {
    pm::FrameD fs;
    pm::load_frame_host(fs, frame);
    for (int64_t idx = 0; idx < n; idx++) {
        double v[PM_N_PLANES];
        for (int k = 0; k < PM_N_PLANES; k++) v[k] = NAN;
        ArraySink sink{v};
        pm::map_cell(fs, lon[idx], lat[idx], mask, sink);
        for (int k = 0; k < PM_N_PLANES; k++) out[(int64_t)k * n + idx] = v[k];
    }
    return 0;
}

int hc_xy2lonlat(const PMFrame *frame, const double *x, const double *y, int64_t n, double *lon, double *lat,
                 int64_t *missed) {
    pm::FrameD fs;
    pm::load_frame_host(fs, frame);
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++) {
        lon[i] = lat[i] = NAN;
        if (fabs(x[i]) < INFINITY && fabs(y[i]) < INFINITY)
            if (!pm::xy2lonlat_point(fs, x[i], y[i], lon[i], lat[i])) m++;
    }
    *missed = m;
    return 0;
}

int hc_lonlat2xy(const PMFrame *frame, const double *lon, const double *lat, int64_t n, double alt, uint32_t flags,
                 double *x, double *y) {
    pm::FrameD fs;
    pm::load_frame_host(fs, frame);
    for (int64_t i = 0; i < n; i++) {
        x[i] = y[i] = NAN;
        if (fabs(lon[i]) < INFINITY && fabs(lat[i]) < INFINITY)
            pm::lonlat2xy_point(fs, lon[i], lat[i], alt, (flags & PM_FLAG_NOT_VISIBLE_NAN) != 0,
                                (flags & PM_FLAG_PLANETOCENTRIC) != 0, x[i], y[i]);
    }
    return 0;
}

// pm_transform's per-point code; aux13 as in include/pm_b200.h (never NULL here)
int hc_transform(const PMFrame *frame, int src, int dst, const double *a, const double *b, int64_t n, double alt,
                 uint32_t flags, const double *aux13, double *oa, double *ob, int64_t *missed) {
    pm::FrameD fs;
    pm::load_frame_host(fs, frame);
    pm::TransformAux aux;
    for (int i = 0; i < 9; i++) aux.Mc[i] = aux13[i];
    for (int i = 0; i < 4; i++) aux.km2ang[i] = aux13[9 + i];
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++)
        if (!pm::point_transform(fs, aux, src, dst, a[i], b[i], alt, flags, oa[i], ob[i])) m++;
    *missed = m;
    return 0;
}

// kind as pm_math_probe (include/pm_b200.h)
int hc_math(int kind, const double *a, const double *b, int64_t n, double *out) {
    for (int64_t i = 0; i < n; i++) {
        const double x = a[i], y = b ? b[i] : 0.0;
        double s, c, r = NAN;
        switch (kind) {
            case 0: r = pm::fast_rcp(x); break;
            case 1: r = pm::fast_rsqrt(x); break;
            case 2: r = pm::fast_sqrt(x); break;
            case 3: pm::sincos_small(x, s, c); r = s; break;
            case 4: pm::sincos_small(x, s, c); r = c; break;
            case 5: r = pm::fast_atan2(x, y); break;
            case 10: r = pm::fast_atan2_ypos(fabs(x), y); break;
            case 11: r = pm::fast_atan2_xpos(x, fabs(y)); break;
            case 6: r = pm::fast_acos(x); break;
            case 7: r = pm::fast_div(x, y); break;
            case 8: pm::sincos_full(x, s, c); r = s; break;
            case 9: pm::sincos_full(x, s, c); r = c; break;
            default: break;
        }
        out[i] = r;
    }
    return 0;
}
}
