// proj_kernels.cu - inverse map projections on the GPU.
//
// Replaces pyproj.Transformer.transform(xx, yy, direction='INVERSE')
// (planetmapper/body_xy.py:3126) for the projection strings the reference builds at
// body_xy.py:2899-2969: ellipsoidal orthographic (+proj=ortho +a +b +to_meter=a
// +y_0=...), spherical azimuthal equidistant (+proj=aeqd, +to_meter = a pi) and
// spherical Lambert azimuthal equal area (+proj=laea, +to_meter = 2a).  The maths
// follows PROJ 9's published inverse formulas (EPSG method 9840 with a Newton solve
// for the oblique ellipsoidal case).  One thread per grid node, coalesced I/O.
#include "pm_device.cuh"
#include "pm_kernels.h"

namespace pm {

struct ProjParams {
    double a, b, lon0_deg, lat0_deg, lon_sign;
};

__device__ __forceinline__ bool ortho_sph_inverse(double x, double y, double phi0, double sinph0,
                                                  double cosph0, double &phi, double &lam) {
    double rh = hypot(x, y), sinc = rh;
    if (sinc > 1.0) {
        if (sinc - 1.0 > 1e-10) return false;
        sinc = 1.0;
    }
    double cosc = sqrt(1.0 - sinc * sinc);
    if (fabs(rh) <= 1e-10) {
        phi = phi0;
        lam = 0.0;
        return true;
    }
    double p;
    bool obl_or_eq = true;
    if (fabs(fabs(phi0) - kHalfPi) < 1e-10) {
        obl_or_eq = false;
        if (phi0 > 0.0) {
            y = -y;
            p = acos(sinc);
        } else {
            p = -acos(sinc);
        }
    } else {
        if (fabs(phi0) < 1e-10) {
            p = y * sinc / rh;
            x *= sinc;
            y = cosc * rh;
        } else {
            p = cosc * sinph0 + y * sinc * cosph0 / rh;
            y = (cosc - sinph0 * p) * rh;
            x *= sinc * cosph0;
        }
        p = (fabs(p) >= 1.0) ? (p < 0.0 ? -kHalfPi : kHalfPi) : asin(p);
    }
    phi = p;
    if (y == 0.0 && obl_or_eq)
        lam = (x == 0.0) ? 0.0 : (x < 0.0 ? -kHalfPi : kHalfPi);
    else
        lam = atan2(x, y);
    return true;
}

__device__ __forceinline__ bool proj_inverse_one(int kind, const ProjParams &pp, double xx, double yy,
                                                 double &lon_deg, double &lat_deg) {
    if (!(isfinite(xx) && isfinite(yy))) return false;
    const double phi0 = pp.lat0_deg * kRpd;
    double sinph0, cosph0;
    sincos(phi0, &sinph0, &cosph0);
    double x = (pp.lon_sign < 0.0) ? -xx : xx;  // +axis=wnu
    double y = yy;
    double lam = 0.0, phi = 0.0;
    const bool polar = fabs(fabs(phi0) - kHalfPi) < 1e-10;
    const bool equit = fabs(phi0) < 1e-10;
    if (kind == PM_PROJ_ORTHOGRAPHIC) {
        y = yy - (pp.b / pp.a - 1.0) * sin((pp.lat0_deg * 2.0) * kRpd);  // +y_0, +to_meter=a
        const double es = 1.0 - (pp.b * pp.b) / (pp.a * pp.a);
        const double one_es = 1.0 - es;
        if (es == 0.0) {
            if (!ortho_sph_inverse(x, y, phi0, sinph0, cosph0, phi, lam)) return false;
        } else if (polar) {
            const double sgn = phi0 > 0.0 ? 1.0 : -1.0;
            double rh2 = x * x + y * y;
            if (rh2 >= 1.0 - 1e-15) {
                if (rh2 - 1.0 > 1e-10) return false;
                phi = 0.0;
            } else {
                phi = acos(sqrt(rh2 * one_es / (1.0 - es * rh2))) * sgn;
            }
            lam = atan2(x, y * -sgn);
        } else if (equit) {
            double ys = y * (pp.a / pp.b);
            if (x * x + ys * ys > 1.0 + 1e-11) return false;
            double q = one_es / y;
            double sinphi2 = (y == 0.0) ? 0.0 : 1.0 / (q * q + es);
            if (sinphi2 > 1.0 - 1e-11) {
                phi = kHalfPi * (y > 0.0 ? 1.0 : -1.0);
                lam = 0.0;
            } else {
                phi = asin(sqrt(sinphi2)) * (y > 0.0 ? 1.0 : -1.0);
                double sinlam = x * sqrt((1.0 - es * sinphi2) / (1.0 - sinphi2));
                if (fabs(sinlam) - 1.0 > -1e-15)
                    lam = kHalfPi * (x > 0.0 ? 1.0 : -1.0);
                else
                    lam = asin(sinlam);
            }
        } else {
            const double nu0 = 1.0 / sqrt(1.0 - es * sinph0 * sinph0);
            const double y_shift = es * nu0 * sinph0 * cosph0;
            const double y_scale = 1.0 / sqrt(1.0 - es * cosph0 * cosph0);
            double yr = (y - y_shift) / y_scale;
            if (x * x + yr * yr > 1.0 + 1e-11) return false;
            if (!ortho_sph_inverse(x, yr, phi0, sinph0, cosph0, phi, lam)) return false;
            bool ok = false;
            // Newton on (phi, lam): own FP64 primitives (pm_math.cuh; 1-2 ulp) instead of libm's sincos /
            // division / sqrt slow paths - this loop is the whole cost of an oblique orthographic map
            for (int it = 0; it < 20; it++) {
                double sp, cp, sl, cl;
                sincos_full(phi, sp, cp);
                sincos_full(lam, sl, cl);
                double w = 1.0 - es * sp * sp;
                double nu = fast_rsqrt(w);
                double fx = nu * cp * sl;
                double fy = nu * (sp * cosph0 - cp * sinph0 * cl) + es * (nu0 * sinph0 - nu * sp) * cosph0;
                double rho = one_es * nu * (nu * nu);
                double J11 = -rho * sp * sl, J12 = nu * cp * cl;
                double J21 = rho * (cp * cosph0 + sp * sinph0 * cl), J22 = nu * sinph0 * cp * sl;
                double iD = fast_rcp(J11 * J22 - J12 * J21);
                double dx = x - fx, dy = y - fy;
                double dphi = (J22 * dx - J12 * dy) * iD;
                double dlam = (-J21 * dx + J11 * dy) * iD;
                phi += dphi;
                if (phi > kHalfPi)
                    phi = kHalfPi - (phi - kHalfPi);
                else if (phi < -kHalfPi)
                    phi = -kHalfPi + (-kHalfPi - phi);
                lam += dlam;
                if (fabs(dphi) < 1e-12 && fabs(dlam) < 1e-12) {
                    ok = true;
                    break;
                }
            }
            if (!ok) return false;
        }
    } else if (kind == PM_PROJ_AZIMUTHAL) {
        x *= kPi;
        y *= kPi;
        double c_rh = hypot(x, y);
        bool at_origin = false;
        if (c_rh > kPi) {
            if (c_rh - 1e-10 > kPi) return false;
            c_rh = kPi;
        } else if (c_rh < 1e-10) {
            at_origin = true;
        }
        if (at_origin) {
            phi = phi0;
            lam = 0.0;
        } else if (polar) {
            if (phi0 > 0.0) {
                phi = kHalfPi - c_rh;
                lam = atan2(x, -y);
            } else {
                phi = c_rh - kHalfPi;
                lam = atan2(x, y);
            }
        } else {
            double sinc, cosc;
            sincos(c_rh, &sinc, &cosc);
            double arg;
            if (equit) {
                arg = y * sinc / c_rh;
                x *= sinc;
                y = cosc * c_rh;
            } else {
                arg = cosc * sinph0 + y * sinc * cosph0 / c_rh;
                arg = fmax(-1.0, fmin(1.0, arg));
                y = (cosc - sinph0 * arg) * c_rh;
                x *= sinc * cosph0;
            }
            arg = fmax(-1.0, fmin(1.0, arg));
            phi = asin(arg);
            lam = (y == 0.0) ? 0.0 : atan2(x, y);
        }
    } else if (kind == PM_PROJ_AZIMUTHAL_EQUAL_AREA) {
        x *= 2.0;
        y *= 2.0;
        double rh = hypot(x, y);
        double half = rh * 0.5;
        if (half > 1.0) return false;
        double z = 2.0 * asin(half);
        if (polar) {
            if (phi0 > 0.0) {
                y = -y;
                phi = kHalfPi - z;
            } else {
                phi = z - kHalfPi;
            }
            lam = atan2(x, y);
        } else {
            double sinz, cosz;
            sincos(z, &sinz, &cosz);
            if (equit) {
                phi = (fabs(rh) <= 1e-10) ? 0.0 : asin(y * sinz / rh);
                x *= sinz;
                y = cosz * rh;
            } else {
                phi = (fabs(rh) <= 1e-10) ? phi0 : asin(cosz * sinph0 + y * sinz * cosph0 / rh);
                x *= sinz * cosph0;
                y = (cosz - sin(phi) * sinph0) * rh;
            }
            lam = (y == 0.0) ? 0.0 : atan2(x, y);
        }
    } else {
        return false;
    }
    double lon_rad = lam + pp.lon0_deg * kRpd;
    if (fabs(lon_rad) > kPi + 1e-12) {
        lon_rad += kPi;
        lon_rad -= kTwoPi * floor(lon_rad / kTwoPi);
        lon_rad -= kPi;
    }
    lon_deg = lon_rad * kDpr;
    lat_deg = phi * kDpr;
    return true;
}

// Forward direction (pyproj.Transformer.transform(lon, lat), the default direction of the transformer
// generate_map_coordinates hands back, body_xy.py:3141-3153): planetographic lon / lat (degrees) -> the
// projection's own units.  Same parameter conventions as the inverse; points the projection cannot show
// (far side for ortho, antipode for the azimuthal ones) give NaN (pyproj: inf, mapped to NaN by the reference).
__device__ __forceinline__ bool proj_forward_one(int kind, const ProjParams &pp, double lon_deg, double lat_deg,
                                                 double &xx, double &yy) {
    if (!(isfinite(lon_deg) && isfinite(lat_deg)) || fabs(lat_deg) > 90.0 + 1e-12) return false;
    const double phi0 = pp.lat0_deg * kRpd, phi = fmax(-kHalfPi, fmin(kHalfPi, lat_deg * kRpd));
    double lam = (lon_deg - pp.lon0_deg) * kRpd;
    lam -= kTwoPi * floor((lam + kPi) / kTwoPi);   // (-pi, pi]
    double sinph0, cosph0, sinphi, cosphi, sinlam, coslam;
    sincos(phi0, &sinph0, &cosph0);
    sincos(phi, &sinphi, &cosphi);
    sincos(lam, &sinlam, &coslam);
    const double cosc = sinph0 * sinphi + cosph0 * cosphi * coslam;   // cosine of the angular distance from the centre
    double x, y;
    if (kind == PM_PROJ_ORTHOGRAPHIC) {
        if (cosc < -1e-10) return false;   // behind the projection plane
        const double es = 1.0 - (pp.b * pp.b) / (pp.a * pp.a);
        const double nu = 1.0 / sqrt(1.0 - es * sinphi * sinphi);
        const double nu0 = 1.0 / sqrt(1.0 - es * sinph0 * sinph0);
        x = nu * cosphi * sinlam;
        y = nu * (sinphi * cosph0 - cosphi * sinph0 * coslam) + es * (nu0 * sinph0 - nu * sinphi) * cosph0;
        y += (pp.b / pp.a - 1.0) * sin((pp.lat0_deg * 2.0) * kRpd);   // +y_0 with +to_meter = a
    } else if (kind == PM_PROJ_AZIMUTHAL) {
        if (cosc <= -1.0 + 1e-14) return false;   // the antipode has no image
        const double c = acos(fmax(-1.0, fmin(1.0, cosc)));
        const double sinc = sin(c);
        const double k = c < 1e-10 ? 1.0 : c / sinc;
        x = k * cosphi * sinlam / kPi;            // +to_meter = a pi
        y = k * (cosph0 * sinphi - sinph0 * cosphi * coslam) / kPi;
    } else if (kind == PM_PROJ_AZIMUTHAL_EQUAL_AREA) {
        const double d = 1.0 + cosc;
        if (d <= 1e-10) return false;
        const double k = sqrt(2.0 / d);
        x = k * cosphi * sinlam * 0.5;            // +to_meter = 2 a
        y = k * (cosph0 * sinphi - sinph0 * cosphi * coslam) * 0.5;
    } else {
        return false;
    }
    xx = (pp.lon_sign < 0.0) ? -x : x;   // +axis=wnu
    yy = y;
    return true;
}

__global__ void __launch_bounds__(kBlock) proj_forward_kernel(int kind, ProjParams pp, const double *__restrict__ lon,
                                                              const double *__restrict__ lat, int64_t n,
                                                              double *__restrict__ xx, double *__restrict__ yy) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (int64_t)gridDim.x * blockDim.x) {
        double x = NAN, y = NAN;
        if (!proj_forward_one(kind, pp, lon[idx], lat[idx], x, y)) x = y = NAN;
        xx[idx] = x;
        yy[idx] = y;
    }
}

__global__ void __launch_bounds__(kBlock) proj_inverse_kernel(int kind, ProjParams pp,
                                                              const double *__restrict__ xx,
                                                              const double *__restrict__ yy, int64_t n,
                                                              double *__restrict__ lon,
                                                              double *__restrict__ lat) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (int64_t)gridDim.x * blockDim.x) {
        double lo = NAN, la = NAN;
        if (!proj_inverse_one(kind, pp, xx[idx], yy[idx], lo, la)) {
            lo = NAN;
            la = NAN;
        }
        lon[idx] = lo;
        lat[idx] = la;
    }
}

cudaError_t launch_proj_inverse(int kind, const double *p5, const double *xx, const double *yy, int64_t n,
                                double *lon, double *lat, int sm_count, cudaStream_t st) {
    ProjParams pp{p5[0], p5[1], p5[2], p5[3], p5[4]};
    int64_t blocks = (n + kBlock - 1) / kBlock;
    int64_t cap = (int64_t)sm_count * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    proj_inverse_kernel<<<(int)blocks, kBlock, 0, st>>>(kind, pp, xx, yy, n, lon, lat);
    count_launches(1);
    return cudaGetLastError();
}

cudaError_t launch_proj_forward(int kind, const double *p5, const double *lon, const double *lat, int64_t n,
                                double *xx, double *yy, int sm_count, cudaStream_t st) {
    ProjParams pp{p5[0], p5[1], p5[2], p5[3], p5[4]};
    int64_t blocks = (n + kBlock - 1) / kBlock;
    int64_t cap = (int64_t)sm_count * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    proj_forward_kernel<<<(int)blocks, kBlock, 0, st>>>(kind, pp, lon, lat, n, xx, yy);
    count_launches(1);
    return cudaGetLastError();
}

}  // namespace pm
