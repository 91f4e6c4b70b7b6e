// pm_device.cuh - FP64 device math for the PlanetMapper hot path on sm_100a.
//
// Everything here is per-pixel / per-cell arithmetic that the reference performs
// with one ctypes CSPICE call per pixel per stage ("third-party
// call sites").  Each routine cites the reference call site it replaces
// (file:line under /root/reference/planetmapper).  The code is written for the GPU:
// frame constants and their derived values live in shared memory (FrameS), vectors
// stay in registers, rotations share one sincos, reciprocals of the radii are
// precomputed once per block, and the biaxial (a == b) geodetic case takes a closed
// form instead of the iterative nearest-point solve.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/pm_b200.h"

namespace pm {

constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double kTwoPi = 2.0 * kPi;
constexpr double kHalfPi = 0.5 * kPi;
constexpr double kDpr = 180.0 / kPi;
constexpr double kRpd = kPi / 180.0;
constexpr int kMaxItr = 10;

struct V3 {
    double x, y, z;
};

__device__ __forceinline__ V3 mk(double x, double y, double z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ double norm(V3 a) { return sqrt(dot(a, a)); }
__device__ __forceinline__ V3 ld3(const double *p) { return mk(p[0], p[1], p[2]); }
__device__ __forceinline__ bool finite3(V3 a) { return isfinite(a.x) && isfinite(a.y) && isfinite(a.z); }
__device__ __forceinline__ V3 mxv(const double *m, V3 v) {
    return mk(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z,
              m[6] * v.x + m[7] * v.y + m[8] * v.z);
}
__device__ __forceinline__ V3 mtxv(const double *m, V3 v) {
    return mk(m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z,
              m[2] * v.x + m[5] * v.y + m[8] * v.z);
}
// Python float % (sign of the divisor), divisor > 0
__device__ __forceinline__ double pymod_pos(double a, double m) {
    double r = fmod(a, m);
    if (r < 0.0) r += m;
    return r;
}

// Frame constants + per-block derived values (shared memory)
struct FrameS {
    PMFrame f;
    double inv_r[3];   // 1 / radii
    double nw[3];      // surfnm weights (min_radius / radius)^2
    double k[3];       // unit rotation axis (body frame)
    double wn;         // |omega|
    double rp;         // re (1 - f)
    double e2, ep2;    // first / second eccentricity squared of the recpgr spheroid
    double inv_omf2;   // 1 / (1 - f)^2
    int biaxial;       // intercept ellipsoid == recpgr spheroid -> closed-form geodetic
};

// Cooperative load of one PMFrame into shared memory + derived constants.
__device__ __forceinline__ void load_frame(FrameS &s, const PMFrame *__restrict__ g) {
    const double *src = reinterpret_cast<const double *>(g);
    double *dst = reinterpret_cast<double *>(&s.f);
    for (int i = threadIdx.x; i < PM_FRAME_NDOUBLES; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        const PMFrame &f = s.f;
        double a = f.radii[0], b = f.radii[1], c = f.radii[2];
        s.inv_r[0] = 1.0 / a;
        s.inv_r[1] = 1.0 / b;
        s.inv_r[2] = 1.0 / c;
        double m = fmin(a, fmin(b, c));
        s.nw[0] = (m / a) * (m / a);
        s.nw[1] = (m / b) * (m / b);
        s.nw[2] = (m / c) * (m / c);
        double wn = sqrt(f.omega[0] * f.omega[0] + f.omega[1] * f.omega[1] + f.omega[2] * f.omega[2]);
        s.wn = wn;
        double iw = wn > 0.0 ? 1.0 / wn : 0.0;
        s.k[0] = f.omega[0] * iw;
        s.k[1] = f.omega[1] * iw;
        s.k[2] = f.omega[2] * iw;
        double rp = f.re - f.f * f.re;
        s.rp = rp;
        s.e2 = 1.0 - (rp * rp) / (f.re * f.re);
        s.ep2 = (f.re * f.re) / (rp * rp) - 1.0;
        s.inv_omf2 = 1.0 / ((1.0 - f.f) * (1.0 - f.f));
        s.biaxial = (a == b) && (f.re == a) && (fabs(rp - c) <= 4e-16 * c);
    }
    __syncthreads();
}

// sin(theta), 1 - cos(theta) for the frame rotation over dt, accurate for tiny angles
struct Rot {
    double s, omc;
};
__device__ __forceinline__ Rot make_rot(const FrameS &fs, double dt) {
    double sh, ch;
    sincos(0.5 * fs.wn * dt, &sh, &ch);
    return Rot{2.0 * sh * ch, 2.0 * sh * sh};
}
// w (already R0 v, i.e. body frame at t_ref) -> body frame at t_ref + dt:
// exp(-theta [k]x) w = w - sin(theta) (k x w) - (1 - cos(theta)) (w - k (k.w))
__device__ __forceinline__ V3 spin_fwd(const FrameS &fs, Rot r, V3 w) {
    V3 k = ld3(fs.k);
    V3 kxw = cross(k, w);
    double kw = dot(k, w);
    return mk(w.x - r.s * kxw.x - r.omc * (w.x - k.x * kw), w.y - r.s * kxw.y - r.omc * (w.y - k.y * kw),
              w.z - r.s * kxw.z - r.omc * (w.z - k.z * kw));
}
// body frame at t_ref + dt -> body frame at t_ref (inverse spin)
__device__ __forceinline__ V3 spin_bwd(const FrameS &fs, Rot r, V3 u) {
    V3 k = ld3(fs.k);
    V3 kxu = cross(k, u);
    double ku = dot(k, u);
    return mk(u.x + r.s * kxu.x - r.omc * (u.x - k.x * ku), u.y + r.s * kxu.y - r.omc * (u.y - k.y * ku),
              u.z + r.s * kxu.z - r.omc * (u.z - k.z * ku));
}
// pxform('J2000', body, t_ref + dt) v   (inside sincpt / illumf / spkcpt, body.py:998)
__device__ __forceinline__ V3 to_body(const FrameS &fs, Rot r, V3 v) {
    return spin_fwd(fs, r, mxv(fs.f.R0, v));
}
// pxform(body, 'J2000', t_ref + dt) u   (body.py:940)
__device__ __forceinline__ V3 from_body(const FrameS &fs, Rot r, V3 u) {
    return mtxv(fs.f.R0, spin_bwd(fs, r, u));
}
// target centre relative to the observer at t_ref + dt
__device__ __forceinline__ V3 target_pos(const PMFrame &f, double dt) {
    double h = 0.5 * dt * dt;
    return mk(f.P0[0] + f.VT[0] * dt + f.AT[0] * h, f.P0[1] + f.VT[1] * dt + f.AT[1] * h,
              f.P0[2] + f.VT[2] * dt + f.AT[2] * h);
}

// spice.recrad angles (base.py:902): RA in [0, 2pi), Dec
__device__ __forceinline__ void recrad_angles(V3 v, double &ra, double &dec) {
    double big = fmax(fabs(v.x), fmax(fabs(v.y), fabs(v.z)));
    if (!(big > 0.0)) {
        ra = (big == 0.0) ? 0.0 : NAN;
        dec = ra;
        return;
    }
    double ib = 1.0 / big;
    double x = v.x * ib, y = v.y * ib, z = v.z * ib;
    dec = atan2(z, sqrt(x * x + y * y));
    double lon = (x == 0.0 && y == 0.0) ? 0.0 : atan2(y, x);
    if (lon < 0.0) lon += kTwoPi;
    ra = lon;
}

// spice.radrec(1, ra, dec) (body.py:967, :1369)
__device__ __forceinline__ V3 radrec1(double ra, double dec) {
    double sr, cr, sd, cd;
    sincos(ra, &sr, &cr);
    sincos(dec, &sd, &cd);
    return mk(cr * cd, sr * cd, sd);
}

// spice.vsep for a unit vector pair
__device__ __forceinline__ double vsep_unit(V3 u, V3 v) {
    double d = dot(u, v);
    if (d > 0.0) return 2.0 * asin(0.5 * norm(u - v));
    if (d < 0.0) return kPi - 2.0 * asin(0.5 * norm(u + v));
    return kHalfPi;
}
__device__ __forceinline__ V3 unit(V3 a) {
    double n = norm(a);
    double in = 1.0 / n;
    return mk(a.x * in, a.y * in, a.z * in);
}

// spice.surfpt (inside sincpt, body.py:1010): nearest ray / ellipsoid intersection,
// perpendicular-projection form.  o, u in the body frame.
__device__ __forceinline__ bool surfpt(const FrameS &fs, V3 o, V3 u, V3 &p) {
    V3 x = mk(u.x * fs.inv_r[0], u.y * fs.inv_r[1], u.z * fs.inv_r[2]);
    V3 y = mk(o.x * fs.inv_r[0], o.y * fs.inv_r[1], o.z * fs.inv_r[2]);
    double xn2 = dot(x, x);
    if (!(xn2 > 0.0)) return false;
    double ixn = 1.0 / sqrt(xn2);
    x = ixn * x;
    double yx = dot(y, x);
    V3 pp = mk(y.x - yx * x.x, y.y - yx * x.y, y.z - yx * x.z);
    double pm2 = dot(pp, pp), ym2 = dot(y, y);
    if (!isfinite(pm2)) return false;
    double pmag = sqrt(pm2);
    V3 q;
    if (ym2 > 1.0) {
        if (pmag > 1.0) return false;
        if (yx > 0.0) return false;
        double sc = sqrt(fmax(0.0, 1.0 - pmag * pmag));
        q = mk(pp.x - sc * x.x, pp.y - sc * x.y, pp.z - sc * x.z);
    } else if (ym2 == 1.0) {
        q = y;
    } else {
        double sc = sqrt(fmax(0.0, 1.0 - pmag * pmag));
        q = mk(pp.x + sc * x.x, pp.y + sc * x.y, pp.z + sc * x.z);
    }
    p = mk(q.x * fs.f.radii[0], q.y * fs.f.radii[1], q.z * fs.f.radii[2]);
    return true;
}

// spice.sincpt(..., 'CN', ..., d) (body.py:1008-1020): intercept with the light time
// iterated on the intercept point.  d: J2000 ray direction.
__device__ __forceinline__ bool sincpt(const FrameS &fs, V3 d, V3 &p, double &lt) {
    const PMFrame &f = fs.f;
    V3 d0 = mxv(f.R0, d);  // ray in the body frame at t_ref (spin applied per pass)
    double t = f.et - f.lt0;
    lt = f.lt0;
    for (int i = 0; i < kMaxItr; i++) {
        double dt = t - f.t_ref;
        Rot r = make_rot(fs, dt);
        V3 o = spin_fwd(fs, r, mxv(f.R0, -target_pos(f, dt)));
        V3 u = spin_fwd(fs, r, d0);
        if (!surfpt(fs, o, u, p)) return false;
        double lt_new = norm(p - o) / f.clight;
        double t_new = f.et - lt_new;
        double ltdiff = fabs(t_new - t);
        t = t_new;
        lt = lt_new;
        if (!(ltdiff > 1.0e-17 * fabs(t))) break;
    }
    return true;
}

// Geodetic lon (east-positive), lat, alt w.r.t. the spheroid (re, re, re(1-f)):
// spice.recgeo inside spice.recpgr (body.py:1030, :2592).
__device__ __forceinline__ void recgeo(const FrameS &fs, V3 p, bool on_spheroid, double &lon,
                                       double &lat, double &alt) {
    const PMFrame &f = fs.f;
    double rho = sqrt(p.x * p.x + p.y * p.y);
    lon = (p.x == 0.0 && p.y == 0.0) ? 0.0 : atan2(p.y, p.x);
    if (on_spheroid) {
        // point lies on the spheroid itself: the normal there is the geodetic normal
        lat = atan2(p.z * fs.inv_omf2, rho);
        alt = 0.0;
        return;
    }
    if (f.f == 0.0) {
        lat = atan2(p.z, rho);
        alt = norm(p) - f.re;
        return;
    }
    // Bowring's iteration run to convergence
    double beta = atan2(f.re * p.z, fs.rp * rho);
    double phi = 0.0;
    for (int i = 0; i < 12; i++) {
        double sb, cb;
        sincos(beta, &sb, &cb);
        double nphi = atan2(p.z + fs.ep2 * fs.rp * sb * sb * sb, rho - fs.e2 * f.re * cb * cb * cb);
        double sp, cp;
        sincos(nphi, &sp, &cp);
        bool done = (i > 0) && fabs(nphi - phi) <= 4.0e-16 * fmax(1.0, fabs(nphi));
        phi = nphi;
        if (done) break;
        beta = atan2((1.0 - f.f) * sp, cp);
    }
    double sp, cp;
    sincos(phi, &sp, &cp);
    lat = phi;
    alt = rho * cp + p.z * sp - f.re * sqrt(1.0 - fs.e2 * sp * sp);
}

// spice.recpgr: planetographic lon in [0, 2pi)
__device__ __forceinline__ void recpgr(const FrameS &fs, V3 p, bool on_spheroid, double &lon,
                                       double &lat, double &alt) {
    double l;
    recgeo(fs, p, on_spheroid, l, lat, alt);
    l = fs.f.lon_sign * l;
    if (l < 0.0) l += kTwoPi;
    lon = l;
}

// spice.pgrrec(lon, lat, alt = 0) (body.py:903), radians in
__device__ __forceinline__ V3 pgrrec0(const FrameS &fs, double lon, double lat) {
    const PMFrame &f = fs.f;
    double slat, clat, slon, clon;
    sincos(lat, &slat, &clat);
    sincos(f.lon_sign * lon, &slon, &clon);
    double big = fmax(fabs(f.re * clat), fabs(fs.rp * slat));
    double x = f.re * clat / big, y = fs.rp * slat / big;
    double scale = 1.0 / sqrt(x * x + y * y);
    return mk(scale * clon * x * f.re, scale * slon * x * f.re, scale * y * fs.rp);
}

// spice.reclat angles (body.py:2912)
__device__ __forceinline__ void reclat_angles(V3 v, double &lon, double &lat) {
    double big = fmax(fabs(v.x), fmax(fabs(v.y), fabs(v.z)));
    if (big > 0.0) {
        double ib = 1.0 / big;
        double x = v.x * ib, y = v.y * ib, z = v.z * ib;
        lat = atan2(z, sqrt(x * x + y * y));
        lon = (x == 0.0 && y == 0.0) ? 0.0 : atan2(y, x);
    } else {
        lon = 0.0;
        lat = 0.0;
    }
}

struct PointState {
    double lt;      // light time point -> observer
    double rv;      // radial velocity (spkcpt velocity . unit position)
    double phase, incdnc, emissn;
};

// spice.spkcpt (body.py:2830-2845) and spice.illumf (body.py:1915-1935) for the
// body-fixed point p; lt_start seeds the light time iteration.
template <bool kState, bool kIllum>
__device__ __forceinline__ void point_state(const FrameS &fs, V3 p, double lt_start, PointState &s) {
    const PMFrame &f = fs.f;
    double lt = lt_start, dt = 0.0;
    Rot r;
    V3 q, X;
    for (int i = 0; i < kMaxItr; i++) {
        dt = (f.et - lt) - f.t_ref;
        r = make_rot(fs, dt);
        q = from_body(fs, r, p);
        X = target_pos(f, dt) + q;
        double lt_new = norm(X) / f.clight;
        double diff = fabs(lt_new - lt);
        lt = lt_new;
        if (!(diff > 1.0e-17 * fabs(f.et))) break;
    }
    {
        double dt2 = (f.et - lt) - f.t_ref;
        if (dt2 != dt) {
            dt = dt2;
            r = make_rot(fs, dt);
            q = from_body(fs, r, p);
            X = target_pos(f, dt) + q;
        }
    }
    s.lt = lt;
    if (kState) {
        V3 vrot = from_body(fs, r, cross(ld3(f.omega), p));
        V3 VX = mk(f.VT[0] + f.AT[0] * dt + vrot.x, f.VT[1] + f.AT[1] * dt + vrot.y,
                   f.VT[2] + f.AT[2] * dt + vrot.z);
        double rn = norm(X);
        V3 ph = mk(X.x / rn, X.y / rn, X.z / rn);
        V3 VO = ld3(f.VO);
        double dlt = dot(ph, VX - VO) / (f.clight + dot(ph, VX));
        V3 vel = (1.0 - dlt) * VX - VO;
        s.rv = dot(vel, ph);
    }
    if (kIllum) {
        V3 e_b = to_body(fs, r, -X);  // point -> observer, body frame at the point epoch
        double h = 0.5 * dt * dt;
        V3 Tc = mk(f.VT[0] * dt + f.AT[0] * h + q.x, f.VT[1] * dt + f.AT[1] * h + q.y,
                   f.VT[2] * dt + f.AT[2] * h + q.z);
        double lts = f.lts0;
        V3 sv;
        for (int i = 0; i < kMaxItr; i++) {
            double ds = dt - (lts - f.lts0);
            sv = mk(f.S0[0] + f.VS[0] * ds - Tc.x, f.S0[1] + f.VS[1] * ds - Tc.y,
                    f.S0[2] + f.VS[2] * ds - Tc.z);
            double lts_new = norm(sv) / f.clight;
            double diff = fabs(lts_new - lts);
            lts = lts_new;
            if (!(diff > 1.0e-17 * fabs(f.et))) break;
        }
        {
            double ds = dt - (lts - f.lts0);
            sv = mk(f.S0[0] + f.VS[0] * ds - Tc.x, f.S0[1] + f.VS[1] * ds - Tc.y,
                    f.S0[2] + f.VS[2] * ds - Tc.z);
        }
        V3 s_b = unit(to_body(fs, r, sv));
        V3 n = unit(mk(p.x * fs.nw[0], p.y * fs.nw[1], p.z * fs.nw[2]));  // spice.surfnm
        V3 eu = unit(e_b);
        s.phase = vsep_unit(s_b, eu);
        s.incdnc = vsep_unit(n, s_b);
        s.emissn = vsep_unit(n, eu);
    }
}

// Body._azimuth_angle_from_gie_radians (body.py:2319-2332), radians in/out
__device__ __forceinline__ double azimuth_from_gie(double g, double i, double e) {
    double cg = cos(g), ci = cos(i), ce = cos(e);
    double a = cg - ce * ci;
    double b = sqrt(1.0 - ce * ce) * sqrt(1.0 - ci * ci);
    return kPi - acos(a / b);
}

// Body.local_solar_time_from_lon (body.py:2364-2398) -> spice.et2lst 'planetographic'
__device__ __forceinline__ double local_solar_time(const PMFrame &f, double lon_deg) {
    if (!isfinite(lon_deg)) return NAN;
    double ang = f.lon_sign * (lon_deg * kRpd) - f.sun_lon_lst;
    if (f.prograde == 0.0) ang = -ang;
    ang = fmod(ang, kTwoPi);
    if (ang < 0.0) ang += kTwoPi;
    double sec = ang * (86400.0 / kTwoPi) + 43200.0;
    if (sec >= 86400.0) sec -= 86400.0;
    double hr = floor(sec / 3600.0);
    sec -= hr * 3600.0;
    double mn = floor(sec / 60.0);
    sec -= mn * 60.0;
    double sc = floor(sec);
    return hr + mn / 60.0 + sc / 3600.0;
}

// SpiceBase.calculate_doppler_factor (base.py:524-551)
__device__ __forceinline__ double doppler_factor(const PMFrame &f, double rv) {
    double beta = rv / f.clight;
    return sqrt((1.0 + beta) / (1.0 - beta));
}

// BodyXY._xy2obsvec_norm (body_xy.py:375-377) + Body._angular2obsvec_norm
// (body.py:1363-1373)
__device__ __forceinline__ V3 xy2obsvec_norm(const PMFrame &f, double x, double y) {
    double ax = f.A[0] * x + f.A[1] * y + f.A[2];
    double ay = f.A[3] * x + f.A[4] * y + f.A[5];
    V3 v = radrec1(-((ax / 3600.0) * kRpd), (ay / 3600.0) * kRpd);
    return mtxv(f.M, v);
}

// Body._obsvec2angular (body.py:1345-1361), arcsec
__device__ __forceinline__ void obsvec2angular(const PMFrame &f, V3 ov, double &ax, double &ay) {
    if (!finite3(ov)) {
        ax = NAN;
        ay = NAN;
        return;
    }
    double ra, dec;
    recrad_angles(mxv(f.M, ov), ra, dec);
    double x = pymod_pos(-(ra * kDpr), 360.0);
    if (x > 180.0) x -= 360.0;
    ax = x * 3600.0;
    ay = (dec * kDpr) * 3600.0;
}

// BodyXY._obsvec2xy (body_xy.py:379-382)
__device__ __forceinline__ void obsvec2xy(const PMFrame &f, V3 ov, double &x, double &y) {
    double ax, ay;
    obsvec2angular(f, ov, ax, ay);
    x = f.Ainv[0] * ax + f.Ainv[1] * ay + f.Ainv[2];
    y = f.Ainv[3] * ax + f.Ainv[4] * ay + f.Ainv[5];
}

// Body._obsvec2km (body.py:1645-1650)
__device__ __forceinline__ void obsvec2km(const PMFrame &f, V3 ov, double &kx, double &ky) {
    double ax, ay;
    obsvec2angular(f, ov, ax, ay);
    kx = f.ang2km[0] * ax + f.ang2km[1] * ay;
    ky = f.ang2km[2] * ax + f.ang2km[3] * ay;
}

// Body._targvec2obsvec (body.py:917-948)
__device__ __forceinline__ V3 targvec2obsvec(const FrameS &fs, V3 tv) {
    const PMFrame &f = fs.f;
    V3 off = tv - ld3(f.sub_t);
    double dist_offset = norm(ld3(f.sub_ray) + off) - f.sub_dist;
    double sub_et = f.t_ref + f.sub_dt;
    double tt = sub_et - dist_offset / f.clight;
    Rot r = make_rot(fs, tt - f.t_ref);
    return ld3(f.sub_obs) + from_body(fs, r, off);
}

// Body._obsvec2targvec (body.py:972-1006), including its frame-mixing norm
__device__ __forceinline__ V3 obsvec2targvec(const FrameS &fs, V3 ov) {
    const PMFrame &f = fs.f;
    V3 off = ov - ld3(f.sub_obs);
    double dist_offset = norm(off - ld3(f.sub_ray)) - f.sub_dist;
    double sub_et = f.t_ref + f.sub_dt;
    double tt = sub_et - dist_offset / f.clight;
    Rot r = make_rot(fs, tt - f.t_ref);
    return ld3(f.sub_t) + to_body(fs, r, off);
}

// Body._ring_coordinates_from_obsvec(only_visible=False) (body.py:2577-2615)
__device__ __forceinline__ void ring_coordinates(const FrameS &fs, V3 ov, double &radius,
                                                 double &lon_deg, double &dist) {
    const PMFrame &f = fs.f;
    radius = lon_deg = dist = NAN;
    if (!finite3(ov)) return;
    double nd = dot(ld3(f.ring_n), ov);  // spice.inrypl, vertex at the origin
    if (nd == 0.0) return;
    double s = f.ring_c / nd;
    if (!(s > 0.0) || !isfinite(s)) return;
    V3 X = s * ov;
    V3 tv = obsvec2targvec(fs, X);
    double lon, lat, alt;
    recpgr(fs, tv, false, lon, lat, alt);
    radius = alt + f.r_eq;
    lon_deg = lon * kDpr;
    dist = norm(X);
}

// Body._limb_coordinates_from_obsvec (body.py:2081-2110)
__device__ __forceinline__ void limb_coordinates(const FrameS &fs, V3 ov, double &lon_deg,
                                                 double &lat_deg, double &dist) {
    const PMFrame &f = fs.f;
    lon_deg = lat_deg = dist = NAN;
    if (!finite3(ov)) return;
    double n = norm(ov);
    if (!(n > 0.0)) return;
    V3 u = (1.0 / n) * ov;  // spice.nplnpt(origin, ov, target centre)
    V3 P0 = ld3(f.P0);
    double t = dot(P0, u);
    V3 pn = t * u;
    double near_dist = norm(P0 - pn);
    V3 tv = obsvec2targvec(fs, pn);
    // spice.surfpt(origin, tv, a, b, c): radial surface point
    V3 x = mk(tv.x * fs.inv_r[0], tv.y * fs.inv_r[1], tv.z * fs.inv_r[2]);
    double xn = norm(x);
    if (!(xn > 0.0)) return;
    V3 sp = (1.0 / xn) * tv;
    double lon, lat, alt;
    recpgr(fs, sp, fs.biaxial != 0, lon, lat, alt);
    lon_deg = lon * kDpr;
    lat_deg = lat * kDpr;
    dist = near_dist - norm(sp);
}

}  // namespace pm
