/*
 * pm_oracle.c - CPU restatement (plain C, FP64) of PlanetMapper's per-pixel
 * geometry and mapping hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * path is the CUDA library in planetmapper_b200/csrc and never calls into here.
 *
 * The reference (ortk95/planetmapper v1.14.0) delegates its arithmetic to CSPICE
 * N0067 through spiceypy <= 8.1.2 (requirements.txt:5), to PROJ 9.x through
 * pyproj <= 3.7.2 and to FITPACK through scipy <= 1.18.0.  None of those sources are
 * under /root/reference, and spiceypy / pyproj are installed neither in the authoring
 * container nor on the GPU box, so the reference itself cannot be executed.  Each
 * routine below restates the published algorithm of the CSPICE / PROJ routine the
 * reference calls, citing the reference call site (file:line under
 * /root/reference/planetmapper) it stands in for.
 *
 * PARITY PIN: tests/test_oracle_golden.py checks this file against the reference's
 * own golden FITS outputs (tests/data/outputs/test_nav.fits etc., 26 backplanes,
 * maps and mapped cubes) and 16-digit known-answer literals of the reference's unit
 * tests; see DESIGN.md "Oracle".
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off)
 */
#include "pm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.14159265358979323846264338327950288
#define TWOPI (2.0 * PI)
#define HALFPI (0.5 * PI)
#define DPR (180.0 / PI)
#define RPD (PI / 180.0)

static const double kNaN = NAN;

/* ---------- small vector helpers ---------- */
static inline double dot3(const double a[3], const double b[3]) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static inline void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
/* spice.vnorm.  CSPICE pre-scales by the largest component as an overflow guard; the
 * magnitudes on this path (<= 1e10 km) cannot overflow, so the plain form is used. */
static inline double norm3(const double a[3]) {
    return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
}
static inline void mxv(const double m[9], const double v[3], double o[3]) {
    o[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
    o[1] = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
    o[2] = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
}
static inline void mtxv(const double m[9], const double v[3], double o[3]) {
    o[0] = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
    o[1] = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
    o[2] = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
}
static inline int finite3(const double v[3]) {
    return isfinite(v[0]) && isfinite(v[1]) && isfinite(v[2]);
}
/* Python's float % (result takes the sign of the divisor) */
static inline double pymod(double a, double m) {
    double r = fmod(a, m);
    if (r != 0.0 && ((r < 0.0) != (m < 0.0))) r += m;
    return r;
}

/* ---------- CSPICE primitives ---------- */

/* spice.recrad (base.py:902, body.py:1356): range, RA in [0,2pi), Dec */
static void recrad(const double v[3], double *range, double *ra, double *dec) {
    double big = fmax(fabs(v[0]), fmax(fabs(v[1]), fabs(v[2])));
    if (!(big > 0.0)) {
        *range = big; /* 0 or NaN */
        *ra = (big == 0.0) ? 0.0 : kNaN;
        *dec = (big == 0.0) ? 0.0 : kNaN;
        return;
    }
    double ib = 1.0 / big;
    double x = v[0] * ib, y = v[1] * ib, z = v[2] * ib;
    *range = big * sqrt(x * x + y * y + z * z);
    *dec = atan2(z, sqrt(x * x + y * y));
    double lon = (x == 0.0 && y == 0.0) ? 0.0 : atan2(y, x);
    if (lon < 0.0) lon += TWOPI;
    *ra = lon;
}

/* spice.radrec (body.py:967, :1369) */
static void radrec(double r, double ra, double dec, double o[3]) {
    o[0] = r * cos(ra) * cos(dec);
    o[1] = r * sin(ra) * cos(dec);
    o[2] = r * sin(dec);
}

/* spice.vsep: angle between two vectors, numerically stable form */
static double vsep(const double a[3], const double b[3]) {
    double na = norm3(a), nb = norm3(b);
    if (na == 0.0 || nb == 0.0) return 0.0;
    double ia = 1.0 / na, ib = 1.0 / nb;
    double u[3] = {a[0] * ia, a[1] * ia, a[2] * ia};
    double v[3] = {b[0] * ib, b[1] * ib, b[2] * ib};
    double d = dot3(u, v);
    if (d > 0.0) {
        double w[3] = {u[0] - v[0], u[1] - v[1], u[2] - v[2]};
        return 2.0 * asin(0.5 * norm3(w));
    } else if (d < 0.0) {
        double w[3] = {u[0] + v[0], u[1] + v[1], u[2] + v[2]};
        return PI - 2.0 * asin(0.5 * norm3(w));
    }
    return HALFPI;
}

/* J2000 -> body-fixed at epoch t_ref + dt: v_body = E(dt) R0 v, E = exp(-[omega]x dt).
 * Stands in for spice.pxform / pxfrm2 at a per-point epoch (body.py:940, :998 and
 * inside sincpt / illumf / spkcpt). */
typedef struct { double omega[3]; double k[3]; double wn; int valid; } SpinAxis;
static _Thread_local SpinAxis g_spin;
/* unit rotation axis and rate: a per-frame constant, cached per thread */
static const SpinAxis *spin_axis(const PMFrame *f) {
    SpinAxis *s = &g_spin;
    if (!s->valid || s->omega[0] != f->omega[0] || s->omega[1] != f->omega[1] ||
        s->omega[2] != f->omega[2]) {
        s->omega[0] = f->omega[0]; s->omega[1] = f->omega[1]; s->omega[2] = f->omega[2];
        s->wn = norm3(f->omega);
        double iw = s->wn > 0.0 ? 1.0 / s->wn : 0.0;
        s->k[0] = f->omega[0] * iw; s->k[1] = f->omega[1] * iw; s->k[2] = f->omega[2] * iw;
        s->valid = 1;
    }
    return s;
}
static void rot_to_body(const PMFrame *f, double dt, const double v[3], double o[3]) {
    double w[3];
    mxv(f->R0, v, w);
    const SpinAxis *ax = spin_axis(f);
    double wn = ax->wn;
    if (wn == 0.0 || dt == 0.0) {
        o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
        return;
    }
    /* Rodrigues rotation by -theta about k, written with sin(theta) and
     * 1 - cos(theta) = 2 sin^2(theta/2) so tiny angles lose no precision:
     * w - sin(th) (k x w) - (1 - cos(th)) (w - k (k.w)) */
    const double *k = ax->k;
    double sh = sin(0.5 * wn * dt), ch = cos(0.5 * wn * dt);
    double s = 2.0 * sh * ch, omc = 2.0 * sh * sh;
    double kxw[3];
    cross3(k, w, kxw);
    double kw = dot3(k, w);
    for (int i = 0; i < 3; i++) o[i] = w[i] - s * kxw[i] - omc * (w[i] - k[i] * kw);
}
/* body-fixed at epoch t_ref + dt -> J2000 */
static void rot_from_body(const PMFrame *f, double dt, const double u[3], double o[3]) {
    double w[3] = {u[0], u[1], u[2]};
    const SpinAxis *ax = spin_axis(f);
    double wn = ax->wn;
    if (wn != 0.0 && dt != 0.0) {
        const double *k = ax->k;
        double sh = sin(0.5 * wn * dt), ch = cos(0.5 * wn * dt);
        double s = 2.0 * sh * ch, omc = 2.0 * sh * sh;
        double kxu[3];
        cross3(k, u, kxu);
        double ku = dot3(k, u);
        for (int i = 0; i < 3; i++) w[i] = u[i] + s * kxu[i] - omc * (u[i] - k[i] * ku);
    }
    mtxv(f->R0, w, o);
}

/* target centre relative to the observer (J2000) at epoch t_ref + dt */
static void target_pos(const PMFrame *f, double dt, double P[3]) {
    for (int i = 0; i < 3; i++)
        P[i] = f->P0[i] + f->VT[i] * dt + 0.5 * f->AT[i] * dt * dt;
}

/* spice.surfpt: nearest intersection of ray (o, u) with ellipsoid (a,b,c).
 * Perpendicular-projection form (well conditioned for a distant observer).
 * margin (optional) receives |p_perp| - 1 in scaled space: < 0 hit, > 0 miss. */
static int surfpt(const double o[3], const double u[3], double a, double b, double c,
                  double p[3], double *margin) {
    double ia = 1.0 / a, ib = 1.0 / b, ic = 1.0 / c;
    double x[3] = {u[0] * ia, u[1] * ib, u[2] * ic};
    double y[3] = {o[0] * ia, o[1] * ib, o[2] * ic};
    double xn = norm3(x);
    if (margin) *margin = kNaN;
    if (!(xn > 0.0)) return 0;
    double ixn = 1.0 / xn;
    x[0] *= ixn; x[1] *= ixn; x[2] *= ixn;
    double yx = dot3(y, x);
    double pp[3] = {y[0] - yx * x[0], y[1] - yx * x[1], y[2] - yx * x[2]};
    double pmag = norm3(pp), ymag = norm3(y);
    if (margin) *margin = pmag - 1.0;
    double q[3];
    if (ymag > 1.0) {
        if (pmag > 1.0) return 0;
        if (yx > 0.0) return 0;
        if (pmag == 1.0) {
            q[0] = pp[0]; q[1] = pp[1]; q[2] = pp[2];
        } else {
            double sc = sqrt(fmax(0.0, 1.0 - pmag * pmag));
            for (int i = 0; i < 3; i++) q[i] = pp[i] - sc * x[i];
        }
    } else if (ymag == 1.0) {
        q[0] = y[0]; q[1] = y[1]; q[2] = y[2];
    } else {
        double sc = sqrt(fmax(0.0, 1.0 - pmag * pmag));
        for (int i = 0; i < 3; i++) q[i] = pp[i] + sc * x[i];
    }
    if (!isfinite(pmag)) return 0;
    p[0] = q[0] * a; p[1] = q[1] * b; p[2] = q[2] * c;
    return 1;
}

/* spice.sincpt('ELLIPSOID', ..., 'CN', observer, 'J2000', d) (body.py:1010-1020):
 * ray-ellipsoid intercept with the light time iterated on the intercept point.
 * Returns 1 if found.  lt/dt are the converged light time and epoch offset. */
#define PMO_MAXITR 10
static int sincpt(const PMFrame *f, const double d[3], double p[3], double *lt_out,
                  double *margin_out) {
    const double a = f->radii[0], b = f->radii[1], c = f->radii[2];
    double lt = f->lt0;
    double t = f->et - lt;
    if (margin_out) *margin_out = kNaN;
    for (int i = 0; i < PMO_MAXITR; i++) {
        double dt = t - f->t_ref;
        double P[3], negP[3], o[3], u[3];
        target_pos(f, dt, P);
        negP[0] = -P[0]; negP[1] = -P[1]; negP[2] = -P[2];
        rot_to_body(f, dt, negP, o);
        rot_to_body(f, dt, d, u);
        double margin;
        int found = surfpt(o, u, a, b, c, p, &margin);
        if (i == 0 && margin_out) *margin_out = margin;
        if (!found) return 0;
        double s[3] = {p[0] - o[0], p[1] - o[1], p[2] - o[2]};
        double lt_new = norm3(s) / f->clight;
        double t_new = f->et - lt_new;
        double ltdiff = fabs(t_new - t);
        t = t_new;
        lt = lt_new;
        if (!(ltdiff > 1.0e-17 * fabs(t))) break;
    }
    *lt_out = lt;
    return 1;
}

/* Geodetic coordinates of a point relative to the spheroid (re, re, re(1-f)):
 * the nearest-point construction of spice.recgeo, solved with Bowring's iteration
 * run to convergence.  lon east-positive in (-pi, pi]. */
static void recgeo(const double p[3], double re, double fl, double *lon, double *lat,
                   double *alt) {
    double rp = re - fl * re;
    double rho = sqrt(p[0] * p[0] + p[1] * p[1]);
    double z = p[2];
    *lon = (p[0] == 0.0 && p[1] == 0.0) ? 0.0 : atan2(p[1], p[0]);
    if (fl == 0.0) {
        double r = norm3(p);
        *lat = atan2(z, rho);
        *alt = r - re;
        return;
    }
    double e2 = 1.0 - (rp * rp) / (re * re);
    double ep2 = (re * re) / (rp * rp) - 1.0;
    double beta = atan2(re * z, rp * rho);
    double phi = 0.0;
    for (int i = 0; i < 12; i++) {
        double sb = sin(beta), cb = cos(beta);
        double nphi = atan2(z + ep2 * rp * sb * sb * sb, rho - e2 * re * cb * cb * cb);
        double nbeta = atan2((1.0 - fl) * sin(nphi), cos(nphi));
        int done = fabs(nphi - phi) <= 4.0e-16 * fmax(1.0, fabs(nphi)) && i > 0;
        phi = nphi;
        beta = nbeta;
        if (done) break;
    }
    double sp = sin(phi), cp = cos(phi);
    *lat = phi;
    *alt = rho * cp + z * sp - re * sqrt(1.0 - e2 * sp * sp);
}

/* spice.recpgr (body.py:1030, :2592): planetographic lon in [0,2pi), lat, alt */
static void recpgr(const PMFrame *f, const double p[3], double *lon, double *lat,
                   double *alt) {
    double l;
    recgeo(p, f->re, f->f, &l, lat, alt);
    l = f->lon_sign * l;
    if (l < 0.0) l += TWOPI;
    *lon = l;
}

/* spice.pgrrec (body.py:903): planetographic (radians) -> body-fixed */
static void pgrrec(const PMFrame *f, double lon, double lat, double alt, double o[3]) {
    double re = f->re, rp = f->re - f->f * f->re;
    double lam = f->lon_sign * lon;
    double clat = cos(lat), slat = sin(lat), clon = cos(lam), slon = sin(lam);
    double big = fmax(fabs(re * clat), fabs(rp * slat));
    double x = re * clat / big, y = rp * slat / big;
    double scale = 1.0 / sqrt(x * x + y * y);
    o[0] = scale * clon * x * re + alt * clat * clon;
    o[1] = scale * slon * x * re + alt * clat * slon;
    o[2] = scale * y * rp + alt * slat;
}

/* spice.reclat (body.py:2912) */
static void reclat(const double v[3], double *r, double *lon, double *lat) {
    double big = fmax(fabs(v[0]), fmax(fabs(v[1]), fabs(v[2])));
    if (big > 0.0) {
        double ib = 1.0 / big;
        double x = v[0] * ib, y = v[1] * ib, z = v[2] * ib;
        *r = big * sqrt(x * x + y * y + z * z);
        *lat = atan2(z, sqrt(x * x + y * y));
        *lon = (x == 0.0 && y == 0.0) ? 0.0 : atan2(y, x);
    } else {
        *r = 0.0; *lon = 0.0; *lat = 0.0;
    }
}

/* Light time from a body-fixed point p to the observer: the 'CN' solve inside
 * spice.spkcpt / spice.illumf (body.py:2833, :1925).  Returns lt, and dt = epoch
 * offset (et - lt) - t_ref of the point. */
static double point_light_time(const PMFrame *f, const double p[3], double lt_start,
                               double *dt_out, double X[3]) {
    double lt = lt_start;
    double dt = 0.0;
    for (int i = 0; i < PMO_MAXITR; i++) {
        double t = f->et - lt;
        dt = t - f->t_ref;
        double P[3], q[3];
        target_pos(f, dt, P);
        rot_from_body(f, dt, p, q);
        for (int k = 0; k < 3; k++) X[k] = P[k] + q[k];
        double lt_new = norm3(X) / f->clight;
        double diff = fabs(lt_new - lt);
        lt = lt_new;
        if (!(diff > 1.0e-17 * fabs(f->et))) break;
    }
    *dt_out = (f->et - lt) - f->t_ref;
    return lt;
}

typedef struct {
    double lt;        /* light time point -> observer */
    double pos[3];    /* point relative to observer, J2000 (spkcpt position) */
    double vel[3];    /* spkcpt velocity (with light-time-rate correction)   */
    double phase, incdnc, emissn;
    int visibl, lit;
} PointState;

/* spice.spkcpt (body.py:2830-2845) + spice.illumf (body.py:1915-1935) for a
 * body-fixed surface point p. */
static void point_state(const PMFrame *f, const double p[3], double lt_start,
                        int want_illum, PointState *s) {
    double dt, X[3];
    s->lt = point_light_time(f, p, lt_start, &dt, X);
    /* recompute X at the converged epoch */
    double P[3], q[3];
    target_pos(f, dt, P);
    rot_from_body(f, dt, p, q);
    for (int k = 0; k < 3; k++) s->pos[k] = P[k] + q[k];
    /* inertial velocity of the point: centre + rotation */
    double wxp[3], vrot[3], VX[3];
    cross3(f->omega, p, wxp);
    rot_from_body(f, dt, wxp, vrot);
    for (int k = 0; k < 3; k++) VX[k] = f->VT[k] + f->AT[k] * dt + vrot[k];
    double r = norm3(s->pos);
    double ir = 1.0 / r;
    double ph[3] = {s->pos[0] * ir, s->pos[1] * ir, s->pos[2] * ir};
    double rel[3] = {VX[0] - f->VO[0], VX[1] - f->VO[1], VX[2] - f->VO[2]};
    double dlt = dot3(ph, rel) / (f->clight + dot3(ph, VX));
    for (int k = 0; k < 3; k++) s->vel[k] = VX[k] * (1.0 - dlt) - f->VO[k];
    if (!want_illum) return;

    /* observer and Sun as seen from the point, body-fixed at the point's epoch */
    double negX[3] = {-s->pos[0], -s->pos[1], -s->pos[2]};
    double e_b[3];
    rot_to_body(f, dt, negX, e_b); /* point -> observer */
    /* Sun light time: S(t - lts) - X_ssb(t); all relative to target centre(t_ref) */
    double lts = f->lts0;
    double sv[3];
    double Tc[3];
    for (int k = 0; k < 3; k++)
        Tc[k] = f->VT[k] * dt + 0.5 * f->AT[k] * dt * dt + q[k]; /* point wrt T(t_ref) */
    for (int i = 0; i < PMO_MAXITR; i++) {
        double ds = dt - (lts - f->lts0);
        for (int k = 0; k < 3; k++) sv[k] = f->S0[k] + f->VS[k] * ds - Tc[k];
        double lts_new = norm3(sv) / f->clight;
        double diff = fabs(lts_new - lts);
        lts = lts_new;
        if (!(diff > 1.0e-17 * fabs(f->et))) break;
    }
    {
        double ds = dt - (lts - f->lts0);
        for (int k = 0; k < 3; k++) sv[k] = f->S0[k] + f->VS[k] * ds - Tc[k];
    }
    double s_b[3];
    rot_to_body(f, dt, sv, s_b);
    /* spice.surfnm */
    double a = f->radii[0], b = f->radii[1], c = f->radii[2];
    double m = fmin(a, fmin(b, c));
    double a1 = m / a, b1 = m / b, c1 = m / c;
    double n[3] = {p[0] * (a1 * a1), p[1] * (b1 * b1), p[2] * (c1 * c1)};
    double inn = 1.0 / norm3(n);
    n[0] *= inn; n[1] *= inn; n[2] *= inn;
    s->phase = vsep(s_b, e_b);
    s->incdnc = vsep(n, s_b);
    s->emissn = vsep(n, e_b);
    s->visibl = s->emissn < HALFPI;
    s->lit = s->incdnc < HALFPI;
}

/* Body._azimuth_angle_from_gie_radians (body.py:2319-2332) */
static double azimuth_from_gie(double g, double i, double e) {
    double a = cos(g) - cos(e) * cos(i);
    double b = sqrt(1.0 - cos(e) * cos(e)) * sqrt(1.0 - cos(i) * cos(i));
    return PI - acos(a / b);
}

/* Body.local_solar_time_from_lon (body.py:2364-2398) -> spice.et2lst(..., 'planetographic')
 * lon_deg: planetographic longitude in degrees. */
static double local_solar_time(const PMFrame *f, double lon_deg) {
    if (!isfinite(lon_deg)) return kNaN;
    double lam = f->lon_sign * (lon_deg * RPD);
    double ang = lam - f->sun_lon_lst;
    if (f->prograde == 0.0) ang = -ang;
    ang = fmod(ang, TWOPI);
    if (ang < 0.0) ang += TWOPI;
    double sec = ang * (86400.0 / TWOPI) + 43200.0;
    if (sec >= 86400.0) sec -= 86400.0;
    double hr = floor(sec / 3600.0);
    sec -= hr * 3600.0;
    double mn = floor(sec / 60.0);
    sec -= mn * 60.0;
    double sc = floor(sec);
    return hr + mn / 60.0 + sc / 3600.0;
}

/* SpiceBase.calculate_doppler_factor (base.py:524-551) */
static double doppler_factor(const PMFrame *f, double rv) {
    double beta = rv / f->clight;
    return sqrt((1.0 + beta) / (1.0 - beta));
}

/* Body._angular2obsvec_norm (body.py:1363-1373) after BodyXY._xy2obsvec_norm
 * (body_xy.py:375-377) */
static void xy2obsvec_norm(const PMFrame *f, double x, double y, double d[3]) {
    double ax = f->A[0] * x + f->A[1] * y + f->A[2] * 1.0;
    double ay = f->A[3] * x + f->A[4] * y + f->A[5] * 1.0;
    double v[3];
    radrec(1.0, -((ax / 3600.0) * RPD), (ay / 3600.0) * RPD, v);
    mtxv(f->M, v, d);
}

/* Body._obsvec2angular (body.py:1345-1361), arcsec */
static void obsvec2angular(const PMFrame *f, const double ov[3], double *ax, double *ay) {
    if (!finite3(ov)) {
        *ax = kNaN; *ay = kNaN;
        return;
    }
    double v[3], r, ra, dec;
    mxv(f->M, ov, v);
    recrad(v, &r, &ra, &dec);
    double x = pymod(-(ra * DPR), 360.0);
    if (x > 180.0) x -= 360.0;
    *ax = x * 3600.0;
    *ay = (dec * DPR) * 3600.0;
}

/* BodyXY._obsvec2xy (body_xy.py:379-382) */
static void obsvec2xy(const PMFrame *f, const double ov[3], double *x, double *y) {
    double ax, ay;
    obsvec2angular(f, ov, &ax, &ay);
    *x = f->Ainv[0] * ax + f->Ainv[1] * ay + f->Ainv[2] * 1.0;
    *y = f->Ainv[3] * ax + f->Ainv[4] * ay + f->Ainv[5] * 1.0;
}

/* Body._radec2obsvec_norm (body.py:964-970), degrees in */
static void radec_deg2obsvec_norm(double ra_deg, double dec_deg, double d[3]) {
    if (!(isfinite(ra_deg) && isfinite(dec_deg))) {
        d[0] = d[1] = d[2] = kNaN;
        return;
    }
    radrec(1.0, ra_deg * RPD, dec_deg * RPD, d);
}

/* Body._obsvec2km (body.py:1645-1650) */
static void obsvec2km(const PMFrame *f, const double ov[3], double *kx, double *ky) {
    double ax, ay;
    obsvec2angular(f, ov, &ax, &ay);
    *kx = f->ang2km[0] * ax + f->ang2km[1] * ay;
    *ky = f->ang2km[2] * ax + f->ang2km[3] * ay;
}

/* Body._targvec2obsvec (body.py:917-948) */
static void targvec2obsvec(const PMFrame *f, const double tv[3], double ov[3]) {
    double off[3] = {tv[0] - f->sub_t[0], tv[1] - f->sub_t[1], tv[2] - f->sub_t[2]};
    double w[3] = {f->sub_ray[0] + off[0], f->sub_ray[1] + off[1], f->sub_ray[2] + off[2]};
    double dist_offset = norm3(w) - f->sub_dist;
    double sub_et = f->t_ref + f->sub_dt;
    double tt = sub_et - dist_offset / f->clight;
    double q[3];
    rot_from_body(f, tt - f->t_ref, off, q);
    for (int k = 0; k < 3; k++) ov[k] = f->sub_obs[k] + q[k];
}

/* Body._obsvec2targvec (body.py:972-1006), including its frame-mixing norm */
static void obsvec2targvec(const PMFrame *f, const double ov[3], double tv[3]) {
    double off[3] = {ov[0] - f->sub_obs[0], ov[1] - f->sub_obs[1], ov[2] - f->sub_obs[2]};
    double w[3] = {-f->sub_ray[0] + off[0], -f->sub_ray[1] + off[1], -f->sub_ray[2] + off[2]};
    double dist_offset = norm3(w) - f->sub_dist;
    double sub_et = f->t_ref + f->sub_dt;
    double tt = sub_et - dist_offset / f->clight;
    double q[3];
    rot_to_body(f, tt - f->t_ref, off, q);
    for (int k = 0; k < 3; k++) tv[k] = f->sub_t[k] + q[k];
}

/* Body._ring_coordinates_from_obsvec(only_visible=False) (body.py:2577-2615) */
static void ring_coordinates(const PMFrame *f, const double ov[3], double *radius,
                             double *lon_deg, double *dist) {
    *radius = *lon_deg = *dist = kNaN;
    if (!finite3(ov)) return;
    /* spice.inrypl with vertex at the origin */
    double nd = dot3(f->ring_n, ov);
    if (nd == 0.0) return;
    double s = f->ring_c / nd;
    if (!(s > 0.0) || !isfinite(s)) return;
    double X[3] = {s * ov[0], s * ov[1], s * ov[2]};
    double tv[3], lon, lat, alt;
    obsvec2targvec(f, X, tv);
    recpgr(f, tv, &lon, &lat, &alt);
    *radius = alt + f->r_eq;
    *lon_deg = lon * DPR;
    *dist = norm3(X);
}

/* Body._limb_coordinates_from_obsvec (body.py:2081-2110) */
static void limb_coordinates(const PMFrame *f, const double ov[3], double *lon_deg,
                             double *lat_deg, double *dist) {
    *lon_deg = *lat_deg = *dist = kNaN;
    if (!finite3(ov)) return;
    /* spice.nplnpt(origin, ov, target centre) */
    double n = norm3(ov);
    if (!(n > 0.0)) return;
    double u[3] = {ov[0] / n, ov[1] / n, ov[2] / n};
    double t = dot3(f->P0, u);
    double pn[3] = {t * u[0], t * u[1], t * u[2]};
    double dv[3] = {f->P0[0] - pn[0], f->P0[1] - pn[1], f->P0[2] - pn[2]};
    double near_dist = norm3(dv);
    double tv[3];
    obsvec2targvec(f, pn, tv);
    /* spice.surfpt(origin, tv, a, b, c): radial surface point */
    double a = f->radii[0], b = f->radii[1], c = f->radii[2];
    double x[3] = {tv[0] / a, tv[1] / b, tv[2] / c};
    double xn = norm3(x);
    if (!(xn > 0.0)) return;
    double sp[3] = {tv[0] / xn, tv[1] / xn, tv[2] / xn};
    double lon, lat, alt;
    recpgr(f, sp, &lon, &lat, &alt);
    *lon_deg = lon * DPR;
    *lat_deg = lat * DPR;
    *dist = near_dist - norm3(sp);
}

/* ---------- per-pixel image direction ---------- */
typedef struct { double v[PM_N_PLANES]; double margin; } PixelOut;

static void pixel_backplanes(const PMFrame *f, double x, double y, uint64_t mask,
                             PixelOut *o) {
    for (int k = 0; k < PM_N_PLANES; k++) o->v[k] = kNaN;
    o->margin = kNaN;
    o->v[PM_PIXEL_X] = x; /* BodyXY.get_x_img / get_y_img (body_xy.py:3494-3531) */
    o->v[PM_PIXEL_Y] = y;

    /* BodyXY._get_radec_img (body_xy.py:3413-3418) */
    double d[3], r, ra, dec;
    xy2obsvec_norm(f, x, y, d);
    const uint64_t km_mask = (1ull << PM_KM_X) | (1ull << PM_KM_Y) | (1ull << PM_ANGULAR_X) |
                             (1ull << PM_ANGULAR_Y);
    const uint64_t radec_users =
        (1ull << PM_RA) | (1ull << PM_DEC) | km_mask | (1ull << PM_LIMB_DISTANCE) |
        (1ull << PM_LIMB_LON_GRAPHIC) | (1ull << PM_LIMB_LAT_GRAPHIC) | (1ull << PM_RING_RADIUS) |
        (1ull << PM_RING_LON_GRAPHIC) | (1ull << PM_RING_DISTANCE);
    double d2[3] = {kNaN, kNaN, kNaN};
    if (mask & radec_users) {
        recrad(d, &r, &ra, &dec);
        double ra_deg = ra * DPR, dec_deg = dec * DPR;
        o->v[PM_RA] = ra_deg;
        o->v[PM_DEC] = dec_deg;
        /* _get_obsvec_norm_img (body_xy.py:3263-3272): RA/Dec degrees -> unit vector */
        radec_deg2obsvec_norm(ra_deg, dec_deg, d2);
    }

    /* BodyXY._get_km_xy_img (body_xy.py:3547-3553), angular (:3611-3656) */
    if (mask & km_mask) {
        double kx, ky;
        obsvec2km(f, d2, &kx, &ky);
        o->v[PM_KM_X] = kx;
        o->v[PM_KM_Y] = ky;
        o->v[PM_ANGULAR_X] = kx / f->km_per_arcsec;
        o->v[PM_ANGULAR_Y] = ky / f->km_per_arcsec;
    }

    /* BodyXY._get_targvec_img (body_xy.py:3197-3225) */
    int on_disc = 0;
    double p[3], lt = 0.0;
    double dx = x - f->x0, dy = y - f->y0;
    if (!(f->optimize_speed != 0.0 && (dx * dx + dy * dy) > f->r_cut2)) {
        on_disc = sincpt(f, d, p, &lt, &o->margin);
    }
    const uint64_t surf_mask =
        (1ull << PM_LON_GRAPHIC) | (1ull << PM_LAT_GRAPHIC) | (1ull << PM_LON_CENTRIC) |
        (1ull << PM_LAT_CENTRIC) | (1ull << PM_PHASE) | (1ull << PM_INCIDENCE) |
        (1ull << PM_EMISSION) | (1ull << PM_AZIMUTH) | (1ull << PM_LOCAL_SOLAR_TIME) |
        (1ull << PM_DISTANCE) | (1ull << PM_RADIAL_VELOCITY) | (1ull << PM_DOPPLER) |
        (1ull << PM_RING_RADIUS) | (1ull << PM_RING_LON_GRAPHIC) | (1ull << PM_RING_DISTANCE);
    double distance = kNaN;
    if (on_disc && (mask & surf_mask)) {
        double lon, lat, alt;
        recpgr(f, p, &lon, &lat, &alt); /* _get_lonlat_img (body_xy.py:3284-3288) */
        o->v[PM_LON_GRAPHIC] = lon * DPR;
        o->v[PM_LAT_GRAPHIC] = lat * DPR;
        double rr, clon, clat;
        reclat(p, &rr, &clon, &clat); /* _get_lonlat_centric_img (body_xy.py:3349) */
        o->v[PM_LON_CENTRIC] = clon * DPR;
        o->v[PM_LAT_CENTRIC] = clat * DPR;
        PointState s;
        point_state(f, p, lt, 1, &s);
        /* _get_illumination_gie_img (body_xy.py:3661-3665) */
        double g = s.phase * DPR, i = s.incdnc * DPR, e = s.emissn * DPR;
        o->v[PM_PHASE] = g;
        o->v[PM_INCIDENCE] = i;
        o->v[PM_EMISSION] = e;
        /* get_azimuth_angle_img (body_xy.py:3744-3762): deg -> rad -> formula -> deg */
        o->v[PM_AZIMUTH] = azimuth_from_gie(g * RPD, i * RPD, e * RPD) * DPR;
        o->v[PM_LOCAL_SOLAR_TIME] = local_solar_time(f, o->v[PM_LON_GRAPHIC]);
        distance = s.lt * f->clight; /* get_distance_img (body_xy.py:3870-3880) */
        o->v[PM_DISTANCE] = distance;
        /* get_radial_velocity_img (body_xy.py:3898-3913), body.py:2847-2853 */
        double rn = sqrt(s.pos[0] * s.pos[0] + s.pos[1] * s.pos[1] + s.pos[2] * s.pos[2]);
        double rv = s.vel[0] * (s.pos[0] / rn) + s.vel[1] * (s.pos[1] / rn) +
                    s.vel[2] * (s.pos[2] / rn);
        o->v[PM_RADIAL_VELOCITY] = rv;
        o->v[PM_DOPPLER] = doppler_factor(f, rv);
    }
    /* _get_limb_coordinate_imgs (body_xy.py:3967-3975); uses _get_obsvec_norm_img,
     * i.e. the RA/Dec degree round trip (body_xy.py:3263-3272) */
    if (mask & ((1ull << PM_LIMB_DISTANCE) | (1ull << PM_LIMB_LON_GRAPHIC) |
                (1ull << PM_LIMB_LAT_GRAPHIC))) {
        limb_coordinates(f, d2, &o->v[PM_LIMB_LON_GRAPHIC], &o->v[PM_LIMB_LAT_GRAPHIC],
                         &o->v[PM_LIMB_DISTANCE]);
    }
    /* _get_ring_plane_coordinate_imgs (body_xy.py:4061-4085) */
    if (mask & ((1ull << PM_RING_RADIUS) | (1ull << PM_RING_LON_GRAPHIC) |
                (1ull << PM_RING_DISTANCE))) {
        double rad, rl, rd;
        ring_coordinates(f, d2, &rad, &rl, &rd);
        if (rd > distance) { /* NaN distance compares False: off-disc keeps values */
            rad = rl = rd = kNaN;
        }
        o->v[PM_RING_RADIUS] = rad;
        o->v[PM_RING_LON_GRAPHIC] = rl;
        o->v[PM_RING_DISTANCE] = rd;
    }
}

static int popcount64(uint64_t m) {
    int n = 0;
    while (m) { n += (int)(m & 1ull); m >>= 1; }
    return n;
}

int pmo_backplanes_img(const PMFrame *f, int nx, int ny, uint64_t mask, double *out,
                       double *margin) {
    if (!f || !out || nx <= 0 || ny <= 0) return PM_ERR_BAD_ARG;
    mask &= PM_ALL_PLANES;
    const long npx = (long)nx * ny;
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < npx; i++) {
        int y = (int)(i / nx), x = (int)(i % nx);
        PixelOut o;
        pixel_backplanes(f, (double)x, (double)y, mask, &o);
        int slot = 0;
        for (int k = 0; k < PM_N_PLANES; k++) {
            if (mask & (1ull << k)) {
                out[(long)slot * npx + i] = o.v[k];
                slot++;
            }
        }
        if (margin) margin[i] = o.margin;
    }
    (void)popcount64;
    return PM_OK;
}

/* ---------- map direction ---------- */
static void cell_backplanes(const PMFrame *f, double lon_deg, double lat_deg,
                            uint64_t mask, PixelOut *o) {
    for (int k = 0; k < PM_N_PLANES; k++) o->v[k] = kNaN;
    o->margin = kNaN;
    /* BodyXY._get_lonlat_map (body_xy.py:3293-3300): lons % 360, non-finite -> NaN */
    double lonm = isfinite(lon_deg) ? pymod(lon_deg, 360.0) : kNaN;
    double latm = isfinite(lat_deg) ? lat_deg : kNaN;
    o->v[PM_LON_GRAPHIC] = lonm;
    o->v[PM_LAT_GRAPHIC] = latm;
    /* _get_targvec_map (body_xy.py:3230-3238): skipped when lon is NaN; a NaN lat
     * gives a NaN targvec through lonlat2targvec (body.py:901-902) */
    if (isnan(lonm)) return;
    /* get_local_solar_time_map (body_xy.py:3812-3828) */
    o->v[PM_LOCAL_SOLAR_TIME] = local_solar_time(f, lonm);
    if (!isfinite(latm)) return;
    double tv[3];
    pgrrec(f, lonm * RPD, latm * RPD, 0.0, tv);
    if (isnan(tv[0])) return;
    double rr, clon, clat;
    reclat(tv, &rr, &clon, &clat); /* _get_lonlat_centric_map (body_xy.py:3357-3364) */
    o->v[PM_LON_CENTRIC] = clon * DPR;
    o->v[PM_LAT_CENTRIC] = clat * DPR;
    PointState s;
    point_state(f, tv, f->lt0, 1, &s);
    o->margin = s.emissn - HALFPI;
    /* _get_illumf_map (body_xy.py:3671-3675) */
    double g = s.phase * DPR, i = s.incdnc * DPR, e = s.emissn * DPR;
    o->v[PM_PHASE] = g;
    o->v[PM_INCIDENCE] = i;
    o->v[PM_EMISSION] = e;
    o->v[PM_AZIMUTH] = azimuth_from_gie(g * RPD, i * RPD, e * RPD) * DPR;
    double distance = s.lt * f->clight; /* get_distance_map (body_xy.py:3883-3893) */
    o->v[PM_DISTANCE] = distance;
    double rn = sqrt(s.pos[0] * s.pos[0] + s.pos[1] * s.pos[1] + s.pos[2] * s.pos[2]);
    double rv = s.vel[0] * (s.pos[0] / rn) + s.vel[1] * (s.pos[1] / rn) +
                s.vel[2] * (s.pos[2] / rn);
    o->v[PM_RADIAL_VELOCITY] = rv; /* get_radial_velocity_map (body_xy.py:3917-3936) */
    o->v[PM_DOPPLER] = doppler_factor(f, rv);

    double ov[3];
    targvec2obsvec(f, tv, ov); /* _get_obsvec_map (body_xy.py:3275-3280) */
    if (s.visibl) {
        /* _get_radec_map (body_xy.py:3423-3432) */
        double r, ra, dec;
        recrad(ov, &r, &ra, &dec);
        double ra_deg = ra * DPR, dec_deg = dec * DPR;
        o->v[PM_RA] = ra_deg;
        o->v[PM_DEC] = dec_deg;
        double d2[3];
        radec_deg2obsvec_norm(ra_deg, dec_deg, d2);
        /* _get_xy_map (body_xy.py:3482-3491) */
        double x, y;
        obsvec2xy(f, d2, &x, &y);
        if ((-0.5 < x && x < f->nx - 0.5) && (-0.5 < y && y < f->ny - 0.5)) {
            o->v[PM_PIXEL_X] = x;
            o->v[PM_PIXEL_Y] = y;
        }
        /* _get_km_xy_map (body_xy.py:3557-3565) */
        double kx, ky;
        obsvec2km(f, d2, &kx, &ky);
        o->v[PM_KM_X] = kx;
        o->v[PM_KM_Y] = ky;
        o->v[PM_ANGULAR_X] = kx / f->km_per_arcsec;
        o->v[PM_ANGULAR_Y] = ky / f->km_per_arcsec;
    }
    /* the reference tests `lit` (illumf index 4) here, not `visibl`
     * (body_xy.py:3981-3986, :4097-4106); reproduced as is */
    if (s.lit) {
        if (mask & ((1ull << PM_LIMB_DISTANCE) | (1ull << PM_LIMB_LON_GRAPHIC) |
                    (1ull << PM_LIMB_LAT_GRAPHIC))) {
            limb_coordinates(f, ov, &o->v[PM_LIMB_LON_GRAPHIC],
                             &o->v[PM_LIMB_LAT_GRAPHIC], &o->v[PM_LIMB_DISTANCE]);
        }
        if (mask & ((1ull << PM_RING_RADIUS) | (1ull << PM_RING_LON_GRAPHIC) |
                    (1ull << PM_RING_DISTANCE))) {
            double rad, rl, rd;
            ring_coordinates(f, ov, &rad, &rl, &rd);
            if (rd > distance) rad = rl = rd = kNaN;
            o->v[PM_RING_RADIUS] = rad;
            o->v[PM_RING_LON_GRAPHIC] = rl;
            o->v[PM_RING_DISTANCE] = rd;
        }
    }
}

int pmo_backplanes_map(const PMFrame *f, const double *lon, const double *lat,
                       int64_t n, uint64_t mask, double *out, double *margin) {
    if (!f || !out || !lon || !lat || n < 0) return PM_ERR_BAD_ARG;
    mask &= PM_ALL_PLANES;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++) {
        PixelOut o;
        cell_backplanes(f, lon[i], lat[i], mask, &o);
        int slot = 0;
        for (int k = 0; k < PM_N_PLANES; k++) {
            if (mask & (1ull << k)) {
                out[(int64_t)slot * n + i] = o.v[k];
                slot++;
            }
        }
        if (margin) margin[i] = o.margin;
    }
    return PM_OK;
}

/* BodyXY._xy2lonlat (body_xy.py:482-496) -> Body._obsvec_norm2lonlat
 * (body.py:1058-1081), planetographic degrees; miss -> NaN and counted */
int pmo_xy2lonlat(const PMFrame *f, const double *x, const double *y, int64_t n,
                  double *lon, double *lat, int64_t *n_missed) {
    if (!f || !x || !y || !lon || !lat || n < 0) return PM_ERR_BAD_ARG;
    int64_t missed = 0;
#pragma omp parallel for reduction(+ : missed)
    for (int64_t i = 0; i < n; i++) {
        double d[3], p[3], lt;
        lon[i] = lat[i] = kNaN;
        if (!(isfinite(x[i]) && isfinite(y[i]))) continue;
        xy2obsvec_norm(f, x[i], y[i], d);
        if (!sincpt(f, d, p, &lt, NULL)) {
            missed++;
            continue;
        }
        double lo, la, al;
        recpgr(f, p, &lo, &la, &al);
        lon[i] = lo * DPR;
        lat[i] = la * DPR;
    }
    if (n_missed) *n_missed = missed;
    return PM_OK;
}

/* spice.latsrf on the ELLIPSOID (body.py:2970-2978): the surface point in the direction of
 * planetocentric lon / lat (radians, east positive) */
static void latsrf(const PMFrame *f, double lon, double lat, double p[3]) {
    double d[3] = {cos(lat) * cos(lon), cos(lat) * sin(lon), sin(lat)};
    double a = f->radii[0], b = f->radii[1], c = f->radii[2];
    double s = 1.0 / sqrt((d[0] / a) * (d[0] / a) + (d[1] / b) * (d[1] / b) + (d[2] / c) * (d[2] / c));
    for (int k = 0; k < 3; k++) p[k] = s * d[k];
}

/* Body._test_if_targvec_visible(on_surface=False) (body.py:2131-2150): cast the ray
 * observer -> point; hidden iff it meets the surface and the surface is nearer */
static int raycast_visible(const PMFrame *f, const double tv[3]) {
    double ov[3], p[3], lt;
    targvec2obsvec(f, tv, ov);
    if (!sincpt(f, ov, p, &lt, NULL)) return 1;
    PointState si, sp;
    point_state(f, p, f->lt0, 0, &si);
    point_state(f, tv, f->lt0, 0, &sp);
    return sp.lt < si.lt;
}

/* BodyXY._lonlat2xy (body_xy.py:544-560) -> Body._lonlat2obsvec (body.py:1039-1056):
 * alt == 0 visibility through illumf.visibl (body.py:2124-2130), alt != 0 through the ray
 * cast (body.py:2131-2150); PM_FLAG_PLANETOCENTRIC converts the inputs with
 * Body._centric2graphic_lonlat (body.py:2966-2982) */
int pmo_lonlat2xy_alt(const PMFrame *f, const double *lon, const double *lat, int64_t n, double alt,
                      uint32_t flags, double *x, double *y) {
    if (!f || !x || !y || !lon || !lat || n < 0) return PM_ERR_BAD_ARG;
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++) {
        x[i] = y[i] = kNaN;
        if (!(isfinite(lon[i]) && isfinite(lat[i]))) continue;
        double lo = lon[i] * RPD, la = lat[i] * RPD;
        if (flags & PM_FLAG_PLANETOCENTRIC) {
            double sp[3], al;
            latsrf(f, lo, la, sp);
            if (alt == 0.0) {
                recpgr(f, sp, &lo, &la, &al);
            } else {
                /* Body.targvec2lonlat(targvec, alt=alt) (body.py:1279-1283): recpgr against the
                 * spheroid with every radius raised by alt (_AdjustedSurfaceAltitude, :210-229) */
                double re_a = f->re + alt, rp_a = (f->re - f->f * f->re) + alt;
                recgeo(sp, re_a, (re_a - rp_a) / re_a, &lo, &la, &al);
                lo = f->lon_sign * lo;
                if (lo < 0.0) lo += TWOPI;
            }
            /* Body.targvec2lonlat returns degrees; _lonlat2obsvec converts back */
            lo = (lo * DPR) * RPD;
            la = (la * DPR) * RPD;
        }
        double tv[3];
        pgrrec(f, lo, la, alt, tv);
        if (flags & PM_FLAG_NOT_VISIBLE_NAN) {
            if (alt == 0.0) {
                PointState s;
                point_state(f, tv, f->lt0, 1, &s);
                if (!s.visibl) continue;
            } else if (!raycast_visible(f, tv)) {
                continue;
            }
        }
        double ov[3];
        targvec2obsvec(f, tv, ov);
        obsvec2xy(f, ov, &x[i], &y[i]);
    }
    return PM_OK;
}
int pmo_lonlat2xy(const PMFrame *f, const double *lon, const double *lat, int64_t n,
                  uint32_t flags, double *x, double *y) {
    return pmo_lonlat2xy_alt(f, lon, lat, n, 0.0, flags, x, y);
}

/* ---------- generic point transforms ---------- */
/* Body._obsvec2angular (body.py:1345-1361) with an explicit obsvec -> angular matrix
 * (Body._get_obsvec2angular_matrix for any origin_ra / origin_dec / coordinate_rotation) */
static void obsvec2angular_m(const double M[9], const double ov[3], double *ax, double *ay) {
    if (!finite3(ov)) {
        *ax = kNaN; *ay = kNaN;
        return;
    }
    double v[3], r, ra, dec;
    mxv(M, ov, v);
    recrad(v, &r, &ra, &dec);
    double x = pymod(-(ra * DPR), 360.0);
    if (x > 180.0) x -= 360.0;
    *ax = x * 3600.0;
    *ay = (dec * DPR) * 3600.0;
}
/* Body._angular2obsvec_norm (body.py:1363-1373) */
static void angular2obsvec_m(const double M[9], double ax, double ay, double d[3]) {
    double v[3];
    radrec(1.0, -((ax / 3600.0) * RPD), (ay / 3600.0) * RPD, v);
    mtxv(M, v, d);
}
/* Body._lonlat2obsvec (body.py:1039-1056); returns 0 when hidden.  lo / la: the planetographic radians the
 * point was built from, tv the body-fixed point (the middle of pmo_lonlat2xy_alt above) */
static int lonlat2obsvec(const PMFrame *f, double lon_deg, double lat_deg, double alt, uint32_t flags,
                         double ov[3], double *lo_out, double *la_out, double tv[3]) {
    double lo = lon_deg * RPD, la = lat_deg * RPD;
    if (flags & PM_FLAG_PLANETOCENTRIC) {
        double sp[3], al;
        latsrf(f, lo, la, sp);
        if (alt == 0.0) {
            recpgr(f, sp, &lo, &la, &al);
        } else {
            double re_a = f->re + alt, rp_a = (f->re - f->f * f->re) + alt;
            recgeo(sp, re_a, (re_a - rp_a) / re_a, &lo, &la, &al);
            lo = f->lon_sign * lo;
            if (lo < 0.0) lo += TWOPI;
        }
        lo = (lo * DPR) * RPD;
        la = (la * DPR) * RPD;
    }
    *lo_out = lo;
    *la_out = la;
    pgrrec(f, lo, la, alt, tv);
    if (flags & PM_FLAG_NOT_VISIBLE_NAN) {
        if (alt == 0.0) {
            PointState s;
            point_state(f, tv, f->lt0, 1, &s);
            if (!s.visibl) return 0;
        } else if (!raycast_visible(f, tv)) {
            return 0;
        }
    }
    targvec2obsvec(f, tv, ov);
    return 1;
}

/* One point between any two of the coordinate systems (include/pm_b200.h PM_COORD_*), every pair through
 * the observer-frame vector exactly as the reference composes them:
 *   BodyXY._xy2radec / _radec2xy / _xy2km / _km2xy / _xy2angular / _angular2xy (body_xy.py:409-676),
 *   Body._lonlat2radec / _radec2lonlat (body.py:1131-1221), _radec2angular / _angular2radec (:1425-1478),
 *   _angular2lonlat / _lonlat2angular (:1534-1623), _km2radec / _radec2km (:1672-1701),
 *   _km2lonlat / _lonlat2km (:1752-1830), _km2angular / _angular2km (:1861-1900),
 *   graphic2centric_lonlat (:2915-2947), centric2graphic_lonlat (:2949-2982).
 * aux13 = obsvec -> angular matrix (9) + km -> angular matrix (4). Returns 0 for a ray that misses. */
static int transform_one(const PMFrame *f, const double *aux13, int src, int dst, double a, double b, double alt,
                         uint32_t flags, double *oa, double *ob) {
    const double *Mc = aux13, *km2ang = aux13 + 9;
    *oa = *ob = kNaN;
    if (!(isfinite(a) && isfinite(b))) return 1;
    double ov[3];
    if (src == PM_COORD_LONLAT) {
        double lo, la, tv[3];
        uint32_t fl = flags;
        if (dst == PM_COORD_LONLAT || dst == PM_COORD_CENTRIC) fl &= ~PM_FLAG_NOT_VISIBLE_NAN;
        int vis = lonlat2obsvec(f, a, b, alt, fl, ov, &lo, &la, tv);
        if (dst == PM_COORD_LONLAT) {
            *oa = lo * DPR;
            *ob = la * DPR;
            return 1;
        }
        if (dst == PM_COORD_CENTRIC) {
            double r, lc, bc;
            reclat(tv, &r, &lc, &bc);
            *oa = lc * DPR;
            *ob = bc * DPR;
            return 1;
        }
        if (!vis) return 1;
    } else if (src == PM_COORD_RADEC) {
        radec_deg2obsvec_norm(a, b, ov);
    } else if (src == PM_COORD_ANGULAR) {
        angular2obsvec_m(Mc, a, b, ov);
    } else if (src == PM_COORD_KM) {
        angular2obsvec_m(f->M, km2ang[0] * a + km2ang[1] * b, km2ang[2] * a + km2ang[3] * b, ov);
    } else {
        xy2obsvec_norm(f, a, b, ov);
    }
    if (!finite3(ov)) return 1;
    if (dst == PM_COORD_RADEC) {
        double r, ra, dec;
        recrad(ov, &r, &ra, &dec);
        *oa = ra * DPR;
        *ob = dec * DPR;
    } else if (dst == PM_COORD_ANGULAR) {
        obsvec2angular_m(Mc, ov, oa, ob);
    } else if (dst == PM_COORD_KM) {
        obsvec2km(f, ov, oa, ob);
    } else if (dst == PM_COORD_XY) {
        obsvec2xy(f, ov, oa, ob);
    } else {
        /* Body._obsvec_norm2lonlat (body.py:1058-1081); f carries the radii raised by alt */
        double p[3], lt, lo, la, al;
        if (!sincpt(f, ov, p, &lt, NULL)) return 0;
        recpgr(f, p, &lo, &la, &al);
        *oa = lo * DPR;
        *ob = la * DPR;
        if ((flags & PM_FLAG_PLANETOCENTRIC) || dst == PM_COORD_CENTRIC) {
            /* graphic2centric_lonlat(lon, lat, alt=alt) INSIDE the altitude adjustment (body.py:1066-1080) */
            double tv[3], r, lc, bc;
            pgrrec(f, (*oa) * RPD, (*ob) * RPD, alt, tv);
            reclat(tv, &r, &lc, &bc);
            *oa = lc * DPR;
            *ob = bc * DPR;
        }
    }
    return 1;
}

int pmo_transform(const PMFrame *f, int src, int dst, const double *a, const double *b, int64_t n, double alt,
                  uint32_t flags, const double *aux13, double *out_a, double *out_b, int64_t *n_missed) {
    if (!f || !a || !b || !out_a || !out_b || n < 0) return PM_ERR_BAD_ARG;
    double aux[13];
    if (aux13) {
        for (int i = 0; i < 13; i++) aux[i] = aux13[i];
    } else {
        for (int i = 0; i < 9; i++) aux[i] = f->M[i];
        double det = f->ang2km[0] * f->ang2km[3] - f->ang2km[1] * f->ang2km[2];
        aux[9] = f->ang2km[3] / det;
        aux[10] = -f->ang2km[1] / det;
        aux[11] = -f->ang2km[2] / det;
        aux[12] = f->ang2km[0] / det;
    }
    int64_t missed = 0;
#pragma omp parallel for reduction(+ : missed)
    for (int64_t i = 0; i < n; i++)
        if (!transform_one(f, aux, src, dst, a[i], b[i], alt, flags, &out_a[i], &out_b[i])) missed++;
    if (n_missed) *n_missed = missed;
    return PM_OK;
}

/* ---------- projections ---------- */
/* PROJ's spherical orthographic inverse (ortho_s_inverse); x, y in sphere radii.
 * Returns 0 when the point is outside the projection. */
static int ortho_sph_inverse(double x, double y, double phi0, double *phi, double *lam) {
    double sinphi0 = sin(phi0), cosphi0 = cos(phi0);
    double rh = hypot(x, y), sinc = rh;
    if (sinc > 1.0) {
        if (sinc - 1.0 > 1e-10) return 0;
        sinc = 1.0;
    }
    double cosc = sqrt(1.0 - sinc * sinc);
    if (fabs(rh) <= 1e-10) {
        *phi = phi0;
        *lam = 0.0;
        return 1;
    }
    double p;
    int oblique_or_equit = 1;
    if (fabs(fabs(phi0) - HALFPI) < 1e-10) {
        oblique_or_equit = 0;
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
            p = cosc * sinphi0 + y * sinc * cosphi0 / rh;
            y = (cosc - sinphi0 * p) * rh;
            x *= sinc * cosphi0;
        }
        p = (fabs(p) >= 1.0) ? (p < 0.0 ? -HALFPI : HALFPI) : asin(p);
    }
    *phi = p;
    if (y == 0.0 && oblique_or_equit)
        *lam = (x == 0.0) ? 0.0 : (x < 0.0 ? -HALFPI : HALFPI);
    else
        *lam = atan2(x, y);
    return 1;
}

/* One cell of pyproj.Transformer.transform(xx, yy, direction='INVERSE')
 * (body_xy.py:3126) for the projection strings built at body_xy.py:2899-2969.
 * Restates PROJ 9's ortho (ellipsoidal, EPSG 9840), aeqd and laea (spherical)
 * inverses.  +axis=wnu (west-positive bodies) negates the projected x; the
 * longitude PROJ returns is lon_0 + lambda normalised to [-180, 180]. */
static void proj_inverse_one(int kind, double a, double b, double lon0_deg,
                             double lat0_deg, double lon_sign, double xx, double yy,
                             double *lon_out, double *lat_out) {
    *lon_out = *lat_out = kNaN;
    if (!(isfinite(xx) && isfinite(yy))) return;
    const double phi0 = lat0_deg * RPD;
    double x = (lon_sign < 0.0 ? -xx : xx);
    double y = yy;
    double lam = 0.0, phi = 0.0;
    if (kind == PM_PROJ_ORTHOGRAPHIC) {
        /* +to_meter=a, +y_0=a(b/a-1)sin(2 lat_0)  =>  normalised y = yy - y_0/a */
        y = yy - (b / a - 1.0) * sin((lat0_deg * 2.0) * RPD);
        const double es = 1.0 - (b * b) / (a * a);
        if (es == 0.0) {
            if (!ortho_sph_inverse(x, y, phi0, &phi, &lam)) return;
        } else {
            const double sinph0 = sin(phi0), cosph0 = cos(phi0);
            const double one_es = 1.0 - es;
            if (fabs(fabs(phi0) - HALFPI) < 1e-10) { /* polar */
                const double sgn = phi0 > 0.0 ? 1.0 : -1.0;
                double rh2 = x * x + y * y;
                if (rh2 >= 1.0 - 1e-15) {
                    if (rh2 - 1.0 > 1e-10) return;
                    phi = 0.0;
                } else {
                    phi = acos(sqrt(rh2 * one_es / (1.0 - es * rh2))) * sgn;
                }
                lam = atan2(x, y * -sgn);
            } else if (fabs(phi0) < 1e-10) { /* equatorial */
                double ys = y * (a / b);
                if (x * x + ys * ys > 1.0 + 1e-11) return;
                double sinphi2 = (y == 0.0) ? 0.0 : 1.0 / ((one_es / y) * (one_es / y) + es);
                if (sinphi2 > 1.0 - 1e-11) {
                    phi = HALFPI * (y > 0.0 ? 1.0 : -1.0);
                    lam = 0.0;
                } else {
                    phi = asin(sqrt(sinphi2)) * (y > 0.0 ? 1.0 : -1.0);
                    double sinlam = x * sqrt((1.0 - es * sinphi2) / (1.0 - sinphi2));
                    if (fabs(sinlam) - 1.0 > -1e-15)
                        lam = HALFPI * (x > 0.0 ? 1.0 : -1.0);
                    else
                        lam = asin(sinlam);
                }
            } else { /* oblique */
                const double nu0 = 1.0 / sqrt(1.0 - es * sinph0 * sinph0);
                const double y_shift = es * nu0 * sinph0 * cosph0;
                const double y_scale = 1.0 / sqrt(1.0 - es * cosph0 * cosph0);
                double yr = (y - y_shift) / y_scale;
                if (x * x + yr * yr > 1.0 + 1e-11) return;
                if (!ortho_sph_inverse(x, yr, phi0, &phi, &lam)) return;
                int ok = 0;
                for (int it = 0; it < 20; it++) {
                    double cp = cos(phi), sp = sin(phi), cl = cos(lam), sl = sin(lam);
                    double w = 1.0 - es * sp * sp;
                    double nu = 1.0 / sqrt(w);
                    double fx = nu * cp * sl;
                    double fy = nu * (sp * cosph0 - cp * sinph0 * cl) +
                                es * (nu0 * sinph0 - nu * sp) * cosph0;
                    double rho = one_es * nu / w;
                    double J11 = -rho * sp * sl;
                    double J12 = nu * cp * cl;
                    double J21 = rho * (cp * cosph0 + sp * sinph0 * cl);
                    double J22 = nu * sinph0 * cp * sl;
                    double D = J11 * J22 - J12 * J21;
                    double dx = x - fx, dy = y - fy;
                    double dphi = (J22 * dx - J12 * dy) / D;
                    double dlam = (-J21 * dx + J11 * dy) / D;
                    phi += dphi;
                    if (phi > HALFPI)
                        phi = HALFPI - (phi - HALFPI);
                    else if (phi < -HALFPI)
                        phi = -HALFPI + (-HALFPI - phi);
                    lam += dlam;
                    if (fabs(dphi) < 1e-12 && fabs(dlam) < 1e-12) {
                        ok = 1;
                        break;
                    }
                }
                if (!ok) return;
            }
        }
    } else if (kind == PM_PROJ_AZIMUTHAL) {
        /* +proj=aeqd, sphere of radius a, +to_meter = a pi (aeqd s_inverse) */
        x *= PI;
        y *= PI;
        double c_rh = hypot(x, y);
        int at_origin = 0;
        if (c_rh > PI) {
            if (c_rh - 1e-10 > PI) return;
            c_rh = PI;
        } else if (c_rh < 1e-10) {
            at_origin = 1;
        }
        if (at_origin) {
            phi = phi0;
            lam = 0.0;
        } else if (fabs(fabs(phi0) - HALFPI) < 1e-10) {
            if (phi0 > 0.0) {
                phi = HALFPI - c_rh;
                lam = atan2(x, -y);
            } else {
                phi = c_rh - HALFPI;
                lam = atan2(x, y);
            }
        } else {
            double sinc = sin(c_rh), cosc = cos(c_rh);
            double sinph0 = sin(phi0), cosph0 = cos(phi0);
            double arg;
            if (fabs(phi0) < 1e-10) {
                arg = y * sinc / c_rh;
                x *= sinc;
                y = cosc * c_rh;
            } else {
                arg = cosc * sinph0 + y * sinc * cosph0 / c_rh;
                arg = fmax(-1.0, fmin(1.0, arg));
                y = (cosc - sinph0 * arg) * c_rh; /* sin(asin(arg)) == arg */
                x *= sinc * cosph0;
            }
            arg = fmax(-1.0, fmin(1.0, arg));
            phi = asin(arg);
            lam = (y == 0.0) ? 0.0 : atan2(x, y);
        }
    } else if (kind == PM_PROJ_AZIMUTHAL_EQUAL_AREA) {
        /* +proj=laea, sphere of radius a, +to_meter = 2a (laea s_inverse) */
        x *= 2.0;
        y *= 2.0;
        double rh = hypot(x, y);
        double half = rh * 0.5;
        if (half > 1.0) return;
        double z = 2.0 * asin(half);
        double sinph0 = sin(phi0), cosph0 = cos(phi0);
        if (fabs(fabs(phi0) - HALFPI) < 1e-10) {
            if (phi0 > 0.0) {
                y = -y;
                phi = HALFPI - z;
            } else {
                phi = z - HALFPI;
            }
            lam = atan2(x, y);
        } else {
            double sinz = sin(z), cosz = cos(z);
            if (fabs(phi0) < 1e-10) {
                phi = (fabs(rh) <= 1e-10) ? 0.0 : asin(y * sinz / rh);
                x *= sinz;
                y = cosz * rh;
            } else {
                phi = (fabs(rh) <= 1e-10) ? phi0
                                          : asin(cosz * sinph0 + y * sinz * cosph0 / rh);
                x *= sinz * cosph0;
                y = (cosz - sin(phi) * sinph0) * rh;
            }
            lam = (y == 0.0) ? 0.0 : atan2(x, y);
        }
    } else {
        return;
    }
    /* lon = lon_0 + lambda, wrapped to [-180, 180] like PROJ's adjlon */
    double lon_rad = lam + lon0_deg * RPD;
    if (fabs(lon_rad) > PI + 1e-12) {
        lon_rad += PI;
        lon_rad -= TWOPI * floor(lon_rad / TWOPI);
        lon_rad -= PI;
    }
    *lon_out = lon_rad * DPR;
    *lat_out = phi * DPR;
}

/* Forward direction of the same projections (pyproj's default direction; PROJ's published forward
 * formulas: ellipsoidal ortho e_forward, spherical aeqd / laea s_forward), in the units of the reference's
 * own strings (body_xy.py:2930-2968: to_meter = a, a pi, 2 a; the ortho's y_0). */
static void proj_forward_one(int kind, double a, double b, double lon0_deg, double lat0_deg, double lon_sign,
                             double lon_deg, double lat_deg, double *xx, double *yy) {
    *xx = *yy = kNaN;
    if (!(isfinite(lon_deg) && isfinite(lat_deg)) || fabs(lat_deg) > 90.0 + 1e-12) return;
    double phi0 = lat0_deg * RPD, phi = fmax(-HALFPI, fmin(HALFPI, lat_deg * RPD));
    double lam = (lon_deg - lon0_deg) * RPD;
    lam -= TWOPI * floor((lam + PI) / TWOPI);
    double sinph0 = sin(phi0), cosph0 = cos(phi0), sinphi = sin(phi), cosphi = cos(phi);
    double sinlam = sin(lam), coslam = cos(lam);
    double cosc = sinph0 * sinphi + cosph0 * cosphi * coslam;
    double x, y;
    if (kind == PM_PROJ_ORTHOGRAPHIC) {
        if (cosc < -1e-10) return;
        double es = 1.0 - (b * b) / (a * a);
        double nu = 1.0 / sqrt(1.0 - es * sinphi * sinphi), nu0 = 1.0 / sqrt(1.0 - es * sinph0 * sinph0);
        x = nu * cosphi * sinlam;
        y = nu * (sinphi * cosph0 - cosphi * sinph0 * coslam) + es * (nu0 * sinph0 - nu * sinphi) * cosph0;
        y += (b / a - 1.0) * sin((lat0_deg * 2.0) * RPD);
    } else if (kind == PM_PROJ_AZIMUTHAL) {
        if (cosc <= -1.0 + 1e-14) return;
        double c = acos(fmax(-1.0, fmin(1.0, cosc)));
        double k = c < 1e-10 ? 1.0 : c / sin(c);
        x = k * cosphi * sinlam / PI;
        y = k * (cosph0 * sinphi - sinph0 * cosphi * coslam) / PI;
    } else if (kind == PM_PROJ_AZIMUTHAL_EQUAL_AREA) {
        double d = 1.0 + cosc;
        if (d <= 1e-10) return;
        double k = sqrt(2.0 / d);
        x = k * cosphi * sinlam * 0.5;
        y = k * (cosph0 * sinphi - sinph0 * cosphi * coslam) * 0.5;
    } else {
        return;
    }
    *xx = lon_sign < 0.0 ? -x : x;
    *yy = y;
}

int pmo_proj_forward(int kind, const double *params5, const double *lon, const double *lat, int64_t n,
                     double *xx, double *yy) {
    if (!params5 || !lon || !lat || !xx || !yy || n < 0) return PM_ERR_BAD_ARG;
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++)
        proj_forward_one(kind, params5[0], params5[1], params5[2], params5[3], params5[4], lon[i], lat[i], &xx[i],
                         &yy[i]);
    return PM_OK;
}

int pmo_proj_inverse(int kind, const double *params5, const double *xx, const double *yy,
                     int64_t n, double *lon, double *lat) {
    if (!params5 || !xx || !yy || !lon || !lat || n < 0) return PM_ERR_BAD_ARG;
    if (kind < PM_PROJ_ORTHOGRAPHIC || kind > PM_PROJ_AZIMUTHAL_EQUAL_AREA)
        return PM_ERR_UNSUPPORTED;
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++) {
        proj_inverse_one(kind, params5[0], params5[1], params5[2], params5[3], params5[4],
                         xx[i], yy[i], &lon[i], &lat[i]);
    }
    return PM_OK;
}

/* ---------- gather ---------- */
/* BodyXY._do_nearest_interpolation (body_xy.py:1633-1649) over a cube
 * (Observation._get_mapped_data, observation.py:892-905) */
int pmo_gather_nearest(const double *cube, int n_planes, int ny, int nx,
                       const double *xmap, const double *ymap, int64_t n_cells,
                       double *out) {
    if (!cube || !xmap || !ymap || !out) return PM_ERR_BAD_ARG;
#pragma omp parallel for
    for (int64_t i = 0; i < n_cells; i++) {
        double x = xmap[i], y = ymap[i];
        long xi = -999, yi = -999;
        if (!isnan(x)) xi = (long)nearbyint(x);
        if (!isnan(y)) yi = (long)nearbyint(y);
        for (int l = 0; l < n_planes; l++) {
            double v = kNaN;
            if (xi != -999) {
                /* numpy negative indices wrap; cannot occur for x_map/y_map built by
                 * _get_xy_map, handled for manual maps */
                long xx = xi < 0 ? xi + nx : xi, yy2 = yi < 0 ? yi + ny : yi;
                if (xx >= 0 && xx < nx && yy2 >= 0 && yy2 < ny)
                    v = cube[((int64_t)l * ny + yy2) * nx + xx];
            }
            out[(int64_t)l * n_cells + i] = v;
        }
    }
    return PM_OK;
}
