/*
 * Host-side orbit model for the Doppler schedule: NORAD SGP4 (Spacetrack Report #3, near-earth branch), TLE parsing,
 * Julian dates and the observer's range rate. Evaluated twice per channel per second of signal (reference
 * src/dsp/doppler.c:150-172), so it stays on the CPU in double precision.
 *
 * It stands in for what the reference reaches through src/sgpsdp (Get_Next_Tle_Set, select_ephemeris, SGP4,
 * Convert_Sat_State, Calculate_Obs, Julian_Date*, reference src/dsp/doppler.c:31-42,102-110). Constants (WGS-72, the
 * truncated pi of the original Pascal units) and the order of the floating point operations follow that code, because
 * the Doppler frequency is truncated to an integer number of Hz before it reaches the NCO and must land on the same
 * integer. Deep-space satellites (period >= 225 min, SDP4) are not supported yet: create fails for them.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "orbit.h"

/* constants of the reference's sgp4sdp4.h (WGS-72, values as written there) */
#define K_DE2RA 1.74532925E-2
#define K_PI 3.1415926535898
#define K_PIO2 1.5707963267949
#define K_X3PIO2 4.71238898
#define K_TWOPI 6.2831853071796
#define K_E6A 1.0E-6
#define K_TOTHRD 6.6666667E-1
#define K_XJ3 -2.53881E-6
#define K_XKE 7.43669161E-2
#define K_XKMPER 6.378135E3
#define K_XMNPDA 1.44E3
#define K_AE 1.0
#define K_CK2 5.413079E-4
#define K_CK4 6.209887E-7
#define K_F 3.352779E-3
#define K_S 1.012229
#define K_QOMS2T 1.880279E-09
#define K_SECDAY 8.6400E4
#define K_OMEGA_E 1.0027379
#define K_MFACTOR 7.292115E-5

static double sqr(double x) { return x * x; }

static double frac(double x) { return x - floor(x); }

static double mod_2pi(double x) {
    double r = x;
    const int i = (int) (r / K_TWOPI);
    r -= i * K_TWOPI;
    if (r < 0) {
        r += K_TWOPI;
    }
    return r;
}

static double modulus(double a, double b) {
    double r = a;
    const int i = (int) (r / b);
    r -= i * b;
    if (r < 0) {
        r += b;
    }
    return r;
}

static double arctan4(double sinx, double cosx) {
    if (cosx == 0) {
        return sinx > 0 ? K_PIO2 : K_X3PIO2;
    }
    if (cosx > 0) {
        return sinx > 0 ? atan(sinx / cosx) : K_TWOPI + atan(sinx / cosx);
    }
    return K_PI + atan(sinx / cosx);
}

/* ---- dates ------------------------------------------------------------------------------------------------------------ */

static double julian_date_of_year(double year) {
    year = year - 1;
    long i = (long) (year / 100);
    const long a = i;
    i = a / 4;
    const long b = 2 - a + i;
    i = (long) (365.25 * year);
    i += (long) (30.6001 * 14);
    return (double) i + 1720994.5 + (double) b;
}

static int day_of_year(int yr, int mo, int dy) {
    static const int days[] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    int day = 0;
    for (int i = 0; i < mo - 1; i++) {
        day += days[i];
    }
    day += dy;
    if ((yr % 4 == 0) && ((yr % 100 != 0) || (yr % 400 == 0)) && (mo > 2)) {
        day++;
    }
    return day;
}

double sdrm_julian_date(int year, int month, int mday, int hour, int min, int sec) {
    return julian_date_of_year(year) + day_of_year(year, month, mday) + ((hour + (min + sec / 60.0) / 60.0) / 24.0);
}

double sdrm_julian_date_of_epoch(double epoch) {
    double year;
    const double day = modf(epoch * 1E-3, &year) * 1E3;
    year = year < 57 ? year + 2000 : year + 1900;
    return julian_date_of_year(year) + day;
}

static double theta_g_jd(double jd) {
    const double ut = frac(jd + 0.5);
    jd = jd - ut;
    const double tu = (jd - 2451545.0) / 36525;
    double gmst = 24110.54841 + tu * (8640184.812866 + tu * (0.093104 - tu * 6.2E-6));
    gmst = modulus(gmst + K_SECDAY * K_OMEGA_E * ut, K_SECDAY);
    return K_TWOPI * gmst / K_SECDAY;
}

/* ---- two-line elements ---------------------------------------------------------------------------------------------------- */

static int checksum_ok(const char *line) {
    int sum = 0;
    for (int i = 0; i < 68; i++) {
        if (line[i] >= '0' && line[i] <= '9') {
            sum += line[i] - '0';
        } else if (line[i] == '-') {
            sum += 1;
        }
    }
    return (sum % 10) == (line[68] - '0');
}

static double field(const char *set, int start, int len) {
    char buf[24];
    memcpy(buf, set + start, (size_t) len);
    buf[len] = '\0';
    return atof(buf);
}

/* "SMMMMM-E" style fields: one sign/space, five mantissa digits (implied leading "."), exponent */
static double exp_field(const char *set, int start) {
    char buf[16];
    buf[0] = set[start];
    buf[1] = '.';
    memcpy(buf + 2, set + start + 1, 5);
    buf[7] = 'E';
    memcpy(buf + 8, set + start + 6, 2);
    buf[10] = '\0';
    return atof(buf);
}

int sdrm_orbit_init(const char tle[3][80], sdrm_orbit *orbit) {
    char set[140];
    memset(orbit, 0, sizeof(*orbit));
    memset(set, 0, sizeof(set));
    strncpy(set, tle[1], 70);
    strncpy(set + 69, tle[2], 70);
    set[138] = '\0';
    /* validity checks of the element set: checksums, line numbers, matching catalogue numbers, decimal points */
    if (strlen(tle[1]) < 69 || strlen(tle[2]) < 69 || !checksum_ok(set) || !checksum_ok(set + 69) || set[0] != '1' || set[69] != '2' ||
        strncmp(set + 2, set + 71, 5) != 0 || set[23] != '.' || set[34] != '.' || set[80] != '.' || set[89] != '.' || set[106] != '.' ||
        set[115] != '.' || set[123] != '.' || strncmp(set + 61, " 0 ", 3) != 0) {
        return -1;
    }
    char epoch_text[16];
    memcpy(epoch_text, set + 18, 14);
    epoch_text[14] = '\0';
    if (epoch_text[2] == ' ') epoch_text[2] = '0';
    if (epoch_text[3] == ' ') epoch_text[3] = '0';
    orbit->epoch = atof(epoch_text);
    orbit->bstar = exp_field(set, 53);
    orbit->xincl = field(set, 77, 8);
    orbit->xnodeo = field(set, 86, 8);
    {
        char buf[12];
        buf[0] = '.';
        memcpy(buf + 1, set + 95, 7);
        buf[8] = '\0';
        orbit->eo = atof(buf);
    }
    if (orbit->eo < 1.0e-6) {
        orbit->eo = 1.0e-6;
    }
    orbit->omegao = field(set, 103, 8);
    orbit->xmo = field(set, 112, 8);
    orbit->xno = field(set, 121, 10);

    /* units: degrees -> radians, rev/day -> rad/min */
    orbit->xnodeo *= K_DE2RA;
    orbit->omegao *= K_DE2RA;
    orbit->xmo *= K_DE2RA;
    orbit->xincl *= K_DE2RA;
    const double temp = K_TWOPI / K_XMNPDA / K_XMNPDA;
    orbit->xno = orbit->xno * temp * K_XMNPDA;
    orbit->bstar /= K_AE;

    /* period >= 225 minutes selects the deep-space model, which is not implemented here */
    double dd1 = K_XKE / orbit->xno;
    const double a1 = pow(dd1, K_TOTHRD);
    const double r1 = cos(orbit->xincl);
    dd1 = 1.0 - orbit->eo * orbit->eo;
    const double t = K_CK2 * 1.5f * (r1 * r1 * 3.0 - 1.0) / pow(dd1, 1.5);
    const double del1 = t / (a1 * a1);
    const double ao = a1 * (1.0 - del1 * (K_TOTHRD * 0.5 + del1 * (del1 * 1.654320987654321 + 1.0)));
    const double delo = t / (ao * ao);
    const double xnodp = orbit->xno / (delo + 1.0);
    if (K_TWOPI / xnodp / K_XMNPDA >= .15625) {
        return -2;
    }
    orbit->jul_epoch = sdrm_julian_date_of_epoch(orbit->epoch);
    return 0;
}

/* ---- SGP4 ---------------------------------------------------------------------------------------------------------------- */

static void sgp4_setup(sdrm_orbit *o) {
    o->ready = 1;
    /* recover the original mean motion and semi-major axis */
    const double a1 = pow(K_XKE / o->xno, K_TOTHRD);
    o->cosio = cos(o->xincl);
    const double theta2 = o->cosio * o->cosio;
    o->x3thm1 = 3 * theta2 - 1.0;
    const double eosq = o->eo * o->eo;
    const double betao2 = 1 - eosq;
    const double betao = sqrt(betao2);
    const double del1 = 1.5 * K_CK2 * o->x3thm1 / (a1 * a1 * betao * betao2);
    const double ao = a1 * (1 - del1 * (0.5 * K_TOTHRD + del1 * (1 + 134.0 / 81.0 * del1)));
    const double delo = 1.5 * K_CK2 * o->x3thm1 / (ao * ao * betao * betao2);
    o->xnodp = o->xno / (1.0 + delo);
    o->aodp = ao / (1.0 - delo);
    /* perigee below 220 km: truncated ("simple") equations */
    o->simple = (o->aodp * (1.0 - o->eo) / K_AE) < (220.0 / K_XKMPER + K_AE);
    /* perigee below 156 km alters s and qoms2t */
    double s4 = K_S;
    double qoms24 = K_QOMS2T;
    const double perige = (o->aodp * (1 - o->eo) - K_AE) * K_XKMPER;
    if (perige < 156.0) {
        s4 = perige <= 98.0 ? 20.0 : perige - 78.0;
        qoms24 = pow((120.0 - s4) * K_AE / K_XKMPER, 4);
        s4 = s4 / K_XKMPER + K_AE;
    }
    const double pinvsq = 1.0 / (o->aodp * o->aodp * betao2 * betao2);
    const double tsi = 1.0 / (o->aodp - s4);
    o->eta = o->aodp * o->eo * tsi;
    const double etasq = o->eta * o->eta;
    const double eeta = o->eo * o->eta;
    const double psisq = fabs(1.0 - etasq);
    const double coef = qoms24 * pow(tsi, 4);
    const double coef1 = coef / pow(psisq, 3.5);
    const double c2 = coef1 * o->xnodp *
                      (o->aodp * (1.0 + 1.5 * etasq + eeta * (4.0 + etasq)) +
                       0.75 * K_CK2 * tsi / psisq * o->x3thm1 * (8.0 + 3.0 * etasq * (8 + etasq)));
    o->c1 = c2 * o->bstar;
    o->sinio = sin(o->xincl);
    const double a3ovk2 = -K_XJ3 / K_CK2 * pow(K_AE, 3);
    const double c3 = coef * tsi * a3ovk2 * o->xnodp * K_AE * o->sinio / o->eo;
    o->x1mth2 = 1.0 - theta2;
    o->c4 = 2.0 * o->xnodp * coef1 * o->aodp * betao2 *
            (o->eta * (2.0 + 0.5 * etasq) + o->eo * (0.5 + 2.0 * etasq) -
             2.0 * K_CK2 * tsi / (o->aodp * psisq) *
                 (-3.0 * o->x3thm1 * (1.0 - 2.0 * eeta + etasq * (1.5 - 0.5 * eeta)) +
                  0.75 * o->x1mth2 * (2.0 * etasq - eeta * (1.0 + etasq)) * cos(2.0 * o->omegao)));
    o->c5 = 2.0 * coef1 * o->aodp * betao2 * (1.0 + 2.75 * (etasq + eeta) + eeta * etasq);
    const double theta4 = theta2 * theta2;
    const double temp1 = 3.0 * K_CK2 * pinvsq * o->xnodp;
    const double temp2 = temp1 * K_CK2 * pinvsq;
    const double temp3 = 1.25 * K_CK4 * pinvsq * pinvsq * o->xnodp;
    o->xmdot = o->xnodp + 0.5 * temp1 * betao * o->x3thm1 + 0.0625 * temp2 * betao * (13.0 - 78.0 * theta2 + 137.0 * theta4);
    const double x1m5th = 1.0 - 5.0 * theta2;
    o->omgdot = -0.5 * temp1 * x1m5th + 0.0625 * temp2 * (7.0 - 114.0 * theta2 + 395.0 * theta4) +
                temp3 * (3.0 - 36.0 * theta2 + 49.0 * theta4);
    const double xhdot1 = -temp1 * o->cosio;
    o->xnodot = xhdot1 + (0.5 * temp2 * (4.0 - 19.0 * theta2) + 2.0 * temp3 * (3.0 - 7.0 * theta2)) * o->cosio;
    o->omgcof = o->bstar * c3 * cos(o->omegao);
    o->xmcof = -K_TOTHRD * coef * o->bstar * K_AE / eeta;
    o->xnodcf = 3.5 * betao2 * xhdot1 * o->c1;
    o->t2cof = 1.5 * o->c1;
    o->xlcof = 0.125 * a3ovk2 * o->sinio * (3.0 + 5.0 * o->cosio) / (1.0 + o->cosio);
    o->aycof = 0.25 * a3ovk2 * o->sinio;
    o->delmo = pow(1.0 + o->eta * cos(o->xmo), 3);
    o->sinmo = sin(o->xmo);
    o->x7thm1 = 7.0 * theta2 - 1.0;
    if (!o->simple) {
        const double c1sq = o->c1 * o->c1;
        o->d2 = 4.0 * o->aodp * tsi * c1sq;
        const double temp = o->d2 * tsi * o->c1 / 3.0;
        o->d3 = (17.0 * o->aodp + s4) * temp;
        o->d4 = 0.5 * temp * o->aodp * tsi * (221.0 * o->aodp + 31.0 * s4) * o->c1;
        o->t3cof = o->d2 + 2.0 * c1sq;
        o->t4cof = 0.25 * (3.0 * o->d3 + o->c1 * (12.0 * o->d2 + 10.0 * c1sq));
        o->t5cof = 0.2 * (3.0 * o->d4 + 12.0 * o->c1 * o->d3 + 6.0 * o->d2 * o->d2 + 15.0 * c1sq * (2.0 * o->d2 + c1sq));
    }
}

/* Position (earth radii) and velocity (earth radii / min) in the ECI frame, tsince minutes after the epoch. */
static void sgp4_propagate(sdrm_orbit *o, double tsince, double pos[3], double vel[3]) {
    if (!o->ready) {
        sgp4_setup(o);
    }
    /* secular gravity and atmospheric drag */
    const double xmdf = o->xmo + o->xmdot * tsince;
    const double omgadf = o->omegao + o->omgdot * tsince;
    const double xnoddf = o->xnodeo + o->xnodot * tsince;
    double omega = omgadf;
    double xmp = xmdf;
    const double tsq = tsince * tsince;
    const double xnode = xnoddf + o->xnodcf * tsq;
    double tempa = 1.0 - o->c1 * tsince;
    double tempe = o->bstar * o->c4 * tsince;
    double templ = o->t2cof * tsq;
    if (!o->simple) {
        const double delomg = o->omgcof * tsince;
        const double delm = o->xmcof * (pow(1 + o->eta * cos(xmdf), 3) - o->delmo);
        const double temp = delomg + delm;
        xmp = xmdf + temp;
        omega = omgadf - temp;
        const double tcube = tsq * tsince;
        const double tfour = tsince * tcube;
        tempa = tempa - o->d2 * tsq - o->d3 * tcube - o->d4 * tfour;
        tempe = tempe + o->bstar * o->c5 * (sin(xmp) - o->sinmo);
        templ = templ + o->t3cof * tcube + tfour * (o->t4cof + tsince * o->t5cof);
    }
    const double a = o->aodp * pow(tempa, 2);
    const double e = o->eo - tempe;
    const double xl = xmp + omega + xnode + o->xnodp * templ;
    const double beta = sqrt(1.0 - e * e);
    const double xn = K_XKE / pow(a, 1.5);

    /* long period periodics */
    const double axn = e * cos(omega);
    double temp = 1.0 / (a * beta * beta);
    const double xll = temp * o->xlcof * axn;
    const double aynl = temp * o->aycof;
    const double xlt = xl + xll;
    const double ayn = e * sin(omega) + aynl;

    /* Kepler's equation */
    const double capu = mod_2pi(xlt - xnode);
    double temp2 = capu;
    double sinepw;
    double cosepw;
    double temp3;
    double temp4;
    double temp5;
    double temp6;
    int i = 0;
    do {
        sinepw = sin(temp2);
        cosepw = cos(temp2);
        temp3 = axn * sinepw;
        temp4 = ayn * cosepw;
        temp5 = axn * cosepw;
        temp6 = ayn * sinepw;
        const double epw = (capu - temp4 + temp3 - temp2) / (1.0 - temp5 - temp6) + temp2;
        if (fabs(epw - temp2) <= K_E6A) {
            break;
        }
        temp2 = epw;
    } while (i++ < 10);

    /* short period preliminary quantities */
    const double ecose = temp5 + temp6;
    const double esine = temp3 - temp4;
    const double elsq = axn * axn + ayn * ayn;
    temp = 1.0 - elsq;
    const double pl = a * temp;
    const double r = a * (1.0 - ecose);
    double temp1 = 1.0 / r;
    const double rdot = K_XKE * sqrt(a) * esine * temp1;
    const double rfdot = K_XKE * sqrt(pl) * temp1;
    temp2 = a * temp1;
    const double betal = sqrt(temp);
    temp3 = 1.0 / (1.0 + betal);
    const double cosu = temp2 * (cosepw - axn + ayn * esine * temp3);
    const double sinu = temp2 * (sinepw - ayn - axn * esine * temp3);
    const double u = arctan4(sinu, cosu);
    const double sin2u = 2.0 * sinu * cosu;
    const double cos2u = 2.0 * cosu * cosu - 1.0;
    temp = 1.0 / pl;
    temp1 = K_CK2 * temp;
    temp2 = temp1 * temp;

    /* short periodics */
    const double rk = r * (1.0 - 1.5 * temp2 * betal * o->x3thm1) + 0.5 * temp1 * o->x1mth2 * cos2u;
    const double uk = u - 0.25 * temp2 * o->x7thm1 * sin2u;
    const double xnodek = xnode + 1.5 * temp2 * o->cosio * sin2u;
    const double xinck = o->xincl + 1.5 * temp2 * o->cosio * o->sinio * cos2u;
    const double rdotk = rdot - xn * temp1 * o->x1mth2 * sin2u;
    const double rfdotk = rfdot + xn * temp1 * (o->x1mth2 * cos2u + 1.5 * o->x3thm1);

    /* orientation vectors */
    const double sinuk = sin(uk);
    const double cosuk = cos(uk);
    const double sinik = sin(xinck);
    const double cosik = cos(xinck);
    const double sinnok = sin(xnodek);
    const double cosnok = cos(xnodek);
    const double xmx = -sinnok * cosik;
    const double xmy = cosnok * cosik;
    const double ux = xmx * sinuk + cosnok * cosuk;
    const double uy = xmy * sinuk + sinnok * cosuk;
    const double uz = sinik * sinuk;
    const double vx = xmx * cosuk - cosnok * sinuk;
    const double vy = xmy * cosuk - sinnok * sinuk;
    const double vz = sinik * cosuk;

    pos[0] = rk * ux;
    pos[1] = rk * uy;
    pos[2] = rk * uz;
    vel[0] = rdotk * ux + rfdotk * vx;
    vel[1] = rdotk * uy + rfdotk * vy;
    vel[2] = rdotk * uz + rfdotk * vz;
}

/* ---- observer ---------------------------------------------------------------------------------------------------------------- */

double sdrm_orbit_range_rate(sdrm_orbit *orbit, double jul_utc, double lat_rad, double lon_rad, double alt_km) {
    const double tsince = (jul_utc - orbit->jul_epoch) * K_XMNPDA;
    double pos[3];
    double vel[3];
    sgp4_propagate(orbit, tsince, pos, vel);
    /* to km and km/s */
    const double kv = K_XKMPER * K_XMNPDA / K_SECDAY;
    for (int k = 0; k < 3; k++) {
        pos[k] *= K_XKMPER;
        vel[k] *= kv;
    }
    /* observer position and velocity in the ECI frame (1992 Astronomical Almanac, K11) */
    const double theta = mod_2pi(theta_g_jd(jul_utc) + lon_rad);
    const double c = 1 / sqrt(1 + K_F * (K_F - 2) * sqr(sin(lat_rad)));
    const double sq = sqr(1 - K_F) * c;
    const double achcp = (K_XKMPER * c + alt_km) * cos(lat_rad);
    const double ox = achcp * cos(theta);
    const double oy = achcp * sin(theta);
    const double oz = (K_XKMPER * sq + alt_km) * sin(lat_rad);
    const double ovx = -K_MFACTOR * oy;
    const double ovy = K_MFACTOR * ox;
    const double rx = pos[0] - ox;
    const double ry = pos[1] - oy;
    const double rz = pos[2] - oz;
    const double rvx = vel[0] - ovx;
    const double rvy = vel[1] - ovy;
    const double rvz = vel[2] - 0;
    const double range = sqrt(sqr(rx) + sqr(ry) + sqr(rz));
    return (rx * rvx + ry * rvy + rz * rvz) / range;
}
