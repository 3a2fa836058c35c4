/*
 * Host-side orbit model for the Doppler schedule: NORAD SGP4 and SDP4 (Spacetrack Report #3: near-earth and deep-space
 * branches, the latter with lunar/solar perturbations and the 12 h / 24 h resonance integrator), TLE parsing, Julian dates
 * and the observer's range rate. Evaluated twice per channel per second of signal (reference
 * src/dsp/doppler.c:150-172), so it stays on the CPU in double precision.
 *
 * It stands in for what the reference reaches through src/sgpsdp (Get_Next_Tle_Set, select_ephemeris, SGP4,
 * Convert_Sat_State, Calculate_Obs, Julian_Date*, reference src/dsp/doppler.c:31-42,102-110). Constants (WGS-72, the
 * truncated pi of the original Pascal units) and the order of the floating point operations follow that code, because
 * the Doppler frequency is truncated to an integer number of Hz before it reaches the NCO and must land on the same
 * integer. Known answers: the reference's test/test_sgp4_001.c (SGP4) and test/test_sgp4_002.c (SDP4) in
 * tests/test_orbit.py.
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
/* deep-space constants (sgp4sdp4.h:223-248) */
#define K_ZNS 1.19459E-5
#define K_C1SS 2.9864797E-6
#define K_ZES 1.675E-2
#define K_ZNL 1.5835218E-4
#define K_C1L 4.7968065E-7
#define K_ZEL 5.490E-2
#define K_ZCOSIS 9.1744867E-1
#define K_ZSINIS 3.9785416E-1
#define K_ZSINGS -9.8088458E-1
#define K_ZCOSGS 1.945905E-1
#define K_Q22 1.7891679E-6
#define K_Q31 2.1460748E-6
#define K_Q33 2.2123015E-7
#define K_G22 5.7686396
#define K_G32 9.5240898E-1
#define K_G44 1.8014998
#define K_G52 1.0508330
#define K_G54 4.4108898
#define K_ROOT22 1.7891679E-6
#define K_ROOT32 3.7393792E-7
#define K_ROOT44 7.3636953E-9
#define K_ROOT52 1.1428639E-7
#define K_ROOT54 2.1765803E-9
#define K_THDT 4.3752691E-3
#define K_RESONANCE_STEP 720.0     /* minutes */
#define K_RESONANCE_STEP2 259200.0 /* step^2 / 2 */

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

/* The numeric columns must hold digits, blanks, signs and points only: atof would also read "inf", "nan", hexadecimal and
 * exponent forms, and an element set from the network (RxRequest.doppler.tle) with such a field — its checksum may still
 * add up, letters count as zero — would reach the date arithmetic as a non-finite year. The reference's Good_Elements
 * (src/sgpsdp/sgp_in.c) does not look; such a set is no element set, so it is rejected here. */
static int numeric_ok(const char *set, int start, int len) {
    for (int i = start; i < start + len; i++) {
        const char c = set[i];
        if (!((c >= '0' && c <= '9') || c == ' ' || c == '.' || c == '+' || c == '-')) {
            return 0;
        }
    }
    return 1;
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
    if (!numeric_ok(set, 18, 14) || !numeric_ok(set, 53, 8) || !numeric_ok(set, 77, 8) || !numeric_ok(set, 86, 8) ||
        !numeric_ok(set, 95, 7) || !numeric_ok(set, 103, 8) || !numeric_ok(set, 112, 8) || !numeric_ok(set, 121, 10)) {
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

    /* period >= 225 minutes selects the deep-space model (select_ephemeris, sgp_in.c) */
    double dd1 = K_XKE / orbit->xno;
    const double a1 = pow(dd1, K_TOTHRD);
    const double r1 = cos(orbit->xincl);
    dd1 = 1.0 - orbit->eo * orbit->eo;
    const double t = K_CK2 * 1.5f * (r1 * r1 * 3.0 - 1.0) / pow(dd1, 1.5);
    const double del1 = t / (a1 * a1);
    const double ao = a1 * (1.0 - del1 * (K_TOTHRD * 0.5 + del1 * (del1 * 1.654320987654321 + 1.0)));
    const double delo = t / (ao * ao);
    const double xnodp = orbit->xno / (delo + 1.0);
    orbit->deep = K_TWOPI / xnodp / K_XMNPDA >= .15625;
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

/* Long-period periodics, Kepler's equation, short-period periodics and the orientation vectors: the common tail of SGP4
 * and SDP4 (sgp4sdp4.c:189-270 and :409-499). Position in earth radii, velocity in earth radii / min, ECI. */
static void elements_to_state(const sdrm_orbit *o, double a, double e, double omega, double xnode, double xinc, double xl,
                              double pos[3], double vel[3]) {
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
    const double xinck = xinc + 1.5 * temp2 * o->cosio * o->sinio * cos2u;
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

/* SGP4 (sgp4sdp4.c:22-276): secular gravity and drag, then the common tail */
static void sgp4_propagate(sdrm_orbit *o, double tsince, double pos[3], double vel[3]) {
    if (!o->ready) {
        sgp4_setup(o);
    }
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
    elements_to_state(o, a, e, omega, xnode, o->xincl, xl, pos, vel);
}

/* ---- SDP4 ------------------------------------------------------------------------------------------------------------------ */

/* Greenwich sidereal angle at a TLE epoch, the Spacetrack Report #3 expression the deep-space model is tuned to
 * (sgp_time.c:253-278); also gives days since 1950 Jan 0.0 */
static double theta_g_epoch(double epoch, double *ds50) {
    double year;
    double day = modf(epoch * 1E-3, &year) * 1E3;
    year = year < 57 ? year + 2000 : year + 1900;
    const double ut = modf(day, &day);
    const double jd = julian_date_of_year(year) + day;
    *ds50 = jd - 2433281.5 + ut;
    return mod_2pi(6.3003880987 * *ds50 + 1.72944494);
}

/* what one perturbing body adds: secular rates (se, si, sl, sgh, sh) and its periodic coefficients */
typedef struct {
    double se, si, sl, sgh, sh;
} body_rates;

/* The body is given by the orientation of its orbit (zcosg/zsing: argument of perigee, zcosi/zsini: inclination,
 * zcosh/zsinh: node relative to the satellite's), its perturbation coefficient cc, mean motion zn and eccentricity ze
 * (Deep(), sgp4sdp4.c:585-648). */
static body_rates third_body(const sdrm_orbit *o, double zcosg, double zsing, double zcosi, double zsini, double zcosh,
                             double zsinh, double cc, double zn, double ze, sdrm_body_terms *terms) {
    const sdrm_deep_space *d = &o->ds;
    const double a1 = zcosg * zcosh + zsing * zcosi * zsinh;
    const double a3 = -zsing * zcosh + zcosg * zcosi * zsinh;
    const double a7 = -zcosg * zsinh + zsing * zcosi * zcosh;
    const double a8 = zsing * zsini;
    const double a9 = zsing * zsinh + zcosg * zcosi * zcosh;
    const double a10 = zcosg * zsini;
    const double a2 = o->cosio * a7 + o->sinio * a8;
    const double a4 = o->cosio * a9 + o->sinio * a10;
    const double a5 = -o->sinio * a7 + o->cosio * a8;
    const double a6 = -o->sinio * a9 + o->cosio * a10;
    const double x1 = a1 * d->cosg + a2 * d->sing;
    const double x2 = a3 * d->cosg + a4 * d->sing;
    const double x3 = -a1 * d->sing + a2 * d->cosg;
    const double x4 = -a3 * d->sing + a4 * d->cosg;
    const double x5 = a5 * d->sing;
    const double x6 = a6 * d->sing;
    const double x7 = a5 * d->cosg;
    const double x8 = a6 * d->cosg;
    const double z31 = 12 * x1 * x1 - 3 * x3 * x3;
    const double z32 = 24 * x1 * x2 - 6 * x3 * x4;
    const double z33 = 12 * x2 * x2 - 3 * x4 * x4;
    double z1 = 3 * (a1 * a1 + a2 * a2) + z31 * d->eosq;
    double z2 = 6 * (a1 * a3 + a2 * a4) + z32 * d->eosq;
    double z3 = 3 * (a3 * a3 + a4 * a4) + z33 * d->eosq;
    const double z11 = -6 * a1 * a5 + d->eosq * (-24 * x1 * x7 - 6 * x3 * x5);
    const double z12 = -6 * (a1 * a6 + a3 * a5) + d->eosq * (-24 * (x2 * x7 + x1 * x8) - 6 * (x3 * x6 + x4 * x5));
    const double z13 = -6 * a3 * a6 + d->eosq * (-24 * x2 * x8 - 6 * x4 * x6);
    const double z21 = 6 * a2 * a5 + d->eosq * (24 * x1 * x5 - 6 * x3 * x7);
    const double z22 = 6 * (a4 * a5 + a2 * a6) + d->eosq * (24 * (x2 * x5 + x1 * x6) - 6 * (x4 * x7 + x3 * x8));
    const double z23 = 6 * a4 * a6 + d->eosq * (24 * x2 * x6 - 6 * x4 * x8);
    z1 = z1 + z1 + d->betao2 * z31;
    z2 = z2 + z2 + d->betao2 * z32;
    z3 = z3 + z3 + d->betao2 * z33;
    const double s3 = cc * (1.0 / o->xnodp);
    const double s2 = -0.5 * s3 / d->betao;
    const double s4 = s3 * d->betao;
    const double s1 = -15 * o->eo * s4;
    const double s5 = x1 * x3 + x2 * x4;
    const double s6 = x2 * x3 + x1 * x4;
    const double s7 = x2 * x4 - x1 * x3;
    body_rates r;
    r.se = s1 * zn * s5;
    r.si = s2 * zn * (z11 + z13);
    r.sl = -zn * s3 * (z1 + z3 - 14 - 6 * d->eosq);
    r.sgh = s4 * zn * (z31 + z33 - 6);
    r.sh = -zn * s2 * (z21 + z23);
    if (o->xincl < 5.2359877E-2) {
        r.sh = 0;
    }
    terms->e2 = 2 * s1 * s6;
    terms->e3 = 2 * s1 * s7;
    terms->i2 = 2 * s2 * z12;
    terms->i3 = 2 * s2 * (z13 - z11);
    terms->l2 = -2 * s3 * z2;
    terms->l3 = -2 * s3 * (z3 - z1);
    terms->l4 = -2 * s3 * (-21 - 9 * d->eosq) * ze;
    terms->gh2 = 2 * s4 * z32;
    terms->gh3 = 2 * s4 * (z33 - z31);
    terms->gh4 = -18 * s4 * ze;
    terms->h2 = -2 * s2 * z22;
    terms->h3 = -2 * s2 * (z23 - z21);
    return r;
}

/* Deep-space initialisation (Deep() entry dpinit, sgp4sdp4.c:530-826) */
static void deep_setup(sdrm_orbit *o) {
    sdrm_deep_space *d = &o->ds;
    d->theta2 = o->cosio * o->cosio;
    d->eosq = o->eo * o->eo;
    d->betao2 = 1.0 - d->eosq;
    d->betao = sqrt(d->betao2);
    d->sing = sin(o->omegao);
    d->cosg = cos(o->omegao);
    double ds50;
    d->thgr = theta_g_epoch(o->epoch, &ds50);
    const double eq = o->eo;
    const double xnq = o->xnodp;
    const double aqnv = 1.0 / o->aodp;
    const double xpidot = o->omgdot + o->xnodot;
    const double sinq = sin(o->xnodeo);
    const double cosq = cos(o->xnodeo);

    /* lunar orbit orientation and the mean anomalies of both bodies at the epoch */
    const double day = ds50 + 18261.5; /* days since 1900 Jan 0.5 */
    const double xnodce = 4.5236020 - 9.2422029E-4 * day;
    const double stem = sin(xnodce);
    const double ctem = cos(xnodce);
    const double zcosil = 0.91375164 - 0.03568096 * ctem;
    const double zsinil = sqrt(1.0 - zcosil * zcosil);
    const double zsinhl = 0.089683511 * stem / zsinil;
    const double zcoshl = sqrt(1.0 - zsinhl * zsinhl);
    const double c = 4.7199672 + 0.22997150 * day;
    const double gam = 5.8351514 + 0.0019443680 * day;
    d->zmol = mod_2pi(c - gam);
    double zx = 0.39785416 * stem / zsinil;
    const double zy = zcoshl * ctem + 0.91744867 * zsinhl * stem;
    zx = arctan4(zx, zy);
    zx = gam + zx - xnodce;
    const double zcosgl = cos(zx);
    const double zsingl = sin(zx);
    d->zmos = mod_2pi(6.2565837 + 0.017201977 * day);
    d->savtsn = 1E20;

    /* the sun, then the moon */
    const body_rates sun = third_body(o, K_ZCOSGS, K_ZSINGS, K_ZCOSIS, K_ZSINIS, cosq, sinq, K_C1SS, K_ZNS, K_ZES, &d->sun);
    d->sse = sun.se;
    d->ssi = sun.si;
    d->ssl = sun.sl;
    d->ssh = sun.sh / o->sinio;
    d->ssg = sun.sgh - o->cosio * d->ssh;
    const body_rates moon = third_body(o, zcosgl, zsingl, zcosil, zsinil, zcoshl * cosq + zsinhl * sinq,
                                       sinq * zcoshl - cosq * zsinhl, K_C1L, K_ZNL, K_ZEL, &d->moon);
    d->sse = d->sse + moon.se;
    d->ssi = d->ssi + moon.si;
    d->ssl = d->ssl + moon.sl;
    d->ssg = d->ssg + moon.sgh - o->cosio / o->sinio * moon.sh;
    d->ssh = d->ssh + moon.sh / o->sinio;

    /* geopotential resonance: 24 h (synchronous) and 12 h (eccentric, e.g. Molniya) orbits */
    d->resonant = 0;
    d->synchronous = 0;
    double bfact;
    if (xnq < 0.0052359877 && xnq > 0.0034906585) {
        d->resonant = 1;
        d->synchronous = 1;
        const double g200 = 1 + d->eosq * (-2.5 + 0.8125 * d->eosq);
        const double g310 = 1 + 2 * d->eosq;
        const double g300 = 1 + d->eosq * (-6 + 6.60937 * d->eosq);
        const double f220 = 0.75 * (1 + o->cosio) * (1 + o->cosio);
        const double f311 = 0.9375 * o->sinio * o->sinio * (1 + 3 * o->cosio) - 0.75 * (1 + o->cosio);
        double f330 = 1 + o->cosio;
        f330 = 1.875 * f330 * f330 * f330;
        d->del1 = 3 * xnq * xnq * aqnv * aqnv;
        d->del2 = 2 * d->del1 * f220 * g200 * K_Q22;
        d->del3 = 3 * d->del1 * f330 * g300 * K_Q33 * aqnv;
        d->del1 = d->del1 * f311 * g310 * K_Q31 * aqnv;
        d->xlamo = o->xmo + o->xnodeo + o->omegao - d->thgr;
        bfact = o->xmdot + xpidot - K_THDT;
        bfact = bfact + d->ssl + d->ssg + d->ssh;
    } else {
        if (xnq < 0.00826 || xnq > 0.00924 || eq < 0.5) {
            return;
        }
        d->resonant = 1;
        const double eoc = eq * d->eosq;
        const double g201 = -0.306 - (eq - 0.64) * 0.440;
        double g211, g310, g322, g410, g422, g520, g533, g521, g532;
        if (eq <= 0.65) {
            g211 = 3.616 - 13.247 * eq + 16.290 * d->eosq;
            g310 = -19.302 + 117.390 * eq - 228.419 * d->eosq + 156.591 * eoc;
            g322 = -18.9068 + 109.7927 * eq - 214.6334 * d->eosq + 146.5816 * eoc;
            g410 = -41.122 + 242.694 * eq - 471.094 * d->eosq + 313.953 * eoc;
            g422 = -146.407 + 841.880 * eq - 1629.014 * d->eosq + 1083.435 * eoc;
            g520 = -532.114 + 3017.977 * eq - 5740 * d->eosq + 3708.276 * eoc;
        } else {
            g211 = -72.099 + 331.819 * eq - 508.738 * d->eosq + 266.724 * eoc;
            g310 = -346.844 + 1582.851 * eq - 2415.925 * d->eosq + 1246.113 * eoc;
            g322 = -342.585 + 1554.908 * eq - 2366.899 * d->eosq + 1215.972 * eoc;
            g410 = -1052.797 + 4758.686 * eq - 7193.992 * d->eosq + 3651.957 * eoc;
            g422 = -3581.69 + 16178.11 * eq - 24462.77 * d->eosq + 12422.52 * eoc;
            if (eq <= 0.715) {
                g520 = 1464.74 - 4664.75 * eq + 3763.64 * d->eosq;
            } else {
                g520 = -5149.66 + 29936.92 * eq - 54087.36 * d->eosq + 31324.56 * eoc;
            }
        }
        if (eq < 0.7) {
            g533 = -919.2277 + 4988.61 * eq - 9064.77 * d->eosq + 5542.21 * eoc;
            g521 = -822.71072 + 4568.6173 * eq - 8491.4146 * d->eosq + 5337.524 * eoc;
            g532 = -853.666 + 4690.25 * eq - 8624.77 * d->eosq + 5341.4 * eoc;
        } else {
            g533 = -37995.78 + 161616.52 * eq - 229838.2 * d->eosq + 109377.94 * eoc;
            g521 = -51752.104 + 218913.95 * eq - 309468.16 * d->eosq + 146349.42 * eoc;
            g532 = -40023.88 + 170470.89 * eq - 242699.48 * d->eosq + 115605.82 * eoc;
        }
        const double sini2 = o->sinio * o->sinio;
        const double f220 = 0.75 * (1 + 2 * o->cosio + d->theta2);
        const double f221 = 1.5 * sini2;
        const double f321 = 1.875 * o->sinio * (1 - 2 * o->cosio - 3 * d->theta2);
        const double f322 = -1.875 * o->sinio * (1 + 2 * o->cosio - 3 * d->theta2);
        const double f441 = 35 * sini2 * f220;
        const double f442 = 39.3750 * sini2 * sini2;
        const double f522 = 9.84375 * o->sinio *
                            (sini2 * (1 - 2 * o->cosio - 5 * d->theta2) + 0.33333333 * (-2 + 4 * o->cosio + 6 * d->theta2));
        const double f523 = o->sinio * (4.92187512 * sini2 * (-2 - 4 * o->cosio + 10 * d->theta2) +
                                        6.56250012 * (1 + 2 * o->cosio - 3 * d->theta2));
        const double f542 = 29.53125 * o->sinio * (2 - 8 * o->cosio + d->theta2 * (-12 + 8 * o->cosio + 10 * d->theta2));
        const double f543 = 29.53125 * o->sinio * (-2 - 8 * o->cosio + d->theta2 * (12 + 8 * o->cosio - 10 * d->theta2));
        const double xno2 = xnq * xnq;
        const double ainv2 = aqnv * aqnv;
        double temp1 = 3 * xno2 * ainv2;
        double temp = temp1 * K_ROOT22;
        d->d2201 = temp * f220 * g201;
        d->d2211 = temp * f221 * g211;
        temp1 = temp1 * aqnv;
        temp = temp1 * K_ROOT32;
        d->d3210 = temp * f321 * g310;
        d->d3222 = temp * f322 * g322;
        temp1 = temp1 * aqnv;
        temp = 2 * temp1 * K_ROOT44;
        d->d4410 = temp * f441 * g410;
        d->d4422 = temp * f442 * g422;
        temp1 = temp1 * aqnv;
        temp = temp1 * K_ROOT52;
        d->d5220 = temp * f522 * g520;
        d->d5232 = temp * f523 * g532;
        temp = 2 * temp1 * K_ROOT54;
        d->d5421 = temp * f542 * g521;
        d->d5433 = temp * f543 * g533;
        d->xlamo = o->xmo + o->xnodeo + o->xnodeo - d->thgr - d->thgr;
        bfact = o->xmdot + o->xnodot + o->xnodot - K_THDT - K_THDT;
        bfact = bfact + d->ssl + d->ssh + d->ssh;
    }
    d->xfact = bfact - xnq;
    d->xli = d->xlamo;
    d->xni = xnq;
    d->atime = 0;
}

/* rates of the resonance variables at the integrator's current state (Deep(), sgp4sdp4.c:877-910) */
static void resonance_rates(const sdrm_orbit *o, double *xndot, double *xnddt) {
    const sdrm_deep_space *d = &o->ds;
    if (d->synchronous) {
        const double fasx2 = 0.13130908;
        const double fasx4 = 2.8843198;
        const double fasx6 = 0.37448087;
        *xndot = d->del1 * sin(d->xli - fasx2) + d->del2 * sin(2 * (d->xli - fasx4)) + d->del3 * sin(3 * (d->xli - fasx6));
        *xnddt = d->del1 * cos(d->xli - fasx2) + 2 * d->del2 * cos(2 * (d->xli - fasx4)) + 3 * d->del3 * cos(3 * (d->xli - fasx6));
        return;
    }
    const double xomi = o->omegao + o->omgdot * d->atime;
    const double x2omi = xomi + xomi;
    const double x2li = d->xli + d->xli;
    *xndot = d->d2201 * sin(x2omi + d->xli - K_G22) + d->d2211 * sin(d->xli - K_G22) + d->d3210 * sin(xomi + d->xli - K_G32) +
             d->d3222 * sin(-xomi + d->xli - K_G32) + d->d4410 * sin(x2omi + x2li - K_G44) + d->d4422 * sin(x2li - K_G44) +
             d->d5220 * sin(xomi + d->xli - K_G52) + d->d5232 * sin(-xomi + d->xli - K_G52) + d->d5421 * sin(xomi + x2li - K_G54) +
             d->d5433 * sin(-xomi + x2li - K_G54);
    *xnddt = d->d2201 * cos(x2omi + d->xli - K_G22) + d->d2211 * cos(d->xli - K_G22) + d->d3210 * cos(xomi + d->xli - K_G32) +
             d->d3222 * cos(-xomi + d->xli - K_G32) + d->d5220 * cos(xomi + d->xli - K_G52) + d->d5232 * cos(-xomi + d->xli - K_G52) +
             2 * (d->d4410 * cos(x2omi + x2li - K_G44) + d->d4422 * cos(x2li - K_G44) + d->d5421 * cos(xomi + x2li - K_G54) +
                  d->d5433 * cos(-xomi + x2li - K_G54));
}

/* Secular lunar/solar effects and, for resonant orbits, the numerical integration of mean motion and longitude from the
 * epoch (or from wherever the previous call left the integrator) to t (Deep() entry dpsec, sgp4sdp4.c:828-929).
 * The integrator walks in 720-minute steps away from the epoch; a request closer to the epoch than its state steps back
 * one step at a time; a request on the other side of the epoch restarts it. */
static void deep_secular(sdrm_orbit *o, double t, double *xll, double *omgadf, double *xnode, double *em, double *xinc,
                         double *xn) {
    sdrm_deep_space *d = &o->ds;
    *xll = *xll + d->ssl * t;
    *omgadf = *omgadf + d->ssg * t;
    *xnode = *xnode + d->ssh * t;
    *em = o->eo + d->sse * t;
    *xinc = o->xincl + d->ssi * t;
    if (*xinc < 0) {
        *xinc = -*xinc;
        *xnode = *xnode + K_PI;
        *omgadf = *omgadf - K_PI;
    }
    if (!d->resonant) {
        return;
    }
    double delt = 0;
    double ft = 0;
    double xndot = 0;
    double xnddt = 0;
    double xldot = 0;
    int stepping;
    int backwards;
    do {
        if (d->atime == 0 || (t >= 0 && d->atime < 0) || (t < 0 && d->atime >= 0)) {
            /* restart at the epoch */
            delt = t >= 0 ? K_RESONANCE_STEP : -K_RESONANCE_STEP;
            d->atime = 0;
            d->xni = o->xnodp;
            d->xli = d->xlamo;
        } else if (fabs(t) >= fabs(d->atime)) {
            delt = t > 0 ? K_RESONANCE_STEP : -K_RESONANCE_STEP;
        }
        do {
            stepping = fabs(t - d->atime) >= K_RESONANCE_STEP;
            if (!stepping) {
                ft = t - d->atime;
            }
            backwards = fabs(t) < fabs(d->atime);
            if (backwards) {
                delt = t >= 0 ? -K_RESONANCE_STEP : K_RESONANCE_STEP;
                stepping = 1;
            }
            resonance_rates(o, &xndot, &xnddt);
            xldot = d->xni + d->xfact;
            xnddt = xnddt * xldot;
            if (stepping) {
                d->xli = d->xli + xldot * delt + xndot * K_RESONANCE_STEP2;
                d->xni = d->xni + xndot * delt + xnddt * K_RESONANCE_STEP2;
                d->atime = d->atime + delt;
            }
        } while (stepping && !backwards);
    } while (stepping && backwards);
    *xn = d->xni + xndot * ft + xnddt * ft * ft * 0.5;
    const double xl = d->xli + xldot * ft + xndot * ft * ft * 0.5;
    const double temp = -*xnode + d->thgr + t * K_THDT;
    if (!d->synchronous) {
        *xll = xl + temp + temp;
    } else {
        *xll = xl - *omgadf + temp;
    }
}

/* Lunar/solar periodics (Deep() entry dpper, sgp4sdp4.c:931-1013); recomputed when t moved by 30 minutes or more */
static void deep_periodics(sdrm_orbit *o, double t, double *em, double *xinc, double *omgadf, double *xnode, double *xll) {
    sdrm_deep_space *d = &o->ds;
    const double sinis = sin(*xinc);
    const double cosis = cos(*xinc);
    if (fabs(d->savtsn - t) >= 30) {
        d->savtsn = t;
        double zm = d->zmos + K_ZNS * t;
        double zf = zm + 2 * K_ZES * sin(zm);
        double sinzf = sin(zf);
        double f2 = 0.5 * sinzf * sinzf - 0.25;
        double f3 = -0.5 * sinzf * cos(zf);
        const double ses = d->sun.e2 * f2 + d->sun.e3 * f3;
        const double sis = d->sun.i2 * f2 + d->sun.i3 * f3;
        const double sls = d->sun.l2 * f2 + d->sun.l3 * f3 + d->sun.l4 * sinzf;
        d->sghs = d->sun.gh2 * f2 + d->sun.gh3 * f3 + d->sun.gh4 * sinzf;
        d->shs = d->sun.h2 * f2 + d->sun.h3 * f3;
        zm = d->zmol + K_ZNL * t;
        zf = zm + 2 * K_ZEL * sin(zm);
        sinzf = sin(zf);
        f2 = 0.5 * sinzf * sinzf - 0.25;
        f3 = -0.5 * sinzf * cos(zf);
        const double sel = d->moon.e2 * f2 + d->moon.e3 * f3;
        const double sil = d->moon.i2 * f2 + d->moon.i3 * f3;
        const double sll = d->moon.l2 * f2 + d->moon.l3 * f3 + d->moon.l4 * sinzf;
        d->sghl = d->moon.gh2 * f2 + d->moon.gh3 * f3 + d->moon.gh4 * sinzf;
        d->shl = d->moon.h2 * f2 + d->moon.h3 * f3;
        d->pe = ses + sel;
        d->pinc = sis + sil;
        d->pl = sls + sll;
    }
    double pgh = d->sghs + d->sghl;
    double ph = d->shs + d->shl;
    *xinc = *xinc + d->pinc;
    *em = *em + d->pe;
    if (o->xincl >= 0.2) {
        /* apply the periodics directly */
        ph = ph / o->sinio;
        pgh = pgh - o->cosio * ph;
        *omgadf = *omgadf + pgh;
        *xnode = *xnode + ph;
        *xll = *xll + d->pl;
    } else {
        /* low inclination: Lyddane's modification, with the node kept continuous */
        const double sinok = sin(*xnode);
        const double cosok = cos(*xnode);
        double alfdp = sinis * sinok;
        double betdp = sinis * cosok;
        const double dalf = ph * cosok + d->pinc * cosis * sinok;
        const double dbet = -ph * sinok + d->pinc * cosis * cosok;
        alfdp = alfdp + dalf;
        betdp = betdp + dbet;
        *xnode = mod_2pi(*xnode);
        double xls = *xll + *omgadf + cosis * *xnode;
        const double dls = d->pl + pgh - d->pinc * *xnode * sinis;
        xls = xls + dls;
        const double xnoh = *xnode;
        *xnode = arctan4(alfdp, betdp);
        if (fabs(xnoh - *xnode) > K_PI) {
            if (*xnode < xnoh) {
                *xnode += K_TWOPI;
            } else {
                *xnode -= K_TWOPI;
            }
        }
        *xll = *xll + d->pl;
        *omgadf = xls - *xll - cos(*xinc) * *xnode;
    }
}

/* SDP4 (sgp4sdp4.c:278-509) */
static void sdp4_propagate(sdrm_orbit *o, double tsince, double pos[3], double vel[3]) {
    if (!o->ready) {
        sgp4_setup(o); /* the drag and J2..J4 secular constants are the same expressions in both models */
        deep_setup(o);
    }
    double xmdf = o->xmo + o->xmdot * tsince;
    double omgadf = o->omegao + o->omgdot * tsince;
    const double xnoddf = o->xnodeo + o->xnodot * tsince;
    const double tsq = tsince * tsince;
    double xnode = xnoddf + o->xnodcf * tsq;
    const double tempa = 1.0 - o->c1 * tsince;
    const double tempe = o->bstar * o->c4 * tsince;
    const double templ = o->t2cof * tsq;
    double xn = o->xnodp;
    double em;
    double xinc;
    deep_secular(o, tsince, &xmdf, &omgadf, &xnode, &em, &xinc, &xn);
    const double a = pow(K_XKE / xn, K_TOTHRD) * tempa * tempa;
    em = em - tempe;
    double xmam = xmdf + o->xnodp * templ;
    deep_periodics(o, tsince, &em, &xinc, &omgadf, &xnode, &xmam);
    const double xl = xmam + omgadf + xnode;
    elements_to_state(o, a, em, omgadf, xnode, xinc, xl, pos, vel);
}

/* ---- observer ---------------------------------------------------------------------------------------------------------------- */

void sdrm_orbit_state(sdrm_orbit *orbit, double tsince, double pos[3], double vel[3]) {
    if (orbit->deep) {
        sdp4_propagate(orbit, tsince, pos, vel);
    } else {
        sgp4_propagate(orbit, tsince, pos, vel);
    }
    /* to km and km/s (Convert_Sat_State) */
    const double kv = K_XKMPER * K_XMNPDA / K_SECDAY;
    for (int k = 0; k < 3; k++) {
        pos[k] *= K_XKMPER;
        vel[k] *= kv;
    }
}

double sdrm_orbit_range_rate(sdrm_orbit *orbit, double jul_utc, double lat_rad, double lon_rad, double alt_km) {
    const double tsince = (jul_utc - orbit->jul_epoch) * K_XMNPDA;
    double pos[3];
    double vel[3];
    sdrm_orbit_state(orbit, tsince, pos, vel);
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
