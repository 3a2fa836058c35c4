/* Host-side orbit model used by the Doppler schedule (orbit.c). Internal. */
#ifndef SDRM_ORBIT_H
#define SDRM_ORBIT_H

/* periodic coefficients one perturbing body (sun or moon) contributes to e, i, mean anomaly, perigee, node */
typedef struct {
    double e2, e3, i2, i3, l2, l3, l4, gh2, gh3, gh4, h2, h3;
} sdrm_body_terms;

/* deep-space (SDP4) state: lunar/solar perturbations and the 12 h / 24 h resonance integrator */
typedef struct {
    double theta2, eosq, betao2, betao, sing, cosg;
    double thgr;               /* Greenwich sidereal angle at the epoch */
    double zmos, zmol;         /* solar / lunar mean anomalies at the epoch */
    sdrm_body_terms sun, moon;
    double sse, ssi, ssl, ssg, ssh; /* secular rates, both bodies */
    int resonant;
    int synchronous;
    double d2201, d2211, d3210, d3222, d4410, d4422, d5220, d5232, d5421, d5433; /* 12 h resonance */
    double del1, del2, del3;   /* 24 h resonance */
    double xlamo, xfact;
    double xli, xni, atime;    /* integrator */
    double savtsn;             /* time of the cached periodics */
    double pe, pinc, pl, sghs, shs, sghl, shl;
} sdrm_deep_space;

typedef struct {
    /* elements (after unit conversion) */
    double epoch;
    double jul_epoch;
    double bstar;
    double xincl;
    double xnodeo;
    double eo;
    double omegao;
    double xmo;
    double xno;
    /* SGP4 constants, filled at the first propagation */
    int ready;
    int simple;
    double cosio, sinio, x3thm1, x1mth2, x7thm1, xnodp, aodp, eta, c1, c4, c5, xmdot, omgdot, xnodot, omgcof, xmcof, xnodcf, t2cof,
        xlcof, aycof, delmo, sinmo, d2, d3, d4, t3cof, t4cof, t5cof;
    /* period >= 225 min: SDP4 */
    int deep;
    sdrm_deep_space ds;
} sdrm_orbit;

/* tle: name line + the two element lines. 0 ok, -1 invalid element set. */
int sdrm_orbit_init(const char tle[3][80], sdrm_orbit *orbit);

/* ECI position (km) and velocity (km/s) tsince minutes after the element set's epoch (SGP4 or SDP4 by orbital period) */
void sdrm_orbit_state(sdrm_orbit *orbit, double tsince, double pos_km[3], double vel_km_s[3]);

/* km/s, observer at geodetic (lat, lon) in radians and altitude in km, at Julian date jul_utc */
double sdrm_orbit_range_rate(sdrm_orbit *orbit, double jul_utc, double lat_rad, double lon_rad, double alt_km);

double sdrm_julian_date(int year, int month, int mday, int hour, int min, int sec);
double sdrm_julian_date_of_epoch(double epoch);

#endif
