/* Host-side orbit model used by the Doppler schedule (orbit.c). Internal. */
#ifndef SDRM_ORBIT_H
#define SDRM_ORBIT_H

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
} sdrm_orbit;

/* tle: name line + the two element lines. 0 ok, -1 invalid element set, -2 deep-space orbit (unsupported). */
int sdrm_orbit_init(const char tle[3][80], sdrm_orbit *orbit);

/* km/s, observer at geodetic (lat, lon) in radians and altitude in km, at Julian date jul_utc */
double sdrm_orbit_range_rate(sdrm_orbit *orbit, double jul_utc, double lat_rad, double lon_rad, double alt_km);

double sdrm_julian_date(int year, int month, int mday, int hour, int min, int sec);
double sdrm_julian_date_of_epoch(double epoch);

#endif
