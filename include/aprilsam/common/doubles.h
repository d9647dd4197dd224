/* doubles.h -- 2-D rigid-body (x, y, theta) helpers used by callers of the solver
 * (the reference demo dead-reckons new poses with them:
 * aprilsam/common/doubles_floats_impl.h:498-506,569-575,619-630).  Own code. */
#ifndef ASAM_DOUBLES_H
#define ASAM_DOUBLES_H

#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline double *doubles_dup(const double *v, int len)
{
    if (!v) /* reference: common/doubles_floats_impl.h:66-75 */
        return NULL;
    double *r = (double *) malloc(sizeof(double) * len);
    memcpy(r, v, sizeof(double) * len);
    return r;
}

/* r = a (+) b : pose b expressed in a's frame, composed onto a */
static inline void doubles_xyt_mul(const double a[3], const double b[3], double r[3])
{
    double s = sin(a[2]), c = cos(a[2]);
    double x = c * b[0] - s * b[1] + a[0];
    double y = s * b[0] + c * b[1] + a[1];
    double t = a[2] + b[2];
    r[0] = x; r[1] = y; r[2] = t;
}

/* r = (-) a */
static inline void doubles_xyt_inv(const double a[3], double r[3])
{
    double s = sin(a[2]), c = cos(a[2]);
    double x = -s * a[1] - c * a[0];
    double y = -c * a[1] + s * a[0];
    r[0] = x; r[1] = y; r[2] = -a[2];
}

/* r = (-) a (+) b : b seen from a */
static inline void doubles_xyt_inv_mul(const double a[3], const double b[3], double r[3])
{
    double s = sin(a[2]), c = cos(a[2]);
    double dx = b[0] - a[0], dy = b[1] - a[1];
    r[0] = c * dx + s * dy;
    r[1] = -s * dx + c * dy;
    r[2] = b[2] - a[2];
}

/* wrap to [-pi, pi)  (reference: aprilsam/common/math_util.h:107-122) */
/* degrees <-> radians (common/math_util.h:50-51; the tutorial uses them) */
#ifndef to_radians
#define to_radians(x) ((x) * (3.14159265358979323846 / 180.0))
#define to_degrees(x) ((x) * (180.0 / 3.14159265358979323846))
#endif

static inline double mod2pi(double v)
{
    const double twopi = 6.283185307179586476925287;
    const double pi = 3.141592653589793238462643;
    double w = v + pi;
    return (w - twopi * floor(w / twopi)) - pi;
}

#endif
