/* TEST INFRASTRUCTURE ONLY (oracle/): never linked or loaded by the product library.
 *
 * Literal C restatement of the reference's own 1-D transform code, src/fft/stafft.f90 of
 * matt-frey/ps3d v0.1.3 -- same factorisation order, same trig table, same out-of-place radix
 * sweeps, same sequential post-processing recurrences -- so that the NumPy/SciPy oracle
 * (oracle/ps3d_oracle.py, which uses library FFTs with the reference's packing) and the CUDA
 * transforms can be pinned against the arithmetic the Fortran build performs, operation by
 * operation, and so that the reference's own round-off floor can be measured (tests/test_stafft_lit.py).
 * The Fortran cannot be compiled in this image (no gfortran), this is the closest available pin.
 *
 *   initfft     <- stafft.f90:76-125      factorisen <- :128-187
 *   forfft      <- :196-287               revfft     <- :296-403
 *   dct         <- :410-483               dst        <- :489-550
 *   forrdx4     <- :901-1014              forrdx3    <- :1017-1098     forrdx2 <- :1101-1151
 *   revrdx4     <- :1503-1617             revrdx3    <- :1620-1703     revrdx2 <- :1706-1757
 * Radices 5 and 6 (:561-898, :1160-1500) are not restated: every BASELINE configuration is a power of
 * two (factorisen: 4^a 2^b); lit_initfft returns 2 for lengths that need them.
 *
 * Index conventions: Fortran a(i, r, k) with shape (0:nv-1, 0:R-1, 0:lv-1) is a[i + nv*(r + R*k)];
 * b(i, k, r) with shape (0:nv-1, 0:lv-1, 0:R-1) is b[i + nv*(k + lv*r)]; cosine(k, j) (0:lv-1, 1:R-1)
 * is cs[k + lv*(j-1)].  The OpenMP loop-order variants of each radix routine perform the same
 * operations per element; one order is kept.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static const double rt2 = 1.414213562373095048801688724209698078569671875376948073176679737990732;
static const double rtf12 = 0.7071067811865475244008443621048490392848359376884740365883398689953662;
static const double sinfpi3 = 0.8660254037844386467637231707529361834714026269051903140279034897259665;
static const double f12 = 0.5;

/* stafft.f90:128-187 */
int lit_factorisen(int n, int factors[5]) {
    static const int fac[5] = {6, 4, 2, 3, 5};
    int rem = n;
    for (int i = 0; i < 5; ++i) factors[i] = 0;
    for (int f = 0; f < 5; ++f) {
        while (rem % fac[f] == 0) {
            factors[f] += 1;
            rem /= fac[f];
            if (rem == 1) return 0;
        }
    }
    return 1;
}

/* stafft.f90:76-125.  trig has 2n entries (the Fortran trig(1:2n); the transforms see it as trig(0:2n-1)).
 * Returns 0, 1 (factorisation impossible: the Fortran stops) or 2 (needs a radix not restated here). */
int lit_initfft(int n, int factors[5], double* trig) {
    static const int fac[5] = {6, 4, 2, 3, 5};
    const double twopi = 6.2831853071795864769252867665590057683943387987502116419498891846156328;
    if (lit_factorisen(n, factors)) return 1;
    double ftwopin = twopi / (double)n;
    int rem = n, m = 0;
    for (int i = 0; i < 2 * n; ++i) trig[i] = 0.0;
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < factors[i]; ++j) {
            rem /= fac[i];
            for (int k = 1; k <= fac[i] - 1; ++k)
                for (int l = 0; l <= rem - 1; ++l) trig[m++] = ftwopin * (double)(k * l);
            ftwopin = ftwopin * fac[i];
        }
    for (int i = 0; i < n - 1; ++i) {      /* Fortran i = 1..n-1 on the 1-based array */
        trig[n + i] = -sin(trig[i]);
        trig[i] = cos(trig[i]);
    }
    return (factors[0] || factors[4]) ? 2 : 0;
}

#define A(i, r, k) a[(i) + nv * ((r) + R * (k))]     /* a(0:nv-1, 0:R-1, 0:lv-1) */
#define B(i, k, r) b[(i) + nv * ((k) + lv * (r))]     /* b(0:nv-1, 0:lv-1, 0:R-1) */
#define CS(k, j) cs[(k) + lv * ((j) - 1)]
#define SN(k, j) sn[(k) + lv * ((j) - 1)]

/* stafft.f90:901-1014 */
static void forrdx4(const double* a, double* b, long nv, long lv, const double* cs, const double* sn) {
    enum { R = 4 };
    for (long i = 0; i < nv; ++i) {
        const double t1r = A(i, 0, 0) + A(i, 2, 0);
        const double t2r = A(i, 1, 0) + A(i, 3, 0);
        B(i, 0, 0) = t1r + t2r;
        B(i, 0, 1) = A(i, 0, 0) - A(i, 2, 0);
        B(i, 0, 2) = t1r - t2r;
        B(i, 0, 3) = A(i, 3, 0) - A(i, 1, 0);
    }
    for (long k = 1; k <= (lv - 1) / 2; ++k) {
        const long kc = lv - k;
        const double c1k = CS(k, 1), s1k = SN(k, 1), c2k = CS(k, 2), s2k = SN(k, 2), c3k = CS(k, 3), s3k = SN(k, 3);
        for (long i = 0; i < nv; ++i) {
            const double x1p = c1k * A(i, 1, k) - s1k * A(i, 1, kc);
            const double y1p = c1k * A(i, 1, kc) + s1k * A(i, 1, k);
            const double x2p = c2k * A(i, 2, k) - s2k * A(i, 2, kc);
            const double y2p = c2k * A(i, 2, kc) + s2k * A(i, 2, k);
            const double x3p = c3k * A(i, 3, k) - s3k * A(i, 3, kc);
            const double y3p = c3k * A(i, 3, kc) + s3k * A(i, 3, k);
            const double t1r = A(i, 0, k) + x2p;
            const double t1i = A(i, 0, kc) + y2p;
            const double t2r = x1p + x3p;
            const double t2i = y1p + y3p;
            const double t3r = A(i, 0, k) - x2p;
            const double t3i = A(i, 0, kc) - y2p;
            const double t4r = x3p - x1p;
            const double t4i = y1p - y3p;
            B(i, k, 0) = t1r + t2r;
            B(i, kc, 0) = t3r - t4i;
            B(i, k, 1) = t3r + t4i;
            B(i, kc, 1) = t1r - t2r;
            B(i, k, 2) = t2i - t1i;
            B(i, kc, 2) = t3i + t4r;
            B(i, k, 3) = t4r - t3i;
            B(i, kc, 3) = t1i + t2i;
        }
    }
    if (lv % 2 == 0) {
        const long lvd2 = lv / 2;
        for (long i = 0; i < nv; ++i) {
            const double q1 = rtf12 * (A(i, 1, lvd2) - A(i, 3, lvd2));
            const double q2 = rtf12 * (A(i, 1, lvd2) + A(i, 3, lvd2));
            B(i, lvd2, 0) = A(i, 0, lvd2) + q1;
            B(i, lvd2, 1) = A(i, 0, lvd2) - q1;
            B(i, lvd2, 2) = A(i, 2, lvd2) - q2;
            B(i, lvd2, 3) = -A(i, 2, lvd2) - q2;
        }
    }
}

/* stafft.f90:1017-1098 */
static void forrdx3(const double* a, double* b, long nv, long lv, const double* cs, const double* sn) {
    enum { R = 3 };
    for (long i = 0; i < nv; ++i) {
        const double t1r = A(i, 1, 0) + A(i, 2, 0);
        B(i, 0, 0) = A(i, 0, 0) + t1r;
        B(i, 0, 1) = A(i, 0, 0) - f12 * t1r;
        B(i, 0, 2) = sinfpi3 * (A(i, 2, 0) - A(i, 1, 0));
    }
    for (long k = 1; k <= (lv - 1) / 2; ++k) {
        const long kc = lv - k;
        const double c1k = CS(k, 1), s1k = SN(k, 1), c2k = CS(k, 2), s2k = SN(k, 2);
        for (long i = 0; i < nv; ++i) {
            const double x1p = c1k * A(i, 1, k) - s1k * A(i, 1, kc);
            const double y1p = c1k * A(i, 1, kc) + s1k * A(i, 1, k);
            const double x2p = c2k * A(i, 2, k) - s2k * A(i, 2, kc);
            const double y2p = c2k * A(i, 2, kc) + s2k * A(i, 2, k);
            const double t1r = x1p + x2p;
            const double t1i = y1p + y2p;
            const double t2r = A(i, 0, k) - f12 * t1r;
            const double t2i = f12 * t1i - A(i, 0, kc);
            const double t3r = sinfpi3 * (x2p - x1p);
            const double t3i = sinfpi3 * (y1p - y2p);
            B(i, k, 0) = A(i, 0, k) + t1r;
            B(i, kc, 0) = t2r - t3i;
            B(i, k, 1) = t2r + t3i;
            B(i, kc, 1) = t3r - t2i;
            B(i, k, 2) = t2i + t3r;
            B(i, kc, 2) = A(i, 0, kc) + t1i;
        }
    }
}

/* stafft.f90:1101-1151 (no k = lv/2 branch in the reference: radix 2 only ever runs at odd lv) */
static void forrdx2(const double* a, double* b, long nv, long lv, const double* cs, const double* sn) {
    enum { R = 2 };
    for (long i = 0; i < nv; ++i) {
        B(i, 0, 0) = A(i, 0, 0) + A(i, 1, 0);
        B(i, 0, 1) = A(i, 0, 0) - A(i, 1, 0);
    }
    for (long k = 1; k <= (lv - 1) / 2; ++k) {
        const long kc = lv - k;
        const double c1k = cs[k], s1k = sn[k];
        for (long i = 0; i < nv; ++i) {
            const double x1 = c1k * A(i, 1, k) - s1k * A(i, 1, kc);
            const double y1 = c1k * A(i, 1, kc) + s1k * A(i, 1, k);
            B(i, k, 0) = A(i, 0, k) + x1;
            B(i, kc, 0) = A(i, 0, k) - x1;
            B(i, k, 1) = y1 - A(i, 0, kc);
            B(i, kc, 1) = A(i, 0, kc) + y1;
        }
    }
}
#undef A
#undef B

#define A(i, k, r) a[(i) + nv * ((k) + lv * (r))]     /* a(0:nv-1, 0:lv-1, 0:R-1) */
#define B(i, r, k) b[(i) + nv * ((r) + R * (k))]      /* b(0:nv-1, 0:R-1, 0:lv-1) */

/* stafft.f90:1503-1617 */
static void revrdx4(const double* a, double* b, long nv, long lv, const double* cs, const double* sn) {
    enum { R = 4 };
    for (long i = 0; i < nv; ++i) {
        const double t1r = A(i, 0, 0) + A(i, 0, 2);
        const double t2r = A(i, 0, 1);
        const double t3r = A(i, 0, 0) - A(i, 0, 2);
        const double t4r = A(i, 0, 3);
        B(i, 0, 0) = t1r + t2r;
        B(i, 1, 0) = t3r + t4r;
        B(i, 2, 0) = t1r - t2r;
        B(i, 3, 0) = t3r - t4r;
    }
    for (long k = 1; k <= (lv - 1) / 2; ++k) {
        const long kc = lv - k;
        const double c1k = CS(k, 1), s1k = SN(k, 1), c2k = CS(k, 2), s2k = SN(k, 2), c3k = CS(k, 3), s3k = SN(k, 3);
        for (long i = 0; i < nv; ++i) {
            const double t1r = A(i, k, 0) + A(i, kc, 1);
            const double t1i = A(i, kc, 3) - A(i, k, 2);
            const double t2r = A(i, k, 1) + A(i, kc, 0);
            const double t2i = A(i, kc, 2) - A(i, k, 3);
            const double t3r = A(i, k, 0) - A(i, kc, 1);
            const double t3i = A(i, kc, 3) + A(i, k, 2);
            const double t4r = A(i, k, 1) - A(i, kc, 0);
            const double t4i = A(i, kc, 2) + A(i, k, 3);
            const double x1p = t3r + t4i;
            const double y1p = t3i - t4r;
            const double x2p = t1r - t2r;
            const double y2p = t1i - t2i;
            const double x3p = t3r - t4i;
            const double y3p = t3i + t4r;
            B(i, 0, k) = t1r + t2r;
            B(i, 0, kc) = t1i + t2i;
            B(i, 1, k) = c1k * x1p - s1k * y1p;
            B(i, 1, kc) = c1k * y1p + s1k * x1p;
            B(i, 2, k) = c2k * x2p - s2k * y2p;
            B(i, 2, kc) = c2k * y2p + s2k * x2p;
            B(i, 3, k) = c3k * x3p - s3k * y3p;
            B(i, 3, kc) = c3k * y3p + s3k * x3p;
        }
    }
    if (lv % 2 == 0) {
        const long lvd2 = lv / 2;
        for (long i = 0; i < nv; ++i) {
            B(i, 0, lvd2) = A(i, lvd2, 0) + A(i, lvd2, 1);
            B(i, 2, lvd2) = A(i, lvd2, 3) - A(i, lvd2, 2);
            const double t3r = A(i, lvd2, 0) - A(i, lvd2, 1);
            const double t4r = A(i, lvd2, 3) + A(i, lvd2, 2);
            B(i, 1, lvd2) = rtf12 * (t3r + t4r);
            B(i, 3, lvd2) = rtf12 * (t4r - t3r);
        }
    }
}

/* stafft.f90:1620-1703 */
static void revrdx3(const double* a, double* b, long nv, long lv, const double* cs, const double* sn) {
    enum { R = 3 };
    for (long i = 0; i < nv; ++i) {
        const double t1r = A(i, 0, 1);
        const double t2r = A(i, 0, 0) - f12 * t1r;
        const double t3r = sinfpi3 * A(i, 0, 2);
        B(i, 0, 0) = A(i, 0, 0) + t1r;
        B(i, 1, 0) = t2r + t3r;
        B(i, 2, 0) = t2r - t3r;
    }
    for (long k = 1; k <= (lv - 1) / 2; ++k) {
        const long kc = lv - k;
        const double c1k = CS(k, 1), s1k = SN(k, 1), c2k = CS(k, 2), s2k = SN(k, 2);
        for (long i = 0; i < nv; ++i) {
            const double t1r = A(i, k, 1) + A(i, kc, 0);
            const double t1i = A(i, kc, 1) - A(i, k, 2);
            const double t2r = A(i, k, 0) - f12 * t1r;
            const double t2i = A(i, kc, 2) - f12 * t1i;
            const double t3r = sinfpi3 * (A(i, k, 1) - A(i, kc, 0));
            const double t3i = sinfpi3 * (A(i, kc, 1) + A(i, k, 2));
            const double x1p = t2r + t3i;
            const double y1p = t2i - t3r;
            const double x2p = t2r - t3i;
            const double y2p = t2i + t3r;
            B(i, 0, k) = A(i, k, 0) + t1r;
            B(i, 0, kc) = A(i, kc, 2) + t1i;
            B(i, 1, k) = c1k * x1p - s1k * y1p;
            B(i, 1, kc) = s1k * x1p + c1k * y1p;
            B(i, 2, k) = c2k * x2p - s2k * y2p;
            B(i, 2, kc) = s2k * x2p + c2k * y2p;
        }
    }
}

/* stafft.f90:1706-1757 */
static void revrdx2(const double* a, double* b, long nv, long lv, const double* cs, const double* sn) {
    enum { R = 2 };
    for (long i = 0; i < nv; ++i) {
        B(i, 0, 0) = A(i, 0, 0) + A(i, 0, 1);
        B(i, 1, 0) = A(i, 0, 0) - A(i, 0, 1);
    }
    for (long k = 1; k <= (lv - 1) / 2; ++k) {
        const long kc = lv - k;
        const double c1k = cs[k], s1k = sn[k];
        for (long i = 0; i < nv; ++i) {
            const double x1p = A(i, k, 0) - A(i, kc, 0);
            const double y1p = A(i, kc, 1) + A(i, k, 1);
            B(i, 0, k) = A(i, k, 0) + A(i, kc, 0);
            B(i, 0, kc) = A(i, kc, 1) - A(i, k, 1);
            B(i, 1, k) = c1k * x1p - s1k * y1p;
            B(i, 1, kc) = c1k * y1p + s1k * x1p;
        }
    }
}
#undef A
#undef B
#undef CS
#undef SN

typedef void (*rdx_fn)(const double*, double*, long, long, const double*, const double*);

/* stafft.f90:196-287.  x(0:m*n-1), vector index fastest; trig(0:2n-1). */
int lit_forfft_wk(int m, int n, double* x, const double* trig, const int factors[5], double* wk);
int lit_forfft(int m, int n, double* x, const double* trig, const int factors[5]) {
    double* wk = (double*)malloc(sizeof(double) * (size_t)m * n);      /* the Fortran's automatic array wk(0:m*n-1) */
    const int rc = lit_forfft_wk(m, n, x, trig, factors, wk);
    free(wk);
    return rc;
}
int lit_forfft_wk(int m, int n, double* x, const double* trig, const int factors[5], double* wk) {
    if (factors[0] || factors[4]) return 2;
    int orig = 1;
    long rem = n, cum = 1;
    /* order of use: 5, 3, 2, 4, 6 (:210-273) */
    static const int order[3] = {3, 2, 1};         /* index into factors: 3 -> radix 3, 2 -> radix 2, 1 -> radix 4 */
    static const int radix[5] = {6, 4, 2, 3, 5};
    static const rdx_fn fn[5] = {0, forrdx4, forrdx2, forrdx3, 0};
    for (int o = 0; o < 3; ++o) {
        const int f = order[o], R = radix[f];
        for (int i = 0; i < factors[f]; ++i) {
            rem = rem / R;
            const long iloc = (rem - 1) * R * cum;
            if (orig) fn[f](x, wk, m * rem, cum, trig + iloc, trig + n + iloc);
            else fn[f](wk, x, m * rem, cum, trig + iloc, trig + n + iloc);
            orig = !orig;
            cum = cum * R;
        }
    }
    const double normfac = 1.0 / sqrt((double)n);
    if (orig) for (long i = 0; i < (long)m * n; ++i) x[i] = x[i] * normfac;
    else for (long i = 0; i < (long)m * n; ++i) x[i] = wk[i] * normfac;
    return 0;
}

/* stafft.f90:296-403 */
int lit_revfft_wk(int m, int n, double* x, const double* trig, const int factors[5], double* wk);
int lit_revfft(int m, int n, double* x, const double* trig, const int factors[5]) {
    double* wk = (double*)malloc(sizeof(double) * (size_t)m * n);
    const int rc = lit_revfft_wk(m, n, x, trig, factors, wk);
    free(wk);
    return rc;
}
int lit_revfft_wk(int m, int n, double* x, const double* trig, const int factors[5], double* wk) {
    if (factors[0] || factors[4]) return 2;
    for (long i = (long)(n / 2 + 1) * m; i < (long)n * m; ++i) x[i] = -x[i];
    for (long i = 0; i < m; ++i) x[i] = f12 * x[i];
    if (n % 2 == 0) {
        const long k = (long)m * n / 2;
        for (long i = 0; i < m; ++i) x[k + i] = f12 * x[k + i];
    }
    int orig = 1;
    long cum = 1, rem = n;
    /* order of use: 6, 4, 2, 3, 5 (:326-389) */
    static const int order[3] = {1, 2, 3};
    static const int radix[5] = {6, 4, 2, 3, 5};
    static const rdx_fn fn[5] = {0, revrdx4, revrdx2, revrdx3, 0};
    for (int o = 0; o < 3; ++o) {
        const int f = order[o], R = radix[f];
        for (int i = 0; i < factors[f]; ++i) {
            rem = rem / R;
            const long iloc = (cum - 1) * R * rem;
            if (orig) fn[f](x, wk, m * cum, rem, trig + iloc, trig + n + iloc);
            else fn[f](wk, x, m * cum, rem, trig + iloc, trig + n + iloc);
            orig = !orig;
            cum = cum * R;
        }
    }
    const double normfac = 2.0 / sqrt((double)n);
    if (orig) for (long i = 0; i < (long)m * n; ++i) x[i] = x[i] * normfac;
    else for (long i = 0; i < (long)m * n; ++i) x[i] = wk[i] * normfac;
    return 0;
}

#define X(i, j) x[(i) + (long)m * (j)]
#define WK(i, j) wk[(i) + (long)m * (j)]

/* stafft.f90:410-483.  x(m, 0:n) */
int lit_dct_wk(int m, int n, double* x, const double* trig, const int factors[5], double* wk2);
int lit_dct(int m, int n, double* x, const double* trig, const int factors[5]) {
    double* wk2 = (double*)malloc(sizeof(double) * (size_t)2 * m * n);
    const int rc = lit_dct_wk(m, n, x, trig, factors, wk2);
    free(wk2);
    return rc;
}
/* wk2: 2*m*n doubles (the automatic arrays of dct and of the forfft it calls) */
int lit_dct_wk(int m, int n, double* x, const double* trig, const int factors[5], double* wk2) {
    const double pi = 3.141592653589793238462643383279502884197169399375105820974944592307816;
    double* wk = wk2;
    const double fpin = pi / (double)n;
    const double rtn = sqrt((double)n);
    for (int i = 0; i < m; ++i) WK(i, 0) = f12 * (X(i, 0) + X(i, n));
    for (int j = 1; j <= n - 1; ++j)
        for (int i = 0; i < m; ++i)
            WK(i, j) = f12 * (X(i, j) + X(i, n - j)) - sin((double)j * fpin) * (X(i, j) - X(i, n - j));
    for (int i = 0; i < m; ++i) {
        double rowsum = 0.0;
        rowsum = rowsum + f12 * X(i, 0);
        for (int j = 1; j <= n - 1; ++j) rowsum = rowsum + X(i, j) * cos((double)j * fpin);
        rowsum = rowsum - f12 * X(i, n);
        X(i, n) = rt2 * rowsum / rtn;
    }
    const int rc = lit_forfft_wk(m, n, wk, trig, factors, wk2 + (size_t)m * n);
    if (rc) return rc;
    for (int i = 0; i < m; ++i) X(i, 0) = rt2 * WK(i, 0);
    for (int i = 0; i < m; ++i) X(i, 1) = X(i, n);
    if (n % 2 == 0) {
        const int nd2 = n / 2;
        for (int j = 1; j <= nd2 - 1; ++j)
            for (int i = 0; i < m; ++i) {
                X(i, 2 * j) = rt2 * WK(i, j);
                X(i, 2 * j + 1) = X(i, 2 * j - 1) - rt2 * WK(i, n - j);
            }
        for (int i = 0; i < m; ++i) X(i, n) = rt2 * WK(i, nd2);
    } else {
        for (int j = 1; j <= (n - 1) / 2; ++j)
            for (int i = 0; i < m; ++i) {
                X(i, 2 * j) = rt2 * WK(i, j);
                X(i, 2 * j + 1) = X(i, 2 * j - 1) - rt2 * WK(i, n - j);
            }
    }
    return 0;
}
#undef X

/* stafft.f90:489-550.  x(m, n) = x(m, 1:n): X(i, j) below takes the Fortran j = 1..n */
#define X(i, j) x[(i) + (long)m * ((j) - 1)]
int lit_dst_wk(int m, int n, double* x, const double* trig, const int factors[5], double* wk2);
int lit_dst(int m, int n, double* x, const double* trig, const int factors[5]) {
    double* wk2 = (double*)malloc(sizeof(double) * (size_t)2 * m * n);
    const int rc = lit_dst_wk(m, n, x, trig, factors, wk2);
    free(wk2);
    return rc;
}
int lit_dst_wk(int m, int n, double* x, const double* trig, const int factors[5], double* wk2) {
    const double pi = 3.141592653589793238462643383279502884197169399375105820974944592307816;
    double* wk = wk2;
    const double fpin = pi / (double)n;
    for (int i = 0; i < m; ++i) WK(i, 0) = 0.0;
    for (int j = 1; j <= n - 1; ++j)
        for (int i = 0; i < m; ++i)
            WK(i, j) = f12 * (X(i, j) - X(i, n - j)) + sin((double)j * fpin) * (X(i, j) + X(i, n - j));
    const int rc = lit_forfft_wk(m, n, wk, trig, factors, wk2 + (size_t)m * n);
    if (rc) return rc;
    for (int i = 0; i < m; ++i) X(i, 1) = WK(i, 0) / rt2;
    if (n % 2 == 0) {
        for (int j = 1; j <= n / 2 - 1; ++j) {
            for (int i = 0; i < m; ++i) X(i, 2 * j) = -rt2 * WK(i, n - j);
            for (int i = 0; i < m; ++i) X(i, 2 * j + 1) = rt2 * WK(i, j) + X(i, 2 * j - 1);
        }
    } else {
        for (int j = 1; j <= (n - 1) / 2 - 1; ++j)
            for (int i = 0; i < m; ++i) {
                X(i, 2 * j) = -rt2 * WK(i, n - j);
                X(i, 2 * j + 1) = rt2 * WK(i, j) + X(i, 2 * j - 1);
            }
        for (int i = 0; i < m; ++i) X(i, n - 1) = -rt2 * WK(i, (n + 1) / 2);
    }
    for (int i = 0; i < m; ++i) X(i, n) = 0.0;
    return 0;
}
#undef X
#undef WK
