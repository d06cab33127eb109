// TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C++/OpenMP restatement of the reference's time-step path (matt-frey/ps3d, one rank) with the reference's own
// sweep structure: every 2-D FFT is four global transposes plus two sweeps of contiguous real FFTs
// (sta3dfft.f90:136-260, fft_pencil.f90:283-330), diffx/diffy go through a reversed copy
// (sta3dfft.f90:304-377, mpi_reverse.f90:332-398), the eight N-sized tables of init_inversion are stored
// (inversion_utils.f90:222-372), the Crank-Nicolson update does its combine -> vdiss -> decompose pairs
// literally (cn2.f90:120-135), the strain eigenvalues come from the cyclic Jacobi iteration (jacobi.f90).
// It is "a restatement, not the Fortran build" (SURVEY.md 8d): the Fortran toolchain is absent from the image.
//
// Used only by tests/ (cross-check against oracle/ps3d_oracle.py) and by bench.py's CPU legs (cpu_baseline,
// --impl reference), where it is timed on all host cores.  Nothing under ps3d_b200/ links or loads it.
//
// Conventions pinned in SURVEY.md a1/a2: forfft output is Hermitian-packed, y[k] = Re X_k / sqrt(n),
// y[n-k] = Im X_k / sqrt(n), X = sum_j x_j exp(-2 pi i jk/n); revfft is its inverse; dst/dct are DST-I / DCT-I
// scaled sqrt(2/n) through the reference's reduction to a real FFT of length n (stafft.f90:410-550).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

// The 1-D transforms are the literal restatement of the reference's own kernels (oracle/stafft_lit.c <- stafft.f90:
// radix-4/2 decimation in time / frequency, out-of-place sweeps, sequential DST/DCT recurrences, sin/cos evaluated per
// element in the pre-processing), one line per call as sta3dfft.f90:161-178 does.  -DPS3D_REF_TEXTBOOK_FFT selects the
// earlier self-contained radix-2 transforms instead.
#ifndef PS3D_REF_TEXTBOOK_FFT
extern "C" {
int lit_initfft(int n, int factors[5], double* trig);
int lit_forfft_wk(int m, int n, double* x, const double* trig, const int factors[5], double* wk);
int lit_revfft_wk(int m, int n, double* x, const double* trig, const int factors[5], double* wk);
int lit_dct_wk(int m, int n, double* x, const double* trig, const int factors[5], double* wk2);
int lit_dst_wk(int m, int n, double* x, const double* trig, const int factors[5], double* wk2);
}
#endif

namespace {

typedef std::vector<double> vec;
typedef std::complex<double> cplx;
const double PI = 3.14159265358979323846264338327950288;

// ---- stafft.f90: real FFT of one contiguous line (power-of-two n), Hermitian-packed ---------------------------
struct Fft {
    int n = 0, h = 0;
    std::vector<cplx> w, wr;     // exp(-2 pi i k / h), k < h/2;   exp(-2 pi i k / n), k <= h
    std::vector<int> rev;
    std::vector<double> trig;    // initfft (stafft.f90:76-125)
    int factors[5] = {0, 0, 0, 0, 0};
    void init(int n_) {
        n = n_; h = n / 2;
#ifndef PS3D_REF_TEXTBOOK_FFT
        trig.assign(2 * (size_t)n, 0.0);
        lit_initfft(n, factors, trig.data());
#endif
        w.resize(std::max(1, h / 2)); wr.resize(h + 1); rev.resize(h);
        for (int k = 0; k < h / 2; ++k) w[k] = std::polar(1.0, -2.0 * PI * k / h);
        for (int k = 0; k <= h; ++k) wr[k] = std::polar(1.0, -2.0 * PI * k / n);
        int bits = 0;
        while ((1 << bits) < h) ++bits;
        for (int i = 0; i < h; ++i) {
            int r = 0;
            for (int b = 0; b < bits; ++b) if (i & (1 << b)) r |= 1 << (bits - 1 - b);
            rev[i] = r;
        }
    }
    // in-place complex FFT of length h, sign = -1 forward, +1 inverse (unnormalised)
    void cfft(cplx* z, int sign) const {
        for (int i = 0; i < h; ++i) if (rev[i] > i) std::swap(z[i], z[rev[i]]);
        for (int len = 2; len <= h; len <<= 1) {
            const int half = len / 2, step = h / len;
            for (int i = 0; i < h; i += len)
                for (int k = 0; k < half; ++k) {
                    cplx t = w[k * step];
                    if (sign > 0) t = std::conj(t);
                    const cplx u = z[i + k], v = z[i + k + half] * t;
                    z[i + k] = u + v; z[i + k + half] = u - v;
                }
        }
    }
    // unnormalised spectrum X_0..X_h of the real line x[0..n)
    void spectrum(const double* x, cplx* X, cplx* z) const {
        for (int j = 0; j < h; ++j) z[j] = cplx(x[2 * j], x[2 * j + 1]);
        cfft(z, -1);
        for (int k = 0; k <= h; ++k) {
            const cplx zk = z[k % h], zm = std::conj(z[(h - k) % h]);
            const cplx e = 0.5 * (zk + zm), o = cplx(0.0, -0.5) * (zk - zm);
            X[k] = e + wr[k] * o;
        }
    }
    void forfft(double* x, cplx* X, cplx* z) const {         // stafft.f90:196-287
#ifndef PS3D_REF_TEXTBOOK_FFT
        (void)z;
        lit_forfft_wk(1, n, x, trig.data(), factors, reinterpret_cast<double*>(X));     // X: n + 2 doubles of work space
        return;
#endif
        spectrum(x, X, z);
        const double s = 1.0 / std::sqrt((double)n);
        x[0] = X[0].real() * s; x[h] = X[h].real() * s;
        for (int k = 1; k < h; ++k) { x[k] = X[k].real() * s; x[n - k] = X[k].imag() * s; }
    }
    void revfft(double* x, cplx* X, cplx* z) const {         // stafft.f90:296-403
#ifndef PS3D_REF_TEXTBOOK_FFT
        (void)z;
        lit_revfft_wk(1, n, x, trig.data(), factors, reinterpret_cast<double*>(X));
        return;
#endif
        X[0] = cplx(x[0], 0.0); X[h] = cplx(x[h], 0.0);
        for (int k = 1; k < h; ++k) X[k] = cplx(x[k], x[n - k]);
        for (int k = 0; k < h; ++k) {
            const cplx xk = X[k], xm = std::conj(X[h - k]);
            const cplx e = 0.5 * (xk + xm), o = 0.5 * (xk - xm) * std::conj(wr[k]);
            z[k] = e + cplx(0.0, 1.0) * o;
        }
        cfft(z, +1);
        const double s = 2.0 / std::sqrt((double)n);          // stafft.f90:393
        for (int j = 0; j < h; ++j) { x[2 * j] = z[j].real() * s; x[2 * j + 1] = z[j].imag() * s; }
    }
};

struct Ref {
    int nx, ny, nz, nzp;
    size_t N, NC;                       // nx*ny*nzp, nx*ny
    double lower[3], extent[3], upper[3], dx[3];
    double dzi, hdzi, ncelli, fnzi;
    int filtering;                      // 0 Hou & Li, 1 2/3-rule
    Fft fx, fy, fz;
    vec hrkx, hrky, rkx, rky, rkz, sinz, cosz;
    vec k2l2, k2l2i, vhdis, vdiss;
    vec filt, green, phim, phip, thetam, thetap, dthetam, dthetap, gamtop, gambot;
    vec svor[3], vor[3], vel[3], svel[3], svorts[3], vortsm[3];
    double ini_mean[2];
    int nnu = 3;
    // rolling mean (rolling_mean.f90)
    vec hist; int inew = 1, iold = 1, rlen = 0; double sma = 0.0; bool filled = false;
    double vorch = 0.0, ggmax = 0.0;

    size_t idx(int x, int y, int z) const { return ((size_t)x * ny + y) * nzp + z; }
};

// ---- transposes (fft_pencil.f90:283-330 on one rank: a permuted full copy each) --------------------------------
// [a][b][c] -> [c][b][a]
static void transpose_cba(const double* in, double* out, int na, int nb, int nc) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int c = 0; c < nc; ++c)
        for (int b = 0; b < nb; ++b)
            for (int a = 0; a < na; ++a) out[((size_t)c * nb + b) * na + a] = in[((size_t)a * nb + b) * nc + c];
}
// [a][b][c] -> [a][c][b]
static void transpose_acb(const double* in, double* out, int na, int nb, int nc) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int a = 0; a < na; ++a)
        for (int c = 0; c < nc; ++c)
            for (int b = 0; b < nb; ++b) out[((size_t)a * nc + c) * nb + b] = in[((size_t)a * nb + b) * nc + c];
}
static void fft_lines(const Fft& f, double* data, size_t nlines, bool inv) {
#pragma omp parallel
    {
        std::vector<cplx> X(f.h + 1), z(std::max(1, f.h));
#pragma omp for schedule(static)
        for (long long l = 0; l < (long long)nlines; ++l) {
            if (inv) f.revfft(data + (size_t)l * f.n, X.data(), z.data());
            else f.forfft(data + (size_t)l * f.n, X.data(), z.data());
        }
    }
}

// fftxyp2s (sta3dfft.f90:136-194): z-pencil -> y-pencil, FFT in y, -> x-pencil, FFT in x, and two transposes back
static void fftxyp2s(const Ref& r, const double* fp, double* fs) {
    vec a(r.N), b(r.N);
    transpose_acb(fp, a.data(), r.nx, r.ny, r.nzp);                   // [x][y][z] -> [x][z][y]
    fft_lines(r.fy, a.data(), (size_t)r.nx * r.nzp, false);
    transpose_cba(a.data(), b.data(), r.nx, r.nzp, r.ny);             // [x][z][ky] -> [ky][z][x]
    fft_lines(r.fx, b.data(), (size_t)r.ny * r.nzp, false);
    transpose_cba(b.data(), a.data(), r.ny, r.nzp, r.nx);             // [ky][z][kx] -> [kx][z][ky]
    transpose_acb(a.data(), fs, r.nx, r.nzp, r.ny);                   // [kx][z][ky] -> [kx][ky][z]
}
// fftxys2p (sta3dfft.f90:202-260)
static void fftxys2p(const Ref& r, const double* fs, double* fp) {
    vec a(r.N), b(r.N);
    transpose_acb(fs, a.data(), r.nx, r.ny, r.nzp);                   // [kx][ky][z] -> [kx][z][ky]
    transpose_cba(a.data(), b.data(), r.nx, r.nzp, r.ny);             // -> [ky][z][kx]
    fft_lines(r.fx, b.data(), (size_t)r.ny * r.nzp, true);
    transpose_cba(b.data(), a.data(), r.ny, r.nzp, r.nx);             // [ky][z][x] -> [x][z][ky]
    fft_lines(r.fy, a.data(), (size_t)r.nx * r.nzp, true);
    transpose_acb(a.data(), fp, r.nx, r.nzp, r.ny);                   // [x][z][y] -> [x][y][z]
}

// ---- stafft.f90:410-550: dct / dst of one column through a real FFT of length n --------------------------------
static void dct_col(const Ref& r, double* x, cplx* X, cplx* z, double* y) {
    const int n = r.nz;
#ifndef PS3D_REF_TEXTBOOK_FFT
    (void)X; (void)z;
    lit_dct_wk(1, n, x, r.fz.trig.data(), r.fz.factors, y);          // y: 2 n doubles of work space
    return;
#endif
    double x1 = 0.5 * (x[0] - x[n]);
    for (int j = 1; j < n; ++j) x1 += x[j] * r.cosz[j];                                  // :440-448
    y[0] = 0.5 * (x[0] + x[n]);
    for (int j = 1; j < n; ++j) y[j] = 0.5 * (x[j] + x[n - j]) - r.sinz[j] * (x[j] - x[n - j]);
    r.fz.spectrum(y, X, z);
    const double s = std::sqrt(2.0 / (double)n);
    x[0] = s * X[0].real(); x[n] = s * X[n / 2].real();
    double run = x1;
    x[1] = s * run;
    for (int k = 1; k < n / 2; ++k) {                                                     // :466-471
        x[2 * k] = s * X[k].real();
        run -= X[k].imag();
        x[2 * k + 1] = s * run;
    }
}
// x points at slot 1 of the reference's x(1:n): transforms slots 1..n-1, sets slot n to 0 (:546-549)
static void dst_col(const Ref& r, double* x1, cplx* X, cplx* z, double* y) {
    const int n = r.nz;
#ifndef PS3D_REF_TEXTBOOK_FFT
    (void)X; (void)z;
    lit_dst_wk(1, n, x1, r.fz.trig.data(), r.fz.factors, y);         // x1 = x(1:n)
    return;
#endif
    double* x = x1 - 1;                                                                   // x[j], j = 1..n
    y[0] = 0.0;
    for (int j = 1; j < n; ++j) y[j] = 0.5 * (x[j] - x[n - j]) + r.sinz[j] * (x[j] + x[n - j]);
    r.fz.spectrum(y, X, z);
    const double s = std::sqrt(2.0 / (double)n);
    double run = 0.5 * X[0].real();
    x[1] = s * run;
    for (int k = 1; k < n / 2; ++k) {                                                     // :526-533
        x[2 * k] = -s * X[k].imag();
        run += X[k].real();
        x[2 * k + 1] = s * run;
    }
    x[n] = 0.0;
}
struct ZWork { std::vector<cplx> X, z; vec y; explicit ZWork(int n) : X(n / 2 + 1), z(std::max(1, n / 2)), y(2 * (size_t)n + 2) {} };

// sta3dfft.f90:264-296 fftsine / fftcosine on every column
static void fftcosine(const Ref& r, double* f) {
#pragma omp parallel
    {
        ZWork w(r.nz);
#pragma omp for schedule(static)
        for (long long c = 0; c < (long long)r.NC; ++c) dct_col(r, f + (size_t)c * r.nzp, w.X.data(), w.z.data(), w.y.data());
    }
}
static void fftsine(const Ref& r, double* f) {
#pragma omp parallel
    {
        ZWork w(r.nz);
#pragma omp for schedule(static)
        for (long long c = 0; c < (long long)r.NC; ++c) dst_col(r, f + (size_t)c * r.nzp + 1, w.X.data(), w.z.data(), w.y.data());
    }
}

// ---- sta3dfft.f90:304-377 diffx / diffy through a reversed copy (mpi_reverse.f90:332-398) ----------------------
static void diffx(const Ref& r, const double* fs, double* ds) {
    const int nx = r.nx, nwx = nx / 2;
    const size_t plane = (size_t)r.ny * r.nzp;
    vec gs(r.N);
#pragma omp parallel for schedule(static)
    for (int kx = 0; kx < nx; ++kx) std::memcpy(&gs[(size_t)kx * plane], fs + (size_t)((nx - kx) % nx) * plane, plane * sizeof(double));
#pragma omp parallel for schedule(static)
    for (int kx = 0; kx < nx; ++kx) {
        double* d = ds + (size_t)kx * plane;
        if (kx == 0 || kx == nwx) { std::fill(d, d + plane, 0.0); continue; }
        const int dkx = std::min(2 * kx, 2 * (nx - kx));
        const double f = ((kx >= nwx + 1) ? 1.0 : -1.0) * r.hrkx[dkx - 1];
        const double* g = &gs[(size_t)kx * plane];
        for (size_t i = 0; i < plane; ++i) d[i] = f * g[i];
    }
}
static void diffy(const Ref& r, const double* fs, double* ds) {
    const int ny = r.ny, nwy = ny / 2;
    vec gs(r.N);
#pragma omp parallel for collapse(2) schedule(static)
    for (int kx = 0; kx < r.nx; ++kx)
        for (int ky = 0; ky < ny; ++ky)
            std::memcpy(&gs[r.idx(kx, ky, 0)], fs + r.idx(kx, (ny - ky) % ny, 0), r.nzp * sizeof(double));
#pragma omp parallel for collapse(2) schedule(static)
    for (int kx = 0; kx < r.nx; ++kx)
        for (int ky = 0; ky < ny; ++ky) {
            double* d = ds + r.idx(kx, ky, 0);
            if (ky == 0 || ky == nwy) { std::fill(d, d + r.nzp, 0.0); continue; }
            const int dky = std::min(2 * ky, 2 * (ny - ky));
            const double f = ((ky >= nwy + 1) ? 1.0 : -1.0) * r.hrky[dky - 1];
            const double* g = &gs[r.idx(kx, ky, 0)];
            for (int z = 0; z < r.nzp; ++z) d[z] = f * g[z];
        }
}
static void central_diffz(const Ref& r, const double* fs, double* ds) {        // inversion_utils.f90:653-673
    const int nz = r.nz;
#pragma omp parallel for schedule(static)
    for (long long c = 0; c < (long long)r.NC; ++c) {
        const double* f = fs + (size_t)c * r.nzp;
        double* d = ds + (size_t)c * r.nzp;
        d[0] = r.dzi * (f[1] - f[0]);
        d[nz] = r.dzi * (f[nz] - f[nz - 1]);
        for (int z = 1; z < nz; ++z) d[z] = (f[z + 1] - f[z - 1]) * r.hdzi;
    }
}

// ---- inversion_utils.f90:549-647 ---------------------------------------------------------------------------
static void decompose_semi_spectral(const Ref& r, double* f) {
    const int nz = r.nz;
#pragma omp parallel
    {
        ZWork w(nz);
#pragma omp for schedule(static)
        for (long long c = 0; c < (long long)r.NC; ++c) {
            double* x = f + (size_t)c * r.nzp;
            const double* pm = &r.phim[(size_t)c * r.nzp];
            const double* pp = &r.phip[(size_t)c * r.nzp];
            const double b = x[0], t = x[nz];
            for (int z = 1; z < nz; ++z) x[z] -= b * pm[z] + t * pp[z];
            dst_col(r, x + 1, w.X.data(), w.z.data(), w.y.data());
            x[nz] = t;
        }
    }
}
static void combine_semi_spectral(const Ref& r, double* f) {
    const int nz = r.nz;
#pragma omp parallel
    {
        ZWork w(nz);
#pragma omp for schedule(static)
        for (long long c = 0; c < (long long)r.NC; ++c) {
            double* x = f + (size_t)c * r.nzp;
            const double* pm = &r.phim[(size_t)c * r.nzp];
            const double* pp = &r.phip[(size_t)c * r.nzp];
            const double b = x[0], t = x[nz];
            x[nz] = 0.0;
            dst_col(r, x + 1, w.X.data(), w.z.data(), w.y.data());
            x[nz] = t;
            for (int z = 1; z < nz; ++z) x[z] += b * pm[z] + t * pp[z];
        }
    }
}
static void decompose_physical(const Ref& r, const double* fc, double* sf) { fftxyp2s(r, fc, sf); decompose_semi_spectral(r, sf); }
static void combine_physical(const Ref& r, const double* sf, double* fc) {
    vec t(sf, sf + r.N);
    combine_semi_spectral(r, t.data());
    fftxys2p(r, t.data(), fc);
}

// ---- init (sta3dfft.f90:53-110, inversion_utils.f90:222-542) --------------------------------------------------
static void init_tables(Ref& r) {
    const int nx = r.nx, ny = r.ny, nz = r.nz, nzp = r.nzp;
    r.hrkx.resize(nx); r.hrky.resize(ny); r.rkx.assign(nx, 0.0); r.rky.assign(ny, 0.0); r.rkz.assign(nzp, 0.0);
    for (int j = 1; j <= nx; ++j) r.hrkx[j - 1] = PI / r.extent[0] * j;
    for (int j = 1; j <= ny; ++j) r.hrky[j - 1] = PI / r.extent[1] * j;
    for (int k = 1; k < nx / 2; ++k) { r.rkx[k] = r.hrkx[2 * k - 1]; r.rkx[nx - k] = r.rkx[k]; }
    r.rkx[nx / 2] = r.hrkx[nx - 1];
    for (int k = 1; k < ny / 2; ++k) { r.rky[k] = r.hrky[2 * k - 1]; r.rky[ny - k] = r.rky[k]; }
    r.rky[ny / 2] = r.hrky[ny - 1];
    for (int k = 1; k <= nz; ++k) r.rkz[k] = PI / r.extent[2] * k;
    r.sinz.resize(nzp); r.cosz.resize(nzp);
    for (int j = 0; j <= nz; ++j) { r.sinz[j] = std::sin(PI * j / nz); r.cosz[j] = std::cos(PI * j / nz); }
    r.k2l2.resize(r.NC); r.k2l2i.resize(r.NC);
    for (int kx = 0; kx < nx; ++kx)
        for (int ky = 0; ky < ny; ++ky) {
            const double v = r.rkx[kx] * r.rkx[kx] + r.rky[ky] * r.rky[ky];
            r.k2l2[(size_t)kx * ny + ky] = v;
            r.k2l2i[(size_t)kx * ny + ky] = (kx == 0 && ky == 0) ? 0.0 : 1.0 / v;
        }
    const double kxm = *std::max_element(r.rkx.begin(), r.rkx.end()), kym = *std::max_element(r.rky.begin(), r.rky.end());
    const double kzm = r.rkz[nz];
    for (vec* t : {&r.filt, &r.green, &r.phim, &r.phip, &r.thetam, &r.thetap, &r.dthetam, &r.dthetap}) t->assign(r.N, 0.0);
    vec zm(nzp), zp(nzp);
    for (int iz = 0; iz <= nz; ++iz) { const double z = r.lower[2] + r.dx[2] * iz; zm[iz] = r.upper[2] - z; zp[iz] = z - r.lower[2]; }
#pragma omp parallel for collapse(2) schedule(static)
    for (int kx = 0; kx < nx; ++kx)
        for (int ky = 0; ky < ny; ++ky) {
            const size_t c = (size_t)kx * ny + ky, o = c * nzp;
            // filter (inversion_utils.f90:377-455)
            double f2;
            if (r.filtering == 0) f2 = std::exp(-36.0 * std::pow(r.rkx[kx] / kxm, 36) - 36.0 * std::pow(r.rky[ky] / kym, 36));
            else f2 = ((r.rkx[kx] <= 2.0 / 3.0 * kxm) ? 1.0 : 0.0) * ((r.rky[ky] <= 2.0 / 3.0 * kym) ? 1.0 : 0.0);
            for (int z = 0; z <= nz; ++z) {
                double fz = 1.0;
                if (z >= 1 && z < nz)
                    fz = (r.filtering == 0) ? std::exp(-36.0 * std::pow(r.rkz[z] / kzm, 36)) : ((r.rkz[z] <= 2.0 / 3.0 * kzm) ? 1.0 : 0.0);
                r.filt[o + z] = (kx == 0 && ky == 0) ? 1.0 : f2 * fz;
                r.green[o + z] = (z == 0) ? -r.k2l2i[c] : -1.0 / (r.k2l2[c] + r.rkz[z] * r.rkz[z]);
            }
            // hyperbolic functions (:484-542); (0,0): linear (:326-346)
            if (kx == 0 && ky == 0) {
                for (int z = 0; z <= nz; ++z) { r.phim[o + z] = zm[z] / r.extent[2]; r.phip[o + z] = zp[z] / r.extent[2]; }
                continue;
            }
            const double kl = std::sqrt(r.k2l2[c]), ef = std::exp(-kl * r.extent[2]), div = 1.0 / (1.0 - ef * ef);
            const double k2ifac = 0.5 * r.k2l2i[c], Q = div * (1.0 + ef * ef), R = div * 2.0 * ef;
            for (int z = 0; z <= nz; ++z) {
                const double Lm = kl * zm[z], Lp = kl * zp[z], ep = std::exp(-Lp), em = std::exp(-Lm);
                const double pm = div * (ep - ef * em), pp = div * (em - ef * ep);
                const double dpm = -kl * div * (ep + ef * em), dpp = kl * div * (em + ef * ep);
                r.phim[o + z] = pm; r.phip[o + z] = pp;
                r.thetam[o + z] = k2ifac * (R * Lm * pp - Q * Lp * pm);
                r.thetap[o + z] = k2ifac * (R * Lp * pm - Q * Lm * pp);
                r.dthetam[o + z] = -k2ifac * ((Q * Lp - 1.0) * dpm - R * Lm * dpp);
                r.dthetap[o + z] = -k2ifac * ((Q * Lm - 1.0) * dpp - R * Lp * dpm);
            }
        }
    r.gamtop.resize(nzp); r.gambot.resize(nzp);
    for (int z = 0; z <= nz; ++z) { const double ph = zp[z] / r.extent[2]; r.gamtop[z] = 0.5 * r.extent[2] * (ph * ph - 1.0 / 3.0); }
    for (int z = 0; z <= nz; ++z) r.gambot[z] = r.gamtop[nz - z];
}

// ---- field_diagnostics.f90 ----------------------------------------------------------------------------------
static double trap_sum3(const Ref& r, const vec& a, const vec& b, const vec& c, int mode) {
    // mode 0: a^2+b^2+c^2; 1: a (b, c ignored)
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (long long col = 0; col < (long long)r.NC; ++col) {
        const size_t o = (size_t)col * r.nzp;
        double t = 0.0;
        for (int z = 0; z <= r.nz; ++z) {
            const double w = (z == 0 || z == r.nz) ? 0.5 : 1.0;
            const double v = mode == 0 ? a[o + z] * a[o + z] + b[o + z] * b[o + z] + c[o + z] * c[o + z] : a[o + z];
            t += w * v;
        }
        s += t;
    }
    return s;
}
static void calc_vorticity_mean(const Ref& r, double savg[2]) {                 // field_diagnostics.f90:584-599
    ZWork w(r.nz);
    for (int nc = 0; nc < 2; ++nc) {
        vec wk(r.nz + 1, 0.0);
        for (int z = 1; z < r.nz; ++z) wk[z] = r.svor[nc][z];                   // column (0,0)
        dst_col(r, wk.data() + 1, w.X.data(), w.z.data(), w.y.data());
        double s = 0.0;
        for (int z = 1; z < r.nz; ++z) s += wk[z];
        savg[nc] = 0.5 * (r.svor[nc][0] + r.svor[nc][r.nz]) + r.fnzi * s;
    }
}
static void adjust_vorticity_mean(Ref& r) {                                     // :604-619
    double savg[2];
    calc_vorticity_mean(r, savg);
    for (int nc = 0; nc < 2; ++nc) { r.svor[nc][0] += r.ini_mean[nc] - savg[nc]; r.svor[nc][r.nz] += r.ini_mean[nc] - savg[nc]; }
}

// ---- inversion.f90:23-226 -----------------------------------------------------------------------------------
static void vor2vel(Ref& r) {
    const int nz = r.nz, nzp = r.nzp;
    const size_t N = r.N;
    vec as(N), bs(N), cs(N), ds(N), es(N), t1(N), t2(N);
    diffx(r, r.svor[1].data(), as.data());
    diffy(r, r.svor[0].data(), bs.data());
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)N; ++i) ds[i] = as[i] - bs[i];
    cs = r.svor[2];
    combine_semi_spectral(r, cs.data());
    central_diffz(r, cs.data(), es.data());
    decompose_semi_spectral(r, es.data());
    vec ubar(r.svor[0].begin(), r.svor[0].begin() + nzp), vbar(r.svor[1].begin(), r.svor[1].begin() + nzp);
    diffx(r, es.data(), t1.data()); diffy(r, ds.data(), t2.data());
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)N; ++i) r.svor[0][i] = r.k2l2i[i / nzp] * (t1[i] + t2[i]);
    diffy(r, es.data(), t1.data()); diffx(r, ds.data(), t2.data());
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)N; ++i) r.svor[1][i] = r.k2l2i[i / nzp] * (t1[i] - t2[i]);
    std::copy(ubar.begin(), ubar.end(), r.svor[0].begin());
    std::copy(vbar.begin(), vbar.end(), r.svor[1].begin());
    for (int nc = 0; nc < 3; ++nc) combine_physical(r, r.svor[nc].data(), r.vor[nc].data());
    diffy(r, r.svor[0].data(), t1.data()); diffx(r, r.svor[1].data(), t2.data());
#pragma omp parallel for schedule(static)
    for (long long c = 0; c < (long long)r.NC; ++c) {
        const size_t o = (size_t)c * nzp;
        for (int z = 0; z <= nz; ++z) ds[o + z] = t1[o + z] - t2[o + z];
        const double d0 = ds[o], dn = ds[o + nz];
        for (int z = 0; z <= nz; ++z) {
            bs[o + z] = (z >= 1 && z < nz) ? d0 * r.thetam[o + z] + dn * r.thetap[o + z] : 0.0;
            es[o + z] = d0 * r.dthetam[o + z] + dn * r.dthetap[o + z];
        }
        for (int z = 1; z < nz; ++z) ds[o + z] *= r.green[o + z];
        for (int z = 0; z <= nz; ++z) as[o + z] = (z >= 1 && z < nz) ? r.rkz[z] * ds[o + z] : 0.0;
    }
    fftcosine(r, as.data());
    fftsine(r, ds.data());
#pragma omp parallel for schedule(static)
    for (long long c = 0; c < (long long)r.NC; ++c) {
        const size_t o = (size_t)c * nzp;
        ds[o] = 0.0;
        for (int z = 1; z < nz; ++z) ds[o + z] += bs[o + z];
        ds[o + nz] = 0.0;
        for (int z = 0; z <= nz; ++z) es[o + z] += as[o + z];
    }
    cs = r.svor[2];
    combine_semi_spectral(r, cs.data());
    // horizontally averaged flow (:150-165)
    {
        ZWork w(nz);
        vec ub(nzp, 0.0), vb(nzp, 0.0);
        for (int z = 1; z < nz; ++z) { ub[z] = -r.svor[1][z] / r.rkz[z]; vb[z] = r.svor[0][z] / r.rkz[z]; }
        dct_col(r, ub.data(), w.X.data(), w.z.data(), w.y.data());
        dct_col(r, vb.data(), w.X.data(), w.z.data(), w.y.data());
        for (int z = 0; z <= nz; ++z) {
            ubar[z] = ub[z] + r.svor[1][nz] * r.gamtop[z] - r.svor[1][0] * r.gambot[z];
            vbar[z] = vb[z] - r.svor[0][nz] * r.gamtop[z] + r.svor[0][0] * r.gambot[z];
        }
    }
    diffx(r, es.data(), t1.data()); diffy(r, cs.data(), t2.data());
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)N; ++i) r.svel[0][i] = r.k2l2i[i / nzp] * (t1[i] + t2[i]);
    std::copy(ubar.begin(), ubar.end(), r.svel[0].begin());
    fftxys2p(r, r.svel[0].data(), r.vel[0].data());
    diffy(r, es.data(), t1.data()); diffx(r, cs.data(), t2.data());
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)N; ++i) r.svel[1][i] = r.k2l2i[i / nzp] * (t1[i] - t2[i]);
    std::copy(vbar.begin(), vbar.end(), r.svel[1].begin());
    fftxys2p(r, r.svel[1].data(), r.vel[1].data());
    r.svel[2] = ds;
    fftxys2p(r, ds.data(), r.vel[2].data());
}

// ---- inversion.f90:298-371 ----------------------------------------------------------------------------------
static void vorticity_tendency(Ref& r) {
    const size_t N = r.N;
    vec fp(N), rr(N), q(N), p(N), t1(N), t2(N);
    const vec &u = r.vel[0], &v = r.vel[1], &w = r.vel[2], &xi = r.vor[0], &eta = r.vor[1], &zeta = r.vor[2];
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)N; ++i) fp[i] = u[i] * eta[i] - v[i] * xi[i];
    decompose_physical(r, fp.data(), rr.data());
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)N; ++i) fp[i] = w[i] * xi[i] - u[i] * zeta[i];
    decompose_physical(r, fp.data(), q.data());
    diffy(r, rr.data(), t1.data());
    central_diffz(r, fp.data(), t2.data());
    decompose_physical(r, t2.data(), p.data());
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)N; ++i) r.svorts[0][i] = t1[i] - p[i];
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)N; ++i) fp[i] = v[i] * zeta[i] - w[i] * eta[i];
    decompose_physical(r, fp.data(), p.data());
    diffx(r, rr.data(), t1.data());
    central_diffz(r, fp.data(), t2.data());
    decompose_physical(r, t2.data(), rr.data());
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)N; ++i) r.svorts[1][i] = rr[i] - t1[i];
    diffx(r, q.data(), t1.data()); diffy(r, p.data(), t2.data());
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)N; ++i) r.svorts[2][i] = t1[i] - t2[i];
}

// ---- jacobi.f90:19-51, 215-305: eigenvalues of a symmetric 3x3, cyclic Jacobi with Rutishauser rotations -------
static void givens(double aij, double di, double dj, double& s, double& t, double& tau) {
    const double eps = 2.220446049250313e-16;
    const double g = 100.0 * std::fabs(aij), h = dj - di;
    if (std::fabs(h) + g == std::fabs(h)) {
        t = aij / (h + std::copysign(eps, h));
    } else {
        const double theta = 0.5 * h / (aij + std::copysign(eps, aij));
        t = 1.0 / (std::fabs(theta) + std::sqrt(1.0 + theta * theta));
        if (theta < 0.0) t = -t;
    }
    const double c = 1.0 / std::sqrt(1.0 + t * t);
    s = t * c;
    tau = s / (1.0 + c);
}
static double max_abs_eig(double s11, double s12, double s13, double s22, double s23, double s33) {
    double a12 = s12, a13 = s13, a23 = s23, d[3] = {s11, s22, s33}, b[3] = {s11, s22, s33};
    double sm = std::fabs(a12) + std::fabs(a13) + std::fabs(a23);
    int sweeps = 0;
    while (sm > 1.0e-15 && sweeps++ < 100) {
        double zz[3] = {0.0, 0.0, 0.0}, s, t, tau, h, g, hh;
        givens(a12, d[0], d[1], s, t, tau); h = t * a12;
        zz[0] -= h; zz[1] += h; d[0] -= h; d[1] += h; a12 = 0.0;
        g = a13; hh = a23; a13 = g - s * (hh + g * tau); a23 = hh + s * (g - hh * tau);
        givens(a13, d[0], d[2], s, t, tau); h = t * a13;
        zz[0] -= h; zz[2] += h; d[0] -= h; d[2] += h; a13 = 0.0;
        g = a12; hh = a23; a12 = g - s * (hh + g * tau); a23 = hh + s * (g - hh * tau);
        givens(a23, d[1], d[2], s, t, tau); h = t * a23;
        zz[1] -= h; zz[2] += h; d[1] -= h; d[2] += h; a23 = 0.0;
        g = a12; hh = a13; a12 = g - s * (hh + g * tau); a13 = hh + s * (g - hh * tau);
        for (int k = 0; k < 3; ++k) { b[k] += zz[k]; d[k] = b[k]; }
        sm = std::fabs(a12) + std::fabs(a13) + std::fabs(a23);
    }
    return std::max(std::fabs(d[0]), std::max(std::fabs(d[1]), std::fabs(d[2])));
}

// ---- advance.f90:109-410 adapt (pretype 'vorch') ----------------------------------------------------------------
static double adapt(Ref& r, double t, double t_limit, double alpha, int stepper) {
    const size_t N = r.N;
    const int nz = r.nz, nzp = r.nzp;
    const double small = 1.0e-12, cflmax = 0.8;
    const vec &xi = r.vor[0], &eta = r.vor[1], &zeta = r.vor[2];
    const double vortrms = std::sqrt(trap_sum3(r, xi, eta, zeta, 0) / ((double)r.nx * r.ny * r.nz));
    // get_char_vorticity (field_diagnostics.f90:501-545)
    double l1 = 0.0, l2 = 0.0;
#pragma omp parallel for reduction(+ : l1, l2) schedule(static)
    for (long long c = 0; c < (long long)r.NC; ++c) {
        const size_t o = (size_t)c * nzp;
        for (int z = 1; z <= nz; ++z) {
            const double v1 = 0.5 * std::fabs(xi[o + z - 1] + xi[o + z]), v2 = 0.5 * std::fabs(eta[o + z - 1] + eta[o + z]);
            const double v3 = 0.5 * std::fabs(zeta[o + z - 1] + zeta[o + z]);
            if (v1 + v2 + v3 > vortrms) { l1 += v1 + v2 + v3; l2 += v1 * v1 + v2 * v2 + v3 * v3; }
        }
    }
    r.vorch = l2 / (small + l1);
    // velocity strain (advance.f90:199-276)
    vec dudx(N), dudy(N), dwdx(N), dvdy(N), dwdy(N), ts(N);
    diffx(r, r.svel[0].data(), ts.data()); fftxys2p(r, ts.data(), dudx.data());
    diffy(r, r.svel[0].data(), ts.data()); fftxys2p(r, ts.data(), dudy.data());
    diffx(r, r.svel[2].data(), ts.data()); fftxys2p(r, ts.data(), dwdx.data());
    diffy(r, r.svel[1].data(), ts.data()); fftxys2p(r, ts.data(), dvdy.data());
    diffy(r, r.svel[2].data(), ts.data()); fftxys2p(r, ts.data(), dwdy.data());
    double gg = 2.220446049250313e-16, umax = -1e300, vmax = -1e300, wmax = -1e300;
#pragma omp parallel for reduction(max : gg, umax, vmax, wmax) schedule(static)
    for (long long i = 0; i < (long long)N; ++i) {
        const double e = max_abs_eig(dudx[i], dudy[i] + 0.5 * zeta[i], dwdx[i] + 0.5 * eta[i], dvdy[i], dwdy[i] - 0.5 * xi[i],
                                     -(dudx[i] + dvdy[i]));
        gg = std::max(gg, e);
        umax = std::max(umax, r.vel[0][i]); vmax = std::max(vmax, r.vel[1][i]); wmax = std::max(wmax, r.vel[2][i]);
    }
    r.ggmax = gg;
    const double dtcfl = cflmax * std::min(r.dx[0] / (umax + small), std::min(r.dx[1] / (vmax + small), r.dx[2] / (wmax + small)));
    const double bfmax = 0.0;
    const double dt = std::min(std::min(alpha / (gg + small), alpha / (bfmax + small)), std::min(dtcfl, t_limit - t));
    if (stepper == 0) {
        // cn2_set_diffusion (cn2.f90:40-79) with the 'vorch' prefactor
        const double dfac = (r.nnu == 1) ? dt : r.vorch * dt;
        for (size_t c = 0; c < r.NC; ++c) r.vdiss[c] = 1.0 / (1.0 + dfac * r.vhdis[c]);
    } else {
        // impl_rk4_set_diffusion (impl_rk4.f90:37-55)
        for (size_t c = 0; c < r.NC; ++c) r.vdiss[c] = 0.5 * r.vorch * dt * r.vhdis[c];
    }
    return dt;
}

// ---- cn2.f90:92-181 with the literal combine -> vdiss -> decompose pairs ---------------------------------------
static void cn2_update(Ref& r, double dt2) {
    const int nzp = r.nzp;
    vec q(r.N);
    for (int nc = 0; nc < 3; ++nc) {
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)r.N; ++i) q[i] = r.filt[i] * (r.vortsm[nc][i] + dt2 * r.svorts[nc][i]);
        combine_semi_spectral(r, q.data());
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)r.N; ++i) q[i] *= r.vdiss[i / nzp];
        decompose_semi_spectral(r, q.data());
        r.svor[nc] = q;
    }
    adjust_vorticity_mean(r);
}
// impl_rk4.f90:212-364: q <- decompose(fac(ky,kx) * combine(q)), literally
static void cmd(const Ref& r, vec& q, const vec& fac) {
    combine_semi_spectral(r, q.data());
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)r.N; ++i) q[i] *= fac[i / r.nzp];
    decompose_semi_spectral(r, q.data());
}
static void rk4_step(Ref& r, double* t, double dt) {                           // impl_rk4.f90:76-207
    const double dt2 = 0.5 * dt, dt3 = dt / 3.0, dt6 = dt / 6.0;
    const size_t N = r.N;
    vec epq(r.NC), emq(r.NC), q(N);
    for (size_t c = 0; c < r.NC; ++c) { const double e = std::exp(r.vdiss[c]); emq[c] = 1.0 / e; epq[c] = e * r.filt[c * r.nzp]; }
    vec svori[3], svorf[3];
    for (int nc = 0; nc < 3; ++nc) {                                           // substep one
        svori[nc] = r.svor[nc]; svorf[nc].resize(N);
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)N; ++i) {
            r.svorts[nc][i] *= r.filt[(i / r.nzp) * r.nzp];
            q[i] = svori[nc][i] + dt2 * r.svorts[nc][i];
            svorf[nc][i] = svori[nc][i] + dt6 * r.svorts[nc][i];
        }
        cmd(r, q, emq);
        r.svor[nc] = q;
    }
    vor2vel(r); vorticity_tendency(r);
    *t += dt2;
    auto substep = [&](double c1, double c2, bool last) {
        for (int nc = 0; nc < 3; ++nc) {
            cmd(r, r.svorts[nc], epq);
            const vec& base = last ? svorf[nc] : svori[nc];
#pragma omp parallel for schedule(static)
            for (long long i = 0; i < (long long)N; ++i) {
                q[i] = base[i] + c1 * r.svorts[nc][i];
                if (!last) svorf[nc][i] += c2 * r.svorts[nc][i];
            }
            cmd(r, q, emq);
            r.svor[nc] = q;
        }
    };
    substep(dt2, dt3, false);                                                  // substep two
    vor2vel(r); vorticity_tendency(r);
    *t += dt2;
    for (size_t c = 0; c < r.NC; ++c) emq[c] *= emq[c];
    substep(dt, dt3, false);                                                   // substep three
    vor2vel(r); vorticity_tendency(r);
    for (size_t c = 0; c < r.NC; ++c) epq[c] *= epq[c];
    substep(dt6, 0.0, true);                                                   // substep four
    adjust_vorticity_mean(r);
}
static double advance(Ref& r, double* t, double t_limit, double alpha, int stepper) {      // advance.f90:77-104
    vor2vel(r);
    const double dt = adapt(r, *t, t_limit, alpha, stepper);
    vorticity_tendency(r);
    if (stepper != 0) { rk4_step(r, t, dt); return dt; }
    const double dt2 = 0.5 * dt;
    for (int nc = 0; nc < 3; ++nc)
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)r.N; ++i) r.vortsm[nc][i] = r.svor[nc][i] + dt2 * r.svorts[nc][i];
    cn2_update(r, dt2);
    for (int iter = 0; iter < 2; ++iter) {
        vor2vel(r);
        vorticity_tendency(r);
        cn2_update(r, dt2);
    }
    *t += dt;
    return dt;
}

}  // namespace

// ---- C ABI for the tests and bench.py ----------------------------------------------------------------------------
extern "C" {

void* ps3d_ref_create(int nx, int ny, int nz, const double* lower, const double* extent, int filtering) {
    auto pow2 = [](int n) { return n >= 8 && (n & (n - 1)) == 0; };
    if (!pow2(nx) || !pow2(ny) || !pow2(nz)) return nullptr;
    Ref* r = new Ref();
    r->nx = nx; r->ny = ny; r->nz = nz; r->nzp = nz + 1;
    r->N = (size_t)nx * ny * (nz + 1); r->NC = (size_t)nx * ny;
    for (int i = 0; i < 3; ++i) { r->lower[i] = lower[i]; r->extent[i] = extent[i]; r->upper[i] = lower[i] + extent[i]; }
    r->dx[0] = extent[0] / nx; r->dx[1] = extent[1] / ny; r->dx[2] = extent[2] / nz;
    r->dzi = 1.0 / r->dx[2]; r->hdzi = 0.5 / r->dx[2];
    r->ncelli = 1.0 / ((double)nx * ny * nz); r->fnzi = 1.0 / nz;
    r->filtering = filtering;
    r->fx.init(nx); r->fy.init(ny); r->fz.init(nz);
    init_tables(*r);
    for (int c = 0; c < 3; ++c)
        for (vec* f : {&r->svor[c], &r->vor[c], &r->vel[c], &r->svel[c], &r->svorts[c], &r->vortsm[c]}) f->assign(r->N, 0.0);
    r->vhdis.assign(r->NC, 0.0); r->vdiss.assign(r->NC, 0.0);
    return r;
}
void ps3d_ref_destroy(void* h) { delete static_cast<Ref*>(h); }

// setup_fields (utils.f90:136-184) + init_diffusion (inversion_utils.f90:124-218, 'Kolmogorov'); returns ke, en
void ps3d_ref_set_vorticity(void* h, const double* vor, int nnu, double prediss, double* ke_en) {
    Ref& r = *static_cast<Ref*>(h);
    for (int c = 0; c < 3; ++c) {
        std::copy(vor + (size_t)c * r.N, vor + (size_t)(c + 1) * r.N, r.vor[c].begin());
        decompose_physical(r, r.vor[c].data(), r.svor[c].data());
    }
    calc_vorticity_mean(r, r.ini_mean);
    vor2vel(r);
    const double ke = 0.5 * trap_sum3(r, r.vel[0], r.vel[1], r.vel[2], 0) * r.ncelli;
    const double en = 0.5 * trap_sum3(r, r.vor[0], r.vor[1], r.vor[2], 0) * r.ncelli;
    const double kmax = std::max(*std::max_element(r.rkx.begin(), r.rkx.end()), *std::max_element(r.rky.begin(), r.rky.end()));
    const double K2max = kmax * kmax;
    const double vis = prediss * std::cbrt(K2max * ke / en) * std::pow(1.0 / K2max, nnu);
    r.nnu = nnu;
    for (size_t c = 0; c < r.NC; ++c) r.vhdis[c] = (nnu == 1) ? vis * r.k2l2[c] : vis * std::pow(r.k2l2[c], nnu);
    if (ke_en) { ke_en[0] = ke; ke_en[1] = en; }
}
void ps3d_ref_vor2vel(void* h) { vor2vel(*static_cast<Ref*>(h)); }
// stepper: 0 cn2, 1 impl-diff-rk4
double ps3d_ref_advance(void* h, double* t, double t_limit, double alpha, int stepper) {
    return advance(*static_cast<Ref*>(h), t, t_limit, alpha, stepper);
}

// field: 0 svor, 1 vor, 2 vel, 3 svel, 4 svorts  -> out[3][nx][ny][nz+1]
void ps3d_ref_get(void* h, int field, double* out) {
    Ref& r = *static_cast<Ref*>(h);
    vec* f = field == 0 ? r.svor : field == 1 ? r.vor : field == 2 ? r.vel : field == 3 ? r.svel : r.svorts;
    for (int c = 0; c < 3; ++c) std::copy(f[c].begin(), f[c].end(), out + (size_t)c * r.N);
}
// operators for the cross-check: 0 fftxyp2s, 1 fftxys2p, 2 fftsine, 3 fftcosine, 4 diffx, 5 diffy, 6 central_diffz,
// 7 combine_semi_spectral, 8 decompose_semi_spectral
void ps3d_ref_op(void* h, int op, const double* in, double* out) {
    Ref& r = *static_cast<Ref*>(h);
    switch (op) {
        case 0: fftxyp2s(r, in, out); break;
        case 1: fftxys2p(r, in, out); break;
        case 2: std::copy(in, in + r.N, out); fftsine(r, out); break;
        case 3: std::copy(in, in + r.N, out); fftcosine(r, out); break;
        case 4: diffx(r, in, out); break;
        case 5: diffy(r, in, out); break;
        case 6: central_diffz(r, in, out); break;
        case 7: std::copy(in, in + r.N, out); combine_semi_spectral(r, out); break;
        case 8: std::copy(in, in + r.N, out); decompose_semi_spectral(r, out); break;
        default: break;
    }
}
// number of OpenMP threads of the following calls (n <= 0: leave as is); returns the number in effect
int ps3d_ref_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}
void ps3d_ref_diag(void* h, double* out) {   // vorch, ggmax of the last adapt
    Ref& r = *static_cast<Ref*>(h);
    out[0] = r.vorch; out[1] = r.ggmax;
}

}  // extern "C"
