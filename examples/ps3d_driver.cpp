// Stand-in for the host program of the reference (src/ps3d.f90:31-189) over the C ABI: the Fortran
// toolchain is absent from this image, so this C++ program plays pre_run / run / post_run for the
// Beltrami configurations (examples/beltrami_<n>.config + beltrami<n>x<n>x<n>.nml).  It contains no
// numerics of its own beyond the analytic initial condition (src/beltrami.f90:162-181).
//
//   g++ -O2 -std=c++17 examples/ps3d_driver.cpp -Iinclude -Lps3d_b200 -lps3d_cuda -Wl,-rpath,$PWD/ps3d_b200 -o examples/ps3d_driver
//   examples/ps3d_driver --n 64 --stepper cn2 --limit 1.0 [--steps 20]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ps3d_cuda.h"

static void check(int ierr, const char* what) {                 // mpi_exit_on_error (mpi_utils.f90:15-31)
    if (ierr != 0) {
        std::fprintf(stderr, "Error in %s: %s (status %d)\n", what, ps3d_cuda_last_error(), ierr);
        std::exit(1);
    }
}

int main(int argc, char** argv) {
    int n = 32, max_steps = -1;
    double limit = 100.0, alpha = 0.1, prediss = 30.0;          // examples/beltrami_32.config
    std::string stepper = "cn2";
    for (int i = 1; i < argc; ++i) {                            // parse_command_line (ps3d.f90:147-188)
        std::string a = argv[i];
        auto next = [&]() { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return argv[++i]; };
        if (a == "--n") n = std::atoi(next());
        else if (a == "--stepper") stepper = next();
        else if (a == "--limit") limit = std::atof(next());
        else if (a == "--steps") max_steps = std::atoi(next());
        else if (a == "--prediss") prediss = std::atof(next());
        else if (a == "--help") { std::puts("ps3d_driver --n <grid> --stepper cn2|impl-diff-rk4 --limit <t> [--steps <k>]"); return 0; }
    }
    const double pi = std::acos(-1.0);
    const double lower[3] = {-0.5 * pi, -0.5 * pi, -0.5 * pi}, extent[3] = {pi, pi, pi};   // beltrami32x32x32.nml
    check(ps3d_cuda_init(n, n, n, lower, extent, 0, 1, nullptr), "mpi_layout_init/initialise_fft");
    check(ps3d_cuda_init_inversion(PS3D_FILTER_HOU_LI), "init_inversion");

    // beltrami.f90:141-181, k = l = 2, m = 1
    const size_t N = (size_t)n * n * (n + 1);
    std::vector<double> vor(3 * N);
    const double kk = 2, ll = 2, mm = 1, al = std::sqrt(kk * kk + ll * ll + mm * mm), fk2l2 = al / (kk * kk + ll * ll);
    const double dx = pi / n;
    for (int ix = 0; ix < n; ++ix)
        for (int iy = 0; iy < n; ++iy)
            for (int iz = 0; iz <= n; ++iz) {
                const double x = lower[0] + dx * ix, y = lower[1] + dx * iy, z = lower[2] + dx * iz;
                const double cz = std::cos(mm * z), sz = std::sin(mm * z), s = std::sin(kk * x + ll * y), c = std::cos(kk * x + ll * y);
                const size_t i = ((size_t)ix * n + iy) * (n + 1) + iz;
                vor[i] = fk2l2 * (kk * mm * sz - ll * al * cz) * s;
                vor[N + i] = fk2l2 * (ll * mm * sz + kk * al * cz) * s;
                vor[2 * N + i] = al * cz * c;
            }
    // setup_fields (utils.f90:136-184)
    check(ps3d_cuda_upload_vorticity(vor.data()), "field_decompose_physical");
    check(ps3d_cuda_vor2vel(), "vor2vel");
    double d[8], nu = 0.0;
    check(ps3d_cuda_diagnostics(d), "diagnostics");
    check(ps3d_cuda_init_diffusion(3, prediss, PS3D_LSCALE_KOLMOGOROV, d[0], d[1], &nu), "init_diffusion");
    std::printf("Vorticity hyperviscosity nu = %14.7e\n", nu);   // inversion_utils.f90:208-210
    check(ps3d_cuda_stepper_setup(stepper == "cn2" ? PS3D_STEPPER_CN2 : PS3D_STEPPER_IMPL_RK4), "stepper setup");

    // run (ps3d.f90:107-126)
    double t = 0.0, dt = 0.0, diag[16];
    int steps = 0;
    const auto t0 = std::chrono::steady_clock::now();
    while (t < limit && (max_steps < 0 || steps < max_steps)) {
        const double tn = t;
        check(ps3d_cuda_advance(&t, limit, alpha, PS3D_PRE_VORCH, 1000, &dt, diag), "advance");
        std::printf(" At time %22.15e and time step %22.15e\n", tn, dt);        // advance.f90:98-100
        ++steps;
    }
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    check(ps3d_cuda_vor2vel(), "vor2vel");
    check(ps3d_cuda_diagnostics(d), "diagnostics");
    std::printf("steps %d  t %.15e  ke %.15e  en %.15e  helicity %.15e\n", steps, t, d[0], d[1], d[2]);
    std::printf("advance: %.3f s wall, %.3f ms/step, %.3e grid-pt*steps/s, %lld kernel launches\n", wall, 1e3 * wall / steps,
                (double)n * n * n * steps / wall, ps3d_cuda_kernel_launches());
    check(ps3d_cuda_finalise(), "finalise_inversion");
    return 0;
}
