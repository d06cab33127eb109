/* libps3d_cuda — C ABI of the B200 implementation of the ps3d time-step path.
 *
 * This is the drop-in boundary for the Fortran host program (matt-frey/ps3d):
 * each entry point replaces a public module procedure of the reference (cited
 * as file:line under /root/reference) and is meant to be bound from Fortran
 * with `bind(C)` interfaces inside thin replacement module bodies (see
 * INTEGRATION.md).  Conventions:
 *   - every function returns an int status (0 = PS3D_OK); the reference's
 *     `stop` / `mpi_stop` / `MPI_Abort` paths become non-zero statuses and the
 *     Fortran shim maps them to `mpi_exit_on_error` (mpi_utils.f90:15-31);
 *     `ps3d_cuda_last_error()` returns the message;
 *   - one host thread per rank, non-reentrant, one global context per process
 *     (the reference keeps all state in module globals: fields.f90:17-38,
 *     inversion_utils.f90:37-78, sta3dfft.f90:18-28);
 *   - host arrays are plain `double*` in the reference's own layout
 *     `f(0:nz, 0:ny-1, 0:nx-1)` (Fortran order, z fastest; fields.f90:59-65),
 *     i.e. C `f[x][y][z]` with nz+1 contiguous doubles per column;
 *   - the library fails (PS3D_ERR_NO_DEVICE) when no CUDA device is present:
 *     there is no CPU path.
 */
#ifndef PS3D_CUDA_H
#define PS3D_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    PS3D_OK = 0,
    PS3D_ERR_NOT_INITIALISED = 1,
    PS3D_ERR_BAD_ARGUMENT = 2,
    PS3D_ERR_UNSUPPORTED_SIZE = 3,   /* stafft.f90:87-95 "Factorisation not possible" analogue */
    PS3D_ERR_NO_DEVICE = 4,
    PS3D_ERR_DEVICE = 5,
    PS3D_ERR_UNSUPPORTED = 6
};

/* options.f90:67 `filtering` */
enum { PS3D_FILTER_HOU_LI = 0, PS3D_FILTER_23_RULE = 1 };
/* options.f90:61 `length_scale` */
enum { PS3D_LSCALE_KOLMOGOROV = 0, PS3D_LSCALE_GEOPHYSICAL = 1 };
/* options.f90:24 `stepper` */
enum { PS3D_STEPPER_CN2 = 0, PS3D_STEPPER_IMPL_RK4 = 1 };
/* options.f90:55 `pretype` (advance.f90:385-408) */
enum { PS3D_PRE_CONSTANT = 0, PS3D_PRE_VORCH = 1, PS3D_PRE_BFMAX = 2, PS3D_PRE_ROLL_MEAN_MAX_STRAIN = 3,
       PS3D_PRE_MAX_STRAIN = 4, PS3D_PRE_US_MAX_STRAIN = 5,
       PS3D_PRE_ROLL_MEAN_BFMAX = 6 /* buoyancy build only (advance.f90:395-398) */ };
/* resident fields (fields.f90:17-23) for ps3d_cuda_download / ps3d_cuda_upload */
enum { PS3D_F_SVOR = 0, PS3D_F_VOR = 1, PS3D_F_VEL = 2, PS3D_F_SVEL = 3, PS3D_F_SVORTS = 4,
       PS3D_F_PRES = 5, PS3D_F_DELTA = 6,
       /* buoyancy build (fields.f90:28-35): mixed-spectral b', physical b', mixed-spectral tendency; comp ignored */
       PS3D_F_SBUOY = 7, PS3D_F_BUOY = 8, PS3D_F_SBUOYS = 9 };
/* slots of the diag_out[16] array filled by ps3d_cuda_adapt / ps3d_cuda_advance
 * (advance.f90:188-193,315-321,366) */
enum { PS3D_D_VORTMAX = 0, PS3D_D_VORTRMS, PS3D_D_VORCH, PS3D_D_VORMEAN_X, PS3D_D_VORMEAN_Y, PS3D_D_VORMEAN_Z,
       PS3D_D_BFMAX, PS3D_D_GGMAX, PS3D_D_UMAX, PS3D_D_VMAX, PS3D_D_WMAX, PS3D_D_USGGMAX, PS3D_D_LSGGMAX,
       PS3D_D_RMV, PS3D_D_DT, PS3D_D_PREFACTOR };

#define PS3D_MAX_RANKS 8   /* one NVSwitch box: peer-memory tables are sized for it */

const char* ps3d_cuda_last_error(void);

/* mpi_layout_init (mpi_layout.f90:53) + update_parameters (parameters.f90:61) +
 * initialise_fft (sta3dfft.f90:53).  rank/nranks: slab decomposition over the
 * GPUs of one box (nranks <= PS3D_MAX_RANKS, else PS3D_ERR_BAD_ARGUMENT); nccl_id:
 * 128-byte ncclUniqueId shared by all ranks (NULL when nranks == 1).
 * Grid sizes: powers of two in 8..1024 per axis (tuned kernels, any nranks dividing nx and ny/2); other even
 * lengths 2^a 3^b 5^c -- what factorisen accepts, stafft.f90:128-187 -- with nx, ny <= 896, nz <= 1200
 * (coverage kernels); anything else returns PS3D_ERR_UNSUPPORTED_SIZE. */
int ps3d_cuda_init(int nx, int ny, int nz, const double lower[3], const double extent[3],
                   int rank, int nranks, const void* nccl_id);
/* init_inversion (inversion_utils.f90:222) */
int ps3d_cuda_init_inversion(int filtering_id);
/* init_diffusion (inversion_utils.f90:124); returns the (hyper)viscosity */
int ps3d_cuda_init_diffusion(int nnu, double prediss, int length_scale_id, double te, double en, double* nu_out);
/* finalise_inversion + finalise_fft (inversion_utils.f90:459, sta3dfft.f90:112) */
int ps3d_cuda_finalise(void);

/* ---- operator mode: host pointers in, host pointers out (H2D/D2H each call) ---- */
int ps3d_cuda_fftxyp2s(const double* fp, double* fs);                 /* sta3dfft.f90:136 */
int ps3d_cuda_fftxys2p(const double* fs, double* fp);                 /* sta3dfft.f90:202 */
int ps3d_cuda_fftsine(double* fs);                                    /* sta3dfft.f90:264 */
int ps3d_cuda_fftcosine(double* fs);                                  /* sta3dfft.f90:282 */
int ps3d_cuda_diffx(const double* fs, double* ds);                    /* sta3dfft.f90:304 */
int ps3d_cuda_diffy(const double* fs, double* ds);                    /* sta3dfft.f90:345 */
int ps3d_cuda_central_diffz(const double* fs, double* ds);            /* inversion_utils.f90:653 */
int ps3d_cuda_diffz(const double* fs, double* ds);                    /* inversion_utils.f90:683 (ENABLE_BUOYANCY) */
int ps3d_cuda_field_combine_semi_spectral(double* sf);                /* inversion_utils.f90:617 */
int ps3d_cuda_field_decompose_semi_spectral(double* sfc);             /* inversion_utils.f90:563 */
int ps3d_cuda_field_combine_physical(const double* sf, double* fc);   /* inversion_utils.f90:599 */
int ps3d_cuda_field_decompose_physical(const double* fc, double* sf); /* inversion_utils.f90:549 */

/* ---- resident mode: state lives in HBM (what the time loop uses) ---- */
/* setup_fields (utils.f90:160-165): vor(0:nz,y,x,1:3) -> decompose x3, ini_vor_mean */
int ps3d_cuda_upload_vorticity(const double* vor_phys);
/* Streamed form of ps3d_cuda_upload_vorticity for hosts that hand over a new state while the device is busy:
 * _begin queues the host -> device copies of the three components on a copy stream (vor_phys must stay valid and
 * should be pinned) and returns at once -- they overlap a running ps3d_cuda_advance of the previous state; _end
 * waits for them and decomposes (utils.f90:160-165).  The state is replaced only by _end.  Up to two uploads may be
 * queued (first in, first out; two staging sets), so that the copy of state k+1 can be started before _end of
 * state k and also overlaps its decomposition. */
int ps3d_cuda_upload_vorticity_begin(const double* vor_phys);
int ps3d_cuda_upload_vorticity_end(void);
int ps3d_cuda_vor2vel(void);                                          /* inversion.f90:23 */
int ps3d_cuda_source(void);                                           /* inversion.f90:378 */
/* adapt (advance.f90:109) incl. bstep%set_diffusion; pressure/divergence are lazy (see download) */
int ps3d_cuda_adapt(double t, double t_limit, double alpha, int pretype_id, int roll_mean_win_size,
                    double* dt, double diag_out[16]);
int ps3d_cuda_stepper_setup(int stepper_id);                          /* cn2.f90:83 / impl_rk4.f90:57 */
int ps3d_cuda_set_diffusion(double dt, double prefactor);             /* cn2.f90:40 / impl_rk4.f90:37 */
/* needs ps3d_cuda_source for the current state, as bstep%step does (advance.f90:95-102): PS3D_ERR_NOT_INITIALISED
 * when no source call happened since the last step */
int ps3d_cuda_step(double* t, double dt);                             /* cn2.f90:92 / impl_rk4.f90:76 */
/* advance (advance.f90:77-104) minus write_step */
int ps3d_cuda_advance(double* t, double t_limit, double alpha, int pretype_id, int roll_mean_win_size,
                      double* dt_out, double diag_out[16]);

/* ---- physics and the buoyancy build (configure.ac:228-245 --enable-buoyancy; a run-time switch here) ----
 * ps3d_cuda_set_physics: planetary vorticity f_cor(1:3) and squared buoyancy frequency bfsq of physics.f90:88-92,
 * 163-169 (the host evaluates 2 Omega cos/sin(lat) and reads / computes N^2 as physics.f90:300-340 does).  f_cor
 * enters the vorticity tendency (inversion.f90:310-314), the pressure of the buoyancy build (fields_derived.f90:111)
 * and the Rossby numbers of ps3d_cuda_field_stats.  Default: all zero. */
int ps3d_cuda_set_physics(const double f_cor[3], double bfsq);
/* Switches the ENABLE_BUOYANCY code paths on: allocates sbuoy, buoy, sbuoys and the stepper work fields
 * (fields.f90:28-35, cn2.f90:31 bsm, impl_rk4.f90:67-72).  After ps3d_cuda_init_inversion, before the first step.
 * From then on source = buoyancy_tendency + vorticity_tendency with r = u eta - v xi + b (inversion.f90:232-292,
 * 316-331, 378-388), adapt evaluates bfmax (advance.f90:147-168) and the steppers advance sbuoy (cn2.f90:107-117,
 * 151-160; impl_rk4.f90:91-105, 124-131, 154-164, 187-195). */
int ps3d_cuda_enable_buoyancy(void);
/* setup_fields (utils.f90:149-158) after the host removed the basic state: b'(0:nz,y,x) -> sbuoy */
int ps3d_cuda_upload_buoyancy(const double* buoy_phys);
/* the buoy_visc half of init_diffusion (inversion_utils.f90:144-151) and its options (options.f90:55-61):
 * bvisc, bhdis; the pretype / rolling-mean window adapt uses for bval (advance.f90:358-374) */
int ps3d_cuda_init_diffusion_buoyancy(int nnu, double prediss, int length_scale_id, double te, double en,
                                      int pretype_id, int roll_mean_win_size, double* nu_out);
/* the `bf` argument of bstep%set_diffusion (cn2.f90:61-75, impl_rk4.f90:44-50) for hosts that call
 * ps3d_cuda_set_diffusion themselves; ps3d_cuda_adapt / ps3d_cuda_advance do both */
int ps3d_cuda_set_diffusion_buoyancy(double dt, double bf);
/* out = bfmax, rmb (rolling mean of bfmax), bval (buoyancy prefactor), bvisc of the last adapt (advance.f90:168,
 * 358-374) */
int ps3d_cuda_buoyancy_diag(double out[4]);

/* field_netcdf.f90:216-230 reads these; comp = 0..2 (ignored for pres/delta, which are
 * computed on demand from the current state: fields_derived.f90:67,161) */
int ps3d_cuda_download(int field_id, int comp, double* host);
int ps3d_cuda_upload(int field_id, int comp, const double* host);
/* out[0..2] = kinetic energy, enstrophy (field_diagnostics.f90:85,172), helicity
 * (plotting/nc_reader.py:94-101 trapezoid mean of u.omega); out[3..7] = horizontal / vertical kinetic energy,
 * horizontal / vertical enstrophy, max horizontal enstrophy (field_diagnostics.f90:128,153,211,248,233) */
int ps3d_cuda_diagnostics(double out[8]);

/* The 40 scalars of the field-statistics file, update_netcdf_field_diagnostics
 * (field_diagnostics_netcdf.f90:257-439) together with the values `adapt` hands over through
 * set_netcdf_field_diagnostic (advance.f90:188-193, 315-321, 366): out[NC_x - 1] with the reference's indices
 * (field_diagnostics_netcdf.f90:36-75), see the PS3D_NC_* enumerators.  Assumes, like the reference, that the
 * fields are up to date (vor2vel done) and that adapt has run for the current state; evaluates the horizontal
 * divergence (fields_derived.f90:161-182) for NC_USDELRMS.  BFMAX is that of the last adapt (0 without buoyancy); RBFMAX and RIMIN
 * are 0 here (ps3d_cuda_buoyancy_diag returns the rolling mean; the Richardson number, APE and the other NC_APE..NC_MSS
 * statistics of the buoyancy build are host-side reductions over ps3d_cuda_download(PS3D_F_BUOY));
 * ROMIN / ROMAX are min / max zeta divided by f_cor(3) exactly as field_diagnostics.f90:293-320 (f_cor = 0 in the
 * configurations in scope: +-inf or nan, as in the reference). */
enum { PS3D_NC_KE = 0, PS3D_NC_EN, PS3D_NC_OMAX, PS3D_NC_ORMS, PS3D_NC_OCHAR, PS3D_NC_OXMEAN, PS3D_NC_OYMEAN,
       PS3D_NC_OZMEAN, PS3D_NC_KEXY, PS3D_NC_KEZ, PS3D_NC_ENXY, PS3D_NC_ENZ, PS3D_NC_OXMIN, PS3D_NC_OYMIN,
       PS3D_NC_OZMIN, PS3D_NC_OXMAX, PS3D_NC_OYMAX, PS3D_NC_OZMAX, PS3D_NC_HEMAX, PS3D_NC_GMAX, PS3D_NC_BFMAX,
       PS3D_NC_UMAX, PS3D_NC_VMAX, PS3D_NC_WMAX, PS3D_NC_USOXMAX, PS3D_NC_LSOXMAX, PS3D_NC_USOYMAX,
       PS3D_NC_LSOYMAX, PS3D_NC_USOZMAX, PS3D_NC_LSOZMAX, PS3D_NC_USUHMAX, PS3D_NC_USGMAX, PS3D_NC_LSGMAX,
       PS3D_NC_USZRMS, PS3D_NC_USDELRMS, PS3D_NC_RGMAX, PS3D_NC_RBFMAX, PS3D_NC_RIMIN, PS3D_NC_ROMIN,
       PS3D_NC_ROMAX, PS3D_NC_COUNT };
int ps3d_cuda_field_stats(double out[40]);

/* Kinetic-energy spectrum of the current velocity field, the computation of the post-processing program genspec
 * (genspec.f90:55-127): fftxyp2s of u, v, w, cosine transform of u, v and sine transform of w in z, shells
 * m = int(nint(|k|) / dk) with dk = kmax / sqrt((nx/2)^2 + (ny/2)^2 + nz^2), spec(m) = sum |u_k|^2 * 4/3 pi dk^3
 * ((m+1)^3 - m^3) / num(m), normalised so that sum(spec) dk = kinetic energy.  nbins = kmax + 1 values are written
 * to spec and num (bins with num = 0 are left 0 where the reference prints a warning); with spec == NULL only
 * *nbins and *dk are returned.  Needs vor2vel for the current state. */
int ps3d_cuda_genspec(int nmax, double* spec, double* num, int* nbins, double* dk);

/* ---- multi-rank transport ----
 * Slab decomposition over `nranks` GPUs of one box: physical fields are split in x (rank r owns planes
 * r*nx/P .. (r+1)*nx/P - 1), spectral fields in ky (rank r owns rows r*ny/P .. of the paired order
 * ky' = 0, ny/2, 1, ny-1, 2, ny-2, ...), one all-to-all of P equal contiguous blocks per 2-D FFT
 * (replaces transpose_to_pencil, fft_pencil.f90:283-330, and reverse_x/y, mpi_reverse.f90:332-398).
 * Default transport: NCCL (nccl_id given to ps3d_cuda_init).  Alternatively the host supplies its own
 * collectives, e.g. CUDA-aware MPI_Alltoall / MPI_Allreduce on the reference's communicators
 * (mpi_layout.f90:34-36): `alltoall` gets device pointers (block d of `send` goes to rank d and must land
 * as block <sender> of `recv`), `allreduce` a host buffer (op 0 = sum, 1 = max); both return 0 on success. */
typedef int (*ps3d_alltoall_fn)(const void* send, void* recv, size_t bytes_per_rank, void* user);
typedef int (*ps3d_allreduce_fn)(double* buf, int n, int op, void* user);
int ps3d_cuda_set_transport(ps3d_alltoall_fn alltoall, ps3d_allreduce_fn allreduce, void* user);
/* number of all-to-alls issued and bytes this rank sent to other ranks since init */
int ps3d_cuda_comm_stats(long long* n_alltoall, double* bytes_sent);

/* ---- introspection for benchmarks ---- */
/* number of kernels this library has launched since init */
long long ps3d_cuda_kernel_launches(void);
/* how many of them were TMA-staged line sweeps (line_tma.cuh; 0 when PS3D_LINE_TMA=0 or the driver refused the
 * tensor maps and the register-staged sweeps of line_fft.cuh ran instead) */
long long ps3d_cuda_tma_launches(void);
/* time the last ps3d_cuda_advance spent on the device (CUDA events), milliseconds */
double ps3d_cuda_last_advance_ms(void);
/* run `reps` back-to-back launches of one hot kernel on resident data and return the
 * average device time per launch in ms (CUDA events on the library's stream).
 * which: 0 = forward y sweep, 1 = forward x sweep, 2 = inverse x sweep,
 *        3 = inverse y sweep, 4 = vor2vel column kernel, 5 = source column kernel,
 *        6 = forward y sweep with the u x omega product in the load (four input fields),
 *        7 = forward y sweep storing straight into the peers' receive buffers over NVLink (nranks > 1 only;
 *            every rank must make the same call),
 *        8 = plain 16-byte-store copy of one field into the next rank's receive buffer, 9 = the same bytes in the
 *            exchange pattern (block d to rank d): the NVLink ceiling of SM-issued stores beside kernel 7 */
int ps3d_cuda_time_kernel(int which, int reps, double* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif
