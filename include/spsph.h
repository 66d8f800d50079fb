/*
 * spsph.h -- C-ABI of the B200-native Stress-Particle SPH time-step engine.
 *
 * The reference (CaitlinChalk/Stress-Particle-SPH) has no FFI; its seam is the three call
 * sites of PROGRAM SPH_2018:
 *     call Init_sph            code/1_SPH_2018.f90:132  -> spsph_create + spsph_upload
 *     call time_integration    code/1_SPH_2018.f90:173  -> spsph_step
 *     call OutputRes / out_prn code/1_SPH_2018.f90:156,183,188 -> spsph_download (+ host writers)
 * All arrays cross the boundary in the reference's own layout: Fortran column-major,
 * x(ndimn,ntotal2), vel(ndimn,ntotal2), stress(nstre,ntotal2), particle order
 * nodes 1..nnode, stress particles nnode+1..ntotal, dummy particles ntotal+1..ntotal2
 * (example_problems/.../3_SPH_material_2018.f90:961-1026, Setup_Global_Arrays).
 * Plain pointers and sizes only; no torch or CUDA types in any signature.
 */
#ifndef SPSPH_H
#define SPSPH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPSPH_MAX_TCURVES 8
#define SPSPH_MAX_TCURVE_PTS 128
#define SPSPH_MAX_BCS 16
#define SPSPH_NPROP 20
#define SPSPH_NINT_VARS 10 /* nint_Vars, 3_SPH_material_2018.f90:294 (Bui copy) */

/* Which copy of the reference sources the step must reproduce. */
enum {
  SPSPH_VARIANT_CODE = 0, /* code/                                             */
  SPSPH_VARIANT_BUI = 1,  /* example_problems/soil_failure_bui_et_al_2008      */
  SPSPH_VARIANT_VS = 2,   /* example_problems/vertical_slope                   */
  SPSPH_VARIANT_SL = 3    /* example_problems/strain_localisation_in_soil_sample */
};

/* Scalars gathered from input.txt / <name>.dat / <name>.pts by the host reader.
 * Every field cites the reference variable it mirrors. */
typedef struct spsph_params {
  int32_t struct_bytes; /* sizeof(spsph_params); checked by spsph_create */
  int32_t variant;      /* SPSPH_VARIANT_* */

  int32_t ndimn, nstre;                                   /* mat:88,99 (always 2, 4) */
  int32_t nnode, nstress, ntotal, ntotal2, ndummy;        /* mat:440-461,733,804 */
  int32_t ndummy2;                                        /* wall particles of the innermost layer, mat:794-802 */
  int32_t npoints;                                        /* mat:389,396 */
  int32_t sp_sph, inside_approach, sph_shift, vel_vector; /* mat:384-392 (logicals as 0/1) */
  int32_t shift_update, dummy_nodes;                      /* mat:392,423 */
  int32_t skf, sle, cspm, update_x, xsph;                 /* mat:304 */
  int32_t cont_density, art_stress;                       /* mat:311,384 */
  int32_t ntype_eco, ncrit, ntype_solid;                  /* props(1,1), props(1,2), mat:99 */
  int32_t no_bcs, ifsigman, ic_grav, tcurve_grav;         /* mat:184,323,329 */
  /* per-copy semantic switches */
  int32_t bc_loop_ntotal; /* Normal_BCs loop bound: 1 -> 1..ntotal (Bui, VS), 0 -> 1..nnode (code, SL) */
  float ae_threshold;     /* get_derivatives CSPM determinant cut: 1e-03 (code,VS,SL) / 1e-07 (Bui), main:599 */
  int32_t ntcurves;       /* 6_SPH_time_vars_2018.f90 */
  int32_t nptstcurves[SPSPH_MAX_TCURVES];

  double dx, dy, sml, r_x, r_y, disp_tol;    /* mat:413,417,467,628,392 */
  double alpha, beta, damping;               /* mat:318,315 */
  double ft_grav, cgrav[2];                  /* mat:329 */
  double props[SPSPH_NPROP];                 /* material 1, props(1,1:20), mat:151-158 */
  double D11, D22, D12, D33, D41, D42;       /* mat:744-749 */
  double xmin_domain[2], xmax_domain[2];     /* mat:290 */
  double pi;                                 /* (double)(4*atanf(1)), mat:163 */
  double ttcurves[SPSPH_MAX_TCURVES][SPSPH_MAX_TCURVE_PTS]; /* fp64 times  */
  float ftcurves[SPSPH_MAX_TCURVES][SPSPH_MAX_TCURVE_PTS];  /* fp32 factors (6_SPH_time_vars_2018.f90) */
  double bc_list[SPSPH_MAX_BCS][8]; /* bc_list(1:8, ibc), column ibc stored as a row here, mat:196-198 */
} spsph_params;

/* Full particle state in reference layout; used by upload and download.
 * Any pointer may be NULL in spsph_download (that array is skipped). */
typedef struct spsph_state {
  double *x;              /* (2,ntotal2) */
  double *vel;            /* (2,ntotal2) */
  double *stress;         /* (4,ntotal2) */
  double *rho;            /* (ntotal2)   */
  double *mass;           /* (ntotal2)   */
  double *hsml;           /* (ntotal2)   */
  int32_t *itype;         /* (ntotal2): 2 node, 1 stress particle, 25 dummy */
  double *internal_vars;  /* (10,ntotal) */
  double *f_drucker;      /* (ntotal)    */
  double *x00;            /* (2,ntotal2) initial positions for displ */
  double *displ;          /* (2,nnode)   */
  double *x_10;           /* (2,nnode)   */
  double *disp_10;        /* (nnode)     */
  float *wall_position;   /* (ntotal2), only dummy entries meaningful */
  float *horizontal_or_not; /* (ntotal2) */
  float *n_int;           /* (nnode) */
  int32_t *bc_int;        /* (nnode) */
  int32_t *if_out_domain; /* (ntotal2) */
  int32_t *bc_or_not;     /* (ntotal) */
  int32_t *bc_info;       /* (8,ntotal) */
} spsph_state;

typedef struct spsph_handle spsph_handle;

/* create: allocates the device-resident state for one simulation on CUDA device `device`. */
int spsph_create(spsph_handle **h, const spsph_params *p, int device);
/* upload: copies the host arrays in (never retains host pointers). == end of Init_sph. */
int spsph_upload(spsph_handle *h, const spsph_state *s);
/* step: one time_integration (2_SPH_main_2018.f90:78-184). The driver owns the clock:
 * itimestep_sph, time_sph (value *before* this step), dt_sph as in 1_SPH_2018.f90:172-174. */
int spsph_step(spsph_handle *h, int32_t itimestep_sph, double time_sph, double dt_sph);
/* run nsteps consecutive steps with the reference's clock update time_sph += dt_sph. */
int spsph_run(spsph_handle *h, int32_t first_itimestep, double time_sph, double dt_sph, int32_t nsteps,
              double *time_sph_out);
/* download: synchronises and copies state out in reference order. */
int spsph_download(spsph_handle *h, const spsph_state *s);
/* interaction statistics of the last neighbour search (grid_find_NEW, main:1405-1431). */
int spsph_pair_stats(spsph_handle *h, int64_t *npairs, int32_t *maxiac, int32_t *miniac, int32_t *noiac);
/* the last step's pair list in the reference's traversal order, after Pint_Update:
 * arrays of length *npairs (query with all pointers NULL first). ids are 1-based. */
int spsph_pairs(spsph_handle *h, int64_t *npairs, int32_t *pair_i, int32_t *pair_j, int32_t *pint_type,
                float *w, float *dwdx, float *dwdy);
/* device timing helpers for the bench: elapsed GPU ms of the last spsph_run (CUDA events on
 * the engine's stream) and number of kernels launched by it. */
int spsph_last_run_ms(spsph_handle *h, float *ms, int64_t *kernel_launches);
/* optional per-kernel timing: enable, run some steps, then read (total_ms, launches) for kid = 0,1,... until
 * the call returns non-zero. Times are CUDA-event intervals on the engine's stream. */
int spsph_profile(spsph_handle *h, int enable);
int spsph_profile_get(spsph_handle *h, int kid, const char **name, double *total_ms, int64_t *launches);
/* ---- multi-GPU x-slab decomposition: one process per GPU, every rank creates a handle on its own device and
 * uploads the COMPLETE problem, then calls spsph_dist_init with the same slab planes. planes[r] <= x < planes[r+1]
 * is rank r's slab (planes[0] = -inf, planes[nranks] = +inf). After that spsph_step exchanges ghost particles
 * and migrants with the neighbouring slabs over NCCL once per step; spsph_download returns this rank's local
 * view and spsph_dist_flags tells which entries are authoritative (1 owned, 2 ghost, 0 remote/stale).
 * id128: NCCL unique id from spsph_dist_unique_id on rank 0, distributed by the launcher (file, MPI, torch). */
/* The NCCL library is libnccl.so.2 from the loader's search path unless the environment variable SPSPH_NCCL_SO names
 * another file. SPSPH_PEEL=1 (read by spsph_dist_init) switches halo peeling on: the pair sums skip the ghost
 * particles that are too deep in the halo to matter for the remaining sweeps of a step; same results. */
int spsph_dist_unique_id(char *id128);
int spsph_dist_init(spsph_handle *h, int32_t rank, int32_t nranks, const char *id128, const double *planes,
                    int32_t halo_cells, int32_t halo_capacity);
int spsph_dist_flags(spsph_handle *h, int32_t *flags);
/* Dynamic re-slabbing (load balance as the material flows): new slab planes for the following steps, the
 * same array on every rank. The particles whose slab changes travel with the next halo exchange exactly like
 * migrants, so a plane may move by at most half the halo distance per call (call again after a step for more). */
int spsph_dist_set_planes(spsph_handle *h, const double *planes);
/* Row-wise transfer of the TIME-VARYING state (a slab rank moves its slab + halo only instead of the whole problem).
 * ids: n ascending 0-based particle numbers. Every non-NULL array of `s` holds the rows of those particles in that
 * order, with the row width of the full array: x 2, vel 2, stress 4, if_out_domain 1 (n rows); internal_vars 10,
 * f_drucker 1, bc_or_not 1 (rows of the ids below ntotal); displ 2, x_10 2, disp_10 1, n_int 1, bc_int 1 (rows of the
 * ids below nnode). The set-up arrays (mass, rho, hsml, itype, wall data, bc_info, x00) are not touched: they come
 * from the spsph_upload that started the run. spsph_upload_rows resets the pair-list length like spsph_upload; in a
 * multi-GPU run the uploaded rows become this rank's local particles (owned where their key position lies in the
 * slab, ghost inside the halo distance) and every other particle is remote, so ids must cover slab + halo
 * (spsph.dist.local_ids). The driver-side counterpart: each rank reads / writes only its own rows of the Fortran
 * arrays (1_SPH_2018.f90:132,156). */
int spsph_upload_rows(spsph_handle *h, const spsph_state *s, const int32_t *ids, int32_t n);
int spsph_download_rows(spsph_handle *h, const spsph_state *s, const int32_t *ids, int32_t n);
/* Output frame packed on the device (host formats): what OutputRes prints per particle
 * (3_SPH_material_2018.f90:2919-3060: ParaView rows "x, y, [vel], [stress], [strain], [disp_10], [density], [sml]"
 * selected by the *_out switches; GiD results :2930-3008 "disp", "vel", "sigma..", "plastic strain") gathered into one
 * row-major (count, ncols) table of doubles for the particles first .. first+count-1 (0-based reference numbering) and
 * copied out in a single transfer, instead of a spsph_download of every array. cols: ncols <= SPSPH_FRAME_MAX_COLS
 * column codes in the order the writer prints them. Columns that exist only for velocity particles (DISP10, DISPLX/Y)
 * read 0 for the others; wall particles carry x, y, RHO, HSML only. BC_OR_NOT (2 = node on the free surface, as
 * written to surface_points.csv, mat:3061-3078) runs get_nodes_on_free_surface first, like spsph_download. In a
 * multi-GPU run the rows of particles this rank does not own are stale, as in spsph_download (spsph_dist_flags). */
enum {
  SPSPH_COL_X = 0, SPSPH_COL_Y, SPSPH_COL_VX, SPSPH_COL_VY, SPSPH_COL_SXX, SPSPH_COL_SYY, SPSPH_COL_SXY,
  SPSPH_COL_SZZ, SPSPH_COL_EPSP, SPSPH_COL_DISP10, SPSPH_COL_RHO, SPSPH_COL_HSML, SPSPH_COL_DISPLX,
  SPSPH_COL_DISPLY, SPSPH_COL_FDRUCKER, SPSPH_COL_BC_OR_NOT, SPSPH_COL_COUNT
};
#define SPSPH_FRAME_MAX_COLS 16
int spsph_download_frame(spsph_handle *h, const int32_t *cols, int32_t ncols, int32_t first, int32_t count,
                         double *out);
/* particles per species (velocity, stress, wall) this rank processed in the last step: its slab + halo */
int spsph_local_counts(spsph_handle *h, int32_t *nloc3);
/* checkpoint support (the reference has none: SURVEY section 5). Everything a restart needs is in spsph_state except
 * the length the reference's pair list has grown to (grid_find_NEW never shrinks it, main:1210,1361-1383); it decides
 * in which order the next steps walk their pairs. spsph_upload resets it to 0 (a fresh run);
 * restoring the value read at the checkpoint makes the restarted run continue bit for bit. */
int spsph_get_list_capacity(spsph_handle *h, int64_t *m_pairs);
/* diagnostics: how many steps since spsph_create ran on the cell-tile kernels (acceptance masks + staged partner
 * tiles) and how many on the id-list kernels (steps in which the reference's pair list grows, and option
 * combinations the tile kernels do not cover). Both paths give bit-identical results. */
int spsph_path_counts(spsph_handle *h, int64_t *tile_steps, int64_t *list_steps);
int spsph_set_list_capacity(spsph_handle *h, int64_t m_pairs);
int spsph_sync(spsph_handle *h);
int spsph_destroy(spsph_handle *h);
const char *spsph_last_error(spsph_handle *h);
const char *spsph_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SPSPH_H */
