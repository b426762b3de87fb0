/*
 * paramol_b200.h -- C ABI of the B200-native objective-evaluation path (libparamol_b200.so).
 *
 * The reference (JMorado/ParaMol, /root/reference) is pure Python and has no FFI of its own: the
 * hot path sits behind duck-typed Python seams that end in calls into the third-party OpenMM library.
 * Each entry point below names the reference interface it replaces (file:line under /root/reference).
 * INTEGRATION.md shows the ctypes stub a ParaMol maintainer would add to bind them.
 *
 * Conventions
 * -----------
 *  - every function returns 0 on success, non-zero on error; pm_last_error() gives the message
 *    (thread-local).  There is NO CPU fallback: without a CUDA device every compute call fails.
 *  - "dev" pointers are CUDA device pointers owned by the caller (the Python host layer allocates
 *    them as torch tensors); "host" pointers are ordinary host memory.  No torch types cross the ABI.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All launches are
 *    asynchronous on that stream unless the function name ends in _host.
 *  - units are OpenMM's: nm, kJ/mol, rad, elementary charge.
 *  - one handle per (topology, device); handles are not thread-safe, distinct handles are independent.
 *
 * Parameter "slots" (the physical MM parameter table; what an OpenMM Context holds after
 * OpenMMEngine.set_bonded_parameters / set_nonbonded_parameters, ParaMol/MM_engines/openmm.py:590-756),
 * one row of `pm_n_slots()` doubles per parameter vector:
 *    [ (r0, k)            x n_bonds      ]   HarmonicBondForce
 *    [ (theta0, k)        x n_angles     ]   HarmonicAngleForce
 *    [ (phase, k)         x n_torsions   ]   PeriodicTorsionForce (periodicity is topology)
 *    [ (q, sigma, eps)    x n_atoms      ]   NonbondedForce particles
 *    [ (qq, sigma, eps)   x n_exceptions ]   NonbondedForce exceptions
 *
 * Conformation "tiles": conformations are kept resident in HBM as tiles of CPW = pm_conf_per_tile(h) conformations,
 * tile t = doubles [3*n_atoms][CPW] with the conformation index fastest; a whole tile is one contiguous TMA bulk
 * copy into the shared memory of the warp that evaluates it.  CPW = 32 / (lanes per conformation) is chosen per
 * topology so that enough warps fit in shared memory (or fixed by the caller: pm_create's conf_per_tile).  pm_pack_tiles converts
 * from the reference's [S][N][3] C-order arrays (ParaMolSystem.ref_coordinates / ref_forces,
 * ParaMol/System/system.py:85-123).
 */
#ifndef PARAMOL_B200_H
#define PARAMOL_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PM_NPARTIAL 8 /* doubles per (parameter vector) row written by pm_reduce_objective */

typedef struct pm_handle_s *pm_handle;
typedef struct pm_comm_s *pm_comm; /* intra-node communicator of the conformation-sharded objective (see pm_comm_create) */

/* Topology: what app.AmberPrmtopFile(...).createSystem(...) hands to ParaMol
 * (ParaMol/MM_engines/openmm.py:161-254) minus the parameter values.  All arrays are host int32. */
typedef struct pm_topology {
    int n_atoms;
    int n_bonds;      const int *bond_idx;       /* [n_bonds][2]      */
    int n_angles;     const int *angle_idx;      /* [n_angles][3]     */
    int n_torsions;   const int *torsion_idx;    /* [n_torsions][4]   */
                      const int *torsion_per;    /* [n_torsions]      */
    int n_exceptions; const int *exception_idx;  /* [n_exceptions][2] */
                      const int *exception_active; /* [n_exceptions] 1: interacts with its own (qq,sigma,eps)
                                                      (scaled 1-4 pair); 0: pure exclusion            */
    int nonbonded_method;   /* 0 NoCutoff ; 1 CutoffNonPeriodic with reaction field (openmm.py:114) */
    double cutoff;          /* nm, method 1 only */
    double rf_dielectric;   /* method 1 only     */
    double one_4pi_eps0;    /* Coulomb constant, 138.935456 for the OpenMM 7.x the reference pins */
} pm_topology;

const char *pm_last_error(void);
int pm_version(void);
/* number of CUDA devices visible (0 = none; compute calls will fail) */
int pm_device_count(void);

/* ---- handle ------------------------------------------------------------------------------------ */
/* Replaces OpenMMEngine.init_openmm / Context creation (ParaMol/MM_engines/openmm.py:161-254).
 * conf_per_tile: conformations per tile (32, 16, 8 or 4: a conformation is evaluated by a team of 32 / conf_per_tile lanes of one
 * warp), -8: tiles of 32 conformations evaluated by teams of 8 WARPS (one lane per conformation, the warps share the tile), or 0 to
 * let the library choose per topology.  The
 * topology-specialised kernels (pm_set_specialized) need 32. */
int pm_create(const pm_topology *topo, int device, int conf_per_tile, pm_handle *out);
int pm_destroy(pm_handle h);
int pm_n_slots(pm_handle h);          /* doubles per parameter vector (layout above)           */
int pm_n_derived(pm_handle h);        /* doubles per parameter vector of the derived table      */
int pm_n_pairs(pm_handle h);          /* interacting pairs (plain pairs + active exceptions)    */
int pm_conf_per_tile(pm_handle h);    /* CPW: conformations per tile (32, 16, 8 or 4; chosen per topology) */
long pm_n_tiles(pm_handle h, long n_conf); /* ceil(n_conf / CPW)                                 */
long pm_tile_doubles(pm_handle h);    /* 3 * n_atoms * CPW                                       */
/* interacting pair list, bit-exact indexing contract: out [n_pairs][3] = (i, j, exception index or -1) */
int pm_get_pairs(pm_handle h, int *out_host);
/* launch geometry chosen for the E+F kernel (for the roofline bookkeeping in bench.py) */
int pm_launch_info(pm_handle h, int with_forces, int *grid, int *block, int *smem_bytes, int *warps_per_cta);
/* topology program statistics: out[10] = {CPW, lanes per conformation, stars, star steps, axes, axis steps,
 * pair segments, pair entries, pair entry slots (incl. padding), program ints} */
int pm_program_info(pm_handle h, int *out10);
/* host-only (no CUDA call): build the program for a given CPW and check its invariants -- every bond, angle, torsion
 * and interacting pair executed exactly once, records of one step atom-disjoint.  0 = valid. */
int pm_program_check(const pm_topology *topo, int cpw, int *stats10);

/* Per-atom 3x3 force metric M_j used by the force residual  sum_j dF_j^T M_j dF_j :
 * inverse covariance of the QM forces ("components", ParaMol/Objective_function/Properties/force_property.py:85-101,126-173)
 * or I / var_j ("norm", :102-117,175-202).  host [n_atoms][9]. */
int pm_set_force_metric(pm_handle h, const double *metric_host);

/* ---- conformations ---------------------------------------------------------------------------- */
/* [S][N][3] (dev, C order) -> tiles (dev, pm_n_tiles(h, S) * pm_tile_doubles(h) doubles); the last tile is
 * padded by repeating the last conformation.  Replaces the per-conformation context.setPositions of
 * ParaMol/System/system.py:333-338,349-354: coordinates are staged once and stay resident. */
int pm_pack_tiles(pm_handle h, const double *src_dev, long n_conf, double *tiles_dev, void *stream);
int pm_unpack_tiles(pm_handle h, const double *tiles_dev, long n_conf, double *dst_dev, void *stream);

/* ---- parameters -------------------------------------------------------------------------------- */
/* slots [P][pm_n_slots] (dev) -> derived [P][pm_n_derived] (dev): combining rules, c12/c6/Coulomb products
 * per interacting pair, cos/sin of torsion phases.  Replaces force.set*Parameters +
 * updateParametersInContext (ParaMol/MM_engines/openmm.py:507-588, 661-756). */
int pm_prepare_params(pm_handle h, const double *slots_dev, int n_vec, double *derived_dev, void *stream);

/* ---- evaluation -------------------------------------------------------------------------------- */
/* E (+F) of every conformation for every parameter vector.
 *   emm_dev   [P][S]            MM energies                 (ParaMolSystem.get_energies_ensemble, system.py:340-354)
 *   fres_dev  [P][S] or NULL    sum_j dF^T M_j dF per conformation with dF = F_mm - F_ref
 *                               (ForceProperty.calculate_property inner loops, force_property.py:85-117)
 *   fmm_tiles_dev [P][tiles] or NULL  MM forces in tile layout (get_forces_ensemble, system.py:324-338)
 * fref_tiles_dev may be NULL when fres_dev is NULL.  with_forces=0 runs the energy-only kernel. */
int pm_eval(pm_handle h, const double *derived_dev, int n_vec,
            const double *coord_tiles_dev, const double *fref_tiles_dev, long n_conf,
            int with_forces, double *emm_dev, double *fres_dev, double *fmm_tiles_dev, void *stream);

/* pm_eval with a term selection.  mode 0: all terms (= pm_eval).  mode 1: nonbonded terms only -- with emm_dev and
 * fmm_tiles_dev as outputs this builds the nonbonded cache (E_nb [S], F_nb tiles) of ONE parameter vector.  mode 2:
 * bonded terms evaluated, nonbonded energies/forces taken from the cache (nb_e_dev [S], nb_f_tiles_dev): for batches of
 * trial vectors that share their nonbonded parameters -- every bonded-only fit, and the reference's "everything
 * optimizable" setup in which set_nonbonded_parameters pushes nothing (ParaMol/MM_engines/openmm.py:752-754). */
int pm_eval_ex(pm_handle h, const double *derived_dev, int n_vec,
               const double *coord_tiles_dev, const double *fref_tiles_dev, long n_conf,
               int with_forces, double *emm_dev, double *fres_dev, double *fmm_tiles_dev,
               int mode, const double *nb_e_dev, const double *nb_f_tiles_dev, void *stream);

/* Weighted residual partial sums of this rank's conformations, one row of PM_NPARTIAL doubles per vector:
 *   [0] sum d'          d' = (E_ref - E_mm) - d_shift       (energy_property.py:99-108)
 *   [1] sum w d'   [2] sum w d'^2   [3] sum w   [4] sum w * fres   [5] count   [6],[7] reserved (0)
 * weights_dev NULL = w_s = 1.  Rows from different ranks add (one allreduce), then
 *   ENERGY = ([2] - 2 m [1] + m^2 [3]) / var(E_ref),  m = [0]/[5];   FORCE = [4] / (3 N).
 * scratch_dev: at least pm_reduce_scratch_doubles(P) doubles. */
long pm_reduce_scratch_doubles(int n_vec);
int pm_reduce_objective(const double *emm_dev, const double *fres_dev, const double *eref_dev,
                        const double *weights_dev, double d_shift, int n_vec, long n_conf,
                        double *partials_dev, double *scratch_dev, void *stream);

/* ---- fused evaluation + reduction (+ cross-GPU sum): the objective call as ONE pass ----------------------------- */
/* pm_eval followed by pm_reduce_objective without the [P][S] round trip: the E(+F) kernel adds the five weighted sums per warp
 * while it evaluates, a second small kernel adds the warps' rows in a fixed order (deterministic) and writes
 * partials_dev [P][PM_NPARTIAL] exactly as pm_reduce_objective does.  fref_tiles_dev != NULL (with_forces = 1): the force
 * residual enters column [4].  emm_dev / fres_dev [P][S]: optional per-conformation outputs (NULL = not written).
 * comm != NULL: the rows of ALL ranks of the communicator are summed inside that second kernel -- every rank pushes its 8
 * doubles per vector into its peers' inboxes over NVLink (peer-to-peer stores), waits for theirs and adds them in rank order, so
 * every rank holds the same global sums bit for bit and [5] is the global conformation count.  This is the single collective
 * of the conformation-sharded objective (the reference's counterpart: the fork pool's gather + np.sum of
 * ParaMol/Objective_function/objective_function.py:268-311).  scratch_dev: pm_eval_reduce_scratch_doubles(h, P, S) doubles. */
long pm_eval_reduce_scratch_doubles(pm_handle h, int n_vec, long n_conf);
int pm_eval_reduce(pm_handle h, const double *derived_dev, int n_vec, const double *coord_tiles_dev,
                   const double *fref_tiles_dev, long n_conf, int with_forces, const double *eref_dev,
                   const double *weights_dev, double d_shift, double *partials_dev, double *scratch_dev, pm_comm comm,
                   double *emm_dev, double *fres_dev, void *stream);

/* ---- energy-only batches: geometry evaluated once, reused by every parameter vector (K2) ----------------------------- */
/* E[p][s] = sum_t C[p][t] G[t][s]: every MM energy term is linear in coefficients C (from the derived tables of the n_vec vectors)
 * once the geometry features G (r^2, r, theta^2, theta, cos / sin(n phi), r^-12, r^-6, 1/r) of a conformation are known, so a batch
 * of trial vectors -- Monte Carlo / simulated annealing proposals, the evaluations of a finite-difference gradient, WHAM passes;
 * ParaMol/Optimizers/monte_carlo.py:89-101, gradient_descent.py:104-122 -- costs ONE pass over the geometry plus a dense FP64
 * contraction on the tensor-core path instead of n_vec full term passes.  Output as pm_eval_reduce with with_forces = 0:
 * partials_dev [n_vec][PM_NPARTIAL] ([4] = 0), summed over the ranks of `comm` when given; emm_dev [n_vec][n_conf] optional.
 * NoCutoff systems only (pm_energy_batch_supported).  scratch_dev: pm_energy_batch_scratch_doubles(h, n_vec, n_conf) doubles. */
int pm_energy_batch_supported(pm_handle h);
long pm_energy_batch_scratch_doubles(pm_handle h, int n_vec, long n_conf);
int pm_energy_batch_reduce(pm_handle h, const double *derived_dev, int n_vec, const double *coord_tiles_dev, long n_conf,
                           const double *eref_dev, const double *weights_dev, double d_shift, double *partials_dev,
                           double *scratch_dev, pm_comm comm, double *emm_dev, void *stream);

/* ---- intra-node communicator (one process per GPU, NVLink / NVSwitch peer memory) -------------------------------- */
/* pm_comm_create allocates this rank's inbox on `device` and returns its 64-byte CUDA IPC handle in handle_out; the host
 * layer exchanges the handles (torch.distributed all_gather) and passes all of them, in rank order, to pm_comm_connect, which
 * maps the peers' inboxes.  max_vec bounds n_vec of the calls that use the communicator.  All ranks must issue the same
 * sequence of communicating calls (SPMD); a peer that never arrives makes the kernel give up after ~6 s, poison the result with
 * NaN and count a timeout (pm_comm_timeouts) instead of hanging the GPU.
 * pm_comm_allreduce: the same one-shot sum for an arbitrary small buffer (n <= 8 * max_vec doubles), in place. */
int pm_comm_create(int device, int rank, int world, int max_vec, void *handle_out, pm_comm *out);
int pm_comm_connect(pm_comm c, const void *all_handles);
int pm_comm_allreduce(pm_comm c, double *buf_dev, int n, void *stream);
long pm_comm_timeouts(pm_comm c);
int pm_comm_destroy(pm_comm c);

/* ---- RESP (K5) ------------------------------------------------------------------------------------ */
/* Normal-equation sums over conformations, the inverse distances R_m[j][i] = 1/|x_mj - g_mi| being generated on the fly:
 *   out[0 .. N*N)      A = sum_m wt_m R_m R_m^T     (row-major, symmetric)
 *   out[N*N .. N*N+N)  B = sum_m wt_m R_m V_m
 *   out[N*N+N]         C = sum_m wt_m V_m . V_m
 * coords [S][N][3], grid [S][G][3], vref [S][G], wt [S] or NULL (= 1), all device f64, any consistent length unit.
 * Replaces RESP.calculate_inverse_distances / _calculate_a / _calculate_b (ParaMol/MM_engines/resp.py:99-134, 381-421,
 * 463-496); with wt_m = w_m the same sums give the ESP objective of ESPProperty (esp_property.py:89-104) for ANY charge
 * vector q as  C - 2 q.B + q^T A q.  scratch: pm_resp_scratch_doubles(N, S, G) doubles.  Acts on the current device. */
long pm_resp_scratch_doubles(int n_atoms, long n_conf, int n_grid);
int pm_resp_assemble(int n_atoms, const double *coords_dev, const double *grid_dev, const double *vref_dev,
                     const double *wt_dev, long n_conf, int n_grid, double *out_dev, double *scratch_dev, void *stream);
/* V_mm[m][i] = sum_j q_j / r_ij  (ParaMolSystem.get_esp_ensemble, ParaMol/System/system.py:356-382); charges [N]. */
int pm_esp_potential(int n_atoms, const double *coords_dev, const double *grid_dev, const double *charges_dev,
                     long n_conf, int n_grid, double *vmm_dev, void *stream);

/* ---- topology-specialised E+F kernel ------------------------------------------------------------------- */
/* pm_set_specialized installs a cubin image holding `pm_ef_spec_kernel` and `__constant__ double spec_tab[]`, generated for
 * exactly this topology by paramol_b200/csrc/specialize.py (the topology as straight-line code); pm_eval / pm_eval_ex then use
 * it for full evaluations (mode 0; `pm_e_spec_kernel` with warps_energy warps per CTA for with_forces = 0 when the image has it)
 * and the generic kernel for everything else.  Same inputs, same
 * outputs, same reference code replaced as pm_eval.  n_force_tile_atoms: how many atoms keep their force
 * accumulators in a (packed) shared-memory tile; the others are accumulated in registers (0 = all in registers).  warps_grad > 0: the image also holds `pm_grad_spec_kernel` / `pm_grad_spec_reduce`, which pm_grad then
 * uses.  image = NULL removes it.  Needs pm_conf_per_tile(h) == 32.
 * pm_topology_pairs (host only, no CUDA call): the pair list in derived-table order for a given conformations-per-tile value,
 * so that the source can be generated and compiled before a GPU is present. */
int pm_set_specialized(pm_handle h, const void *image, long n_bytes, int warps, int n_force_tile_atoms, int warps_energy,
                       int warps_grad);
int pm_has_specialized(pm_handle h);
int pm_topology_pairs(const pm_topology *topo, int cpw, int *n_pairs, int *out_pairs);

/* ---- analytic parameter gradient ---------------------------------------------------------------------- */
/* d f / d p for ONE parameter vector, where f = sum_s [ g(E_s) + b_s sum_a (F_sa - Fref_sa)^T M_a (F_sa - Fref_sa) ]:
 * the caller passes a_s = d f / d E_s (coef_e_dev [S], or NULL when the objective has no energy part) and b_s (coef_f_dev [S],
 * or NULL; needs the MM force tiles written by pm_eval(..., fmm_tiles_dev) and the reference force tiles).  The reference
 * has no counterpart: it differences ObjectiveFunction.f (scipy jac="2-point", ParaMol/Utils/settings.py:64;
 * ParaMol/Optimizers/gradient_descent.py:106-121); this replaces those P_opt + 1 evaluations of
 * ParaMol/Objective_function/objective_function.py:231-359 with one pass.
 * gout_dev [pm_grad_out_doubles(h)] = bonds (d/d r0, d/d k), angles (d/d theta0, d/d k), torsions (d/d phase, d/d k),
 * pairs in pm_get_pairs order (d/d c12, d/d c6, d/d (K qq)) with c12 = 4 eps sigma^12, c6 = 4 eps sigma^6 as in the derived
 * table of pm_prepare_params.  scratch_dev: pm_grad_scratch_doubles(h) doubles.  Deterministic (no atomics).
 * Fails for the cutoff / reaction-field variant. */
long pm_grad_out_doubles(pm_handle h);
long pm_grad_scratch_doubles(pm_handle h);
int pm_grad(pm_handle h, const double *derived_dev, const double *coord_tiles_dev, const double *fmm_tiles_dev,
            const double *fref_tiles_dev, long n_conf, const double *coef_e_dev, const double *coef_f_dev, double *gout_dev,
            double *scratch_dev, void *stream);

/* ---- delta evaluation ---------------------------------------------------------------------------------- */
/* Energies (and force residuals) of n_vec trial vectors that differ from a stored BASE vector in a few terms each, as the
 * single-parameter moves of the reference's Monte Carlo / simulated annealing do (ParaMol/Optimizers/monte_carlo.py:89-100,
 * simulated_annealing.py:90-131; the reference re-evaluates everything, Objective_function/objective_function.py:231-359).
 *   E_s(p) = e_base[s] + sum_affected (E_t(new) - E_t(old));   fres_s(p) = res_base[s] + the change caused by the force deltas.
 * Proposal p owns entries prop_ptr[p] .. prop_ptr[p+1]-1; an entry is 12 ints: (kind 0 bond / 1 angle / 2 torsion / 3 pair,
 * offset of the term in the derived table, periodicity, 0), the term's atoms (4), and for each of them its index into the
 * proposal's list of affected atoms atom_list[atom_ptr[p] ..] (at most max_loc <= pm_delta_max_local_atoms() of them).
 * derived_base_dev [n_derived] and derived_dev [n_vec][n_derived] come from pm_prepare_params.  fbase_tiles_dev = the MM
 * forces of the base vector in tile layout (NULL: energies only; then res_base_dev / fres_dev are unused).
 * commit = 1 (n_vec = 1): the proposal is folded into e_base, res_base and fbase_tiles instead (accepted move).
 * pm_derived_layout: out5 = offsets of bonds, angles, torsions, pairs in the derived table and the pair stride. */
int pm_delta_max_local_atoms(void);
int pm_derived_layout(pm_handle h, int *out5);
int pm_delta_eval(pm_handle h, const double *derived_base_dev, const double *derived_dev, int n_vec, const int *prop_ptr_dev,
                  const int *entries_dev, const int *atom_ptr_dev, const int *atom_list_dev, int max_loc,
                  const double *coord_tiles_dev, const double *fref_tiles_dev, double *fbase_tiles_dev, double *e_base_dev,
                  double *res_base_dev, long n_conf, double *emm_dev, double *fres_dev, int commit, void *stream);

/* ---- LLS (K3) ------------------------------------------------------------------------------------- */
/* Internal coordinates q[S][n_terms] with the conventions of ParaMol/Utils/geometry.py:4-107: kind 0 distance(a0,a1),
 * 1 angle at a1 (arccos of the clipped cosine), 2 dihedral a0-a1-a2-a3 (atan2 form).  atoms [n_terms][4] int32 device. */
int pm_lls_internal(int n_atoms, const double *coords_dev, long n_conf, int n_terms, const int *kind_dev,
                    const int *atoms_dev, double *q_dev, void *stream);
/* Design-matrix columns A[S][n_cols] (LinearLeastSquare._calculate_a, ParaMol/MM_engines/least_square.py:492-645):
 * column c reads term col_term[c]; col_kind 0/1: 1/2 (q - p0)^2 ; col_kind 2: 1 + cos(per q - p0). */
int pm_lls_columns(const double *q_dev, long n_conf, int n_terms, int n_cols, const int *col_term_dev,
                   const int *col_kind_dev, const double *col_p0_dev, const int *col_per_dev, double *a_dev, void *stream);

/* ---- LLS (K3 + K4 fused): normal equations straight from the resident conformation tiles ---------------------------- */
/* A plan holds the columns of LinearLeastSquare._calculate_a (ParaMol/MM_engines/least_square.py:469-649) as a term table: term t
 * is the internal coordinate of one optimizable force constant (kind 0 bond a0-a1, 1 angle at a1, 2 dihedral a0-a1-a2-a3 of
 * periodicity term_per[t], conventions of ParaMol/Utils/geometry.py) and owns term_ncol[t] = 1 or 2 consecutive columns (2: the
 * "_x"/"_y" expansion used when the equilibrium value / phase is optimizable too).  All arrays here are HOST int32; term_atoms is
 * [n_terms][4].  conf_per_tile: layout of the coordinates handed to the calls below -- the pm_pack_tiles layout of a pm_handle
 * (4, 8, 16, 32) or 1 = plain [S][N][3].  The columns are generated inside the kernels; the [S][M] matrix never exists. */
typedef struct pm_lls_s *pm_lls;
int pm_lls_create(int device, int n_atoms, int conf_per_tile, int n_terms, const int *term_kind, const int *term_atoms,
                  const int *term_per, const int *term_ncol, pm_lls *out);
int pm_lls_n_cols(pm_lls p);
/* Per column (HOST arrays [M]): p0 = r0 / theta0 / phase delta; mean (NULL = 0) is subtracted from the raw column
 * 1/2 (q - p0)^2 or 1 + cos(n phi - delta) (the centring of least_square.py:641-645); scale (NULL = 1) multiplies the centred
 * column (1 / scaling constant, :112-117). */
int pm_lls_set_columns(pm_lls p, const double *p0, const double *mean, const double *scale);
long pm_lls_scratch_doubles(pm_lls p, long n_conf);
/* y[j] = sum_s w_s e_s (a_sj - mean_j)  (e_dev, w_dev NULL = 1: with mean = 0 these are the column sums) and, when minmax_dev
 * != NULL, [n_terms][2] = min and max of the internal coordinate over the conformations (bonds and angles; the reference centres
 * the angle expansion at the observed extrema, :583-586). */
int pm_lls_colstats(pm_lls p, const double *coords_dev, long n_conf, const double *e_dev, const double *w_dev, double *y_dev,
                    double *minmax_dev, double *scratch_dev, void *stream);
/* e[s] = b[s] - sum_j (a_sj - mean_j) z[j]   (b_dev NULL = 0): the residual of a trial solution z in parameter units. */
int pm_lls_residual(pm_lls p, const double *coords_dev, long n_conf, const double *z_dev, const double *b_dev, double *e_dev,
                    void *stream);
/* g[(M+1)][(M+1)] = [A b]^T [A b] with A_sj = rowscale_s * scale_j * (a_sj - mean_j) and b_s = rowscale_s * b_dev[s] as column M
 * (rowscale_dev NULL = 1; b_dev NULL = 0): the weighted normal equations A^T W A, A^T W b and b^T W b of
 * least_square.py:104-135 in one pass on the FP64 tensor-core instruction.  Multi-GPU: add the g of the ranks (one all-reduce). */
int pm_lls_normal(pm_lls p, const double *coords_dev, long n_conf, const double *rowscale_dev, const double *b_dev, double *g_dev,
                  double *scratch_dev, void *stream);
int pm_lls_destroy(pm_lls p);

/* ---- single-vector objective call (the host side of ObjectiveFunction.f, ParaMol/Objective_function/objective_function.py:75-160) - */
/* An instantiated CUDA graph (a cudaGraphExec_t) captured over: pinned slots row -> device, pm_prepare_params, pm_eval_reduce,
 * partial sums -> pinned host.  pm_graph_launch copies in_host [n_in] into pin_in (the graph's pinned input) and launches the graph
 * on stream; pm_graph_wait waits for stream and copies pin_out [n_out] to out_host.  Between the two the caller is free to do its
 * host-only work (the regularisation term).  All pointers are HOST pointers. */
int pm_graph_launch(void *graph_exec, void *stream, const double *in_host, double *pin_in, long n_in);
int pm_graph_wait(void *stream, const double *pin_out, long n_out, double *out_host);

/* ---- measurement helpers ------------------------------------------------------------------------ */
/* Measured FMA-pipe peak of the current device: runs a register-only FMA kernel and returns FLOP/s
 * (2 flop per FMA) in *flops and the kernel time in *ms.  fp64 = 1: DFMA, 0: FFMA. */
int pm_measure_fma_peak(int device, int fp64, int iters, double *flops, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* PARAMOL_B200_H */
