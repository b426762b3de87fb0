// pm_internal.hpp -- definitions shared by the translation units of libparamol_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/paramol_b200.h"
#include "pm_program.hpp"

// =============================================================================================
// Device-visible description of one topology
// =============================================================================================
struct PmProg {
    int n_atoms, n_bonds, n_angles, n_torsions, n_pairs;
    int cpw, L;
    int n_star_steps, n_axis_steps, n_segs;
    int n_pair_pos, off_blocks;  // team mode: 4 x 4 pair blocks (pm_program.hpp)
    int n_int;      // ints in the blob
    int n_derived;  // doubles per parameter vector in the derived table
    int off_star_tab, off_star_hdr, off_star_rec, off_axis_tab, off_axis_hdr, off_axis_rec, off_axis_terms, off_seg_hdr, off_seg_tile,
        off_entries, off_tper;
    int doff_bond, doff_angle, doff_tors, doff_pair;
    int n_slots;
    int soff_bond, soff_angle, soff_tors, soff_part, soff_exc;
    double k_coul;
    double rf_krf, rf_crf, rf_cut2;  // reaction field (CutoffNonPeriodic); pair_stride == 4 when active
    int pair_stride;
    const int *ints;         // device blob
    const double *metric;    // device [9N]
    const int *pair_map;     // device [n_pairs][3] = (i, j, exception or -1)
    const double2 *acos_tab; // device [PM_ACOS_N + 1] = (cos, sin)(q pi / PM_ACOS_N)
    // cached nonbonded contribution (pm_eval_ex mode 2): forces in tile layout and energies [S]; null otherwise
    const double *nb_f_tiles;
    const double *nb_e;
};

// Fused residual reduction (pm_eval_reduce): when rows != nullptr the topology-interpreting E(+F) kernel adds, per warp, the five
// weighted sums of pm_reduce_objective over the conformations it evaluates and writes one row of 5 doubles per (vector, warp of
// the grid).  (The generated kernels keep their round-1 form -- per-conformation energy and residual out, summed by
// pm_reduce_stage1: any epilogue added to their 13k-instruction straight-line body cost 5-8 % of the kernel, measured.)
struct PmRed {
    const double *eref;  // [S] reference energies or null (= 0)
    const double *w;     // [S] weights or null (= 1)
    double d_shift;
    double *rows;        // [P][grid * warps][5] or null: no fused reduction
};

// =============================================================================================
// The object behind pm_handle
// =============================================================================================
struct pm_handle_s {
    int device;
    int n_sm;
    int max_smem;
    PmProg prog;
    PmHostProgram host;
    int *d_ints;
    double *d_metric;
    int *d_pair_map;
    double2 *d_acos;
    int warps[2];  // [with_forces]
    int team_warps;  // 0: a team is 32 / cpw lanes of one warp; 8: a team is 8 warps sharing a tile of 32 conformations
    int maxw[2];   // kernel variant (registers per thread) per [with_forces]
    int rf;        // reaction-field variant
    size_t smem_total[2];
    // raw bonded term lists (the gradient kernel walks flat term lists, pm_grad_kernels.cu)
    std::vector<int> bond_idx, angle_idx, torsion_idx, torsion_per;
    void *grad_state;               // owned by pm_grad_kernels.cu
    void (*grad_free)(void *);
    void *spec_state;               // owned by pm_spec.cu: a topology-specialised E+F kernel, or null
    void (*spec_free)(void *);
    void *egemm_state;              // owned by pm_egemm_kernels.cu: feature table of the energy-only batch contraction, or null
    void (*egemm_free)(void *);
};

int pm_fail_msg(const std::string &msg);  // sets pm_last_error(), returns 1 (pm_kernels.cu)

// pm_spec.cu
int pm_spec_metric_changed(pm_handle h);
int pm_spec_launch_info(pm_handle h, int with_forces, int *block, int *smem_bytes, int *warps);  // 1 when a specialised kernel serves that mode
int pm_spec_can_eval(pm_handle h, int with_forces);
int pm_spec_eval(pm_handle h, const double *derived_dev, int n_vec, const double *xt, const double *fref_t, long n_conf, double *emm,
                 double *fres, double *fmm_t, int with_forces, PmRed red, cudaStream_t stream);
int pm_spec_fused(pm_handle h);                                 // 1 when the generated kernels accumulate the residual sums themselves
long pm_spec_rows(pm_handle h, int with_forces, long n_conf);   // rows of five they write per vector

// pm_comm.cu
int pm_comm_max_vec(pm_comm c);
// rows [P][n_rows][5] -> partials [P][PM_NPARTIAL]; with a communicator the rows of all ranks are summed (one-shot all-reduce
// over NVLink peer memory inside the same kernel) and every rank gets the global sums
int pm_reduce_rows(const double *rows_dev, long n_rows, int n_vec, long n_conf, double *partials_dev, pm_comm comm, cudaStream_t stream);
int pm_spec_can_grad(pm_handle h);
long pm_spec_grad_scratch(pm_handle h);  // doubles
int pm_spec_grad(pm_handle h, const double *derived_dev, const double *xt, const double *fmm_t, const double *fref_t, long n_conf,
                 const double *coef_e, const double *coef_f, double *gout, double *scratch, cudaStream_t stream);
