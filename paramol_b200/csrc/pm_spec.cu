// pm_spec.cu -- loading and launching a topology-SPECIALISED E+F kernel (see paramol_b200/csrc/specialize.py).
//
// The generic kernel interprets the topology as data; for a small molecule (both coordinate and force tile of 32
// conformations per warp fit ~10 times into shared memory) a kernel in which the topology is code -- immediate offsets,
// parameters as constant-bank operands, one large basic block per conformation -- needs ~9.0k instead of ~13.5k instructions
// per conformation (aniline) and is 1.36x faster.  The host layer generates the CUDA source, compiles it with nvcc to a cubin
// (cached) and hands the image to pm_set_specialized; pm_eval then uses it for full E+F evaluations and keeps the generic
// kernel for everything else (energy-only, cached-nonbonded modes, reaction field).
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "pm_device.cuh"
#include "pm_internal.hpp"

#define PMS_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) return pm_fail_msg(std::string(#call) + ": " + cudaGetErrorString(_e)); \
    } while (0)

struct PmSpecState {
    cudaLibrary_t lib;
    cudaKernel_t kernel;    // pm_ef_spec_kernel: energies, forces, residual
    cudaKernel_t kernel_e;  // pm_e_spec_kernel: energies only (null when the image has none)
    cudaKernel_t kernel_g;  // pm_grad_spec_kernel + pm_grad_spec_reduce: analytic parameter gradient (null when absent)
    cudaKernel_t kernel_gr;
    double *tab_dev;  // __constant__ spec_tab[n_derived + 9 N]: derived table of the vector being evaluated, then the metric
    int warps, warps_e, warps_g;
    int ngp;          // padded number of gradient quantities (rows of the partial array are ngp doubles)
    int fused;        // the E+F / energy kernels take (eref, weights, d_shift, rows) and accumulate the residual sums themselves
    size_t smem, smem_e, smem_g;
};

static void pm_spec_free(void *p) {
    PmSpecState *st = (PmSpecState *)p;
    if (!st) return;
    cudaLibraryUnload(st->lib);
    delete st;
}

extern "C" int pm_has_specialized(pm_handle h) { return h && h->spec_state ? 1 : 0; }

extern "C" int pm_set_specialized(pm_handle h, const void *image, long n_bytes, int warps, int n_force_tile_atoms, int warps_energy,
                                  int warps_grad) {
    if (!h) return pm_fail_msg("pm_set_specialized: null handle");
    PMS_CUDA(cudaSetDevice(h->device));
    if (h->spec_state) {
        pm_spec_free(h->spec_state);
        h->spec_state = nullptr;
    }
    if (!image || n_bytes <= 0) return 0;  // NULL image: back to the generic kernel
    if (h->rf || h->prog.cpw != 32) return pm_fail_msg("pm_set_specialized: needs a NoCutoff topology with 32 conformations per tile");
    if (warps < 1 || warps > 32 || n_force_tile_atoms < 0 || n_force_tile_atoms > h->prog.n_atoms)
        return pm_fail_msg("pm_set_specialized: bad warp count or force-tile size");
    PmSpecState *st = new PmSpecState();
    if (cudaLibraryLoadData(&st->lib, image, nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess) {
        delete st;
        return pm_fail_msg(std::string("pm_set_specialized: cudaLibraryLoadData: ") + cudaGetErrorString(cudaGetLastError()));
    }
    size_t bytes = 0;
    void *tab = nullptr;
    const size_t want = sizeof(double) * ((size_t)h->prog.doff_pair + 3 * (size_t)h->prog.n_pairs + 9 * (size_t)h->prog.n_atoms);
    if (cudaLibraryGetKernel(&st->kernel, st->lib, "pm_ef_spec_kernel") != cudaSuccess ||
        cudaLibraryGetGlobal(&tab, &bytes, st->lib, "spec_tab") != cudaSuccess || bytes != want) {
        cudaGetLastError();
        pm_spec_free(st);
        return pm_fail_msg("pm_set_specialized: the image does not hold pm_ef_spec_kernel / spec_tab for this topology");
    }
    st->tab_dev = (double *)tab;
    st->warps = warps;
    const size_t tile = (size_t)3 * h->prog.n_atoms * 32 * sizeof(double);
    // optional: extra tiles per warp of the E+F kernel (the reference-force buffer of the prefetching variant)
    int ef_info[2] = {0, 0};
    {
        void *info = nullptr;
        size_t info_bytes = 0;
        if (cudaLibraryGetGlobal(&info, &info_bytes, st->lib, "spec_ef_info") != cudaSuccess || info_bytes != sizeof(ef_info) ||
            cudaMemcpy(ef_info, info, sizeof(ef_info), cudaMemcpyDeviceToHost) != cudaSuccess || ef_info[0] < 0 || ef_info[0] > 1)
            ef_info[0] = 0;
        cudaGetLastError();
    }
    if (ef_info[1] != 0 && ef_info[1] != 1) ef_info[1] = 0;
    st->fused = ef_info[1];
    const size_t red_bytes = st->fused ? (size_t)5 * 32 * sizeof(double) : 0;   // [5][32] running sums per warp
    st->smem = (((size_t)(PM_ACOS_N + 1) * sizeof(double2) + 127) & ~(size_t)127) +
               (size_t)warps * (128 + red_bytes + tile * (size_t)(1 + ef_info[0]) + (size_t)3 * n_force_tile_atoms * 32 * sizeof(double));
    if (st->smem > (size_t)h->max_smem ||
        cudaFuncSetAttribute((const void *)st->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem) != cudaSuccess) {
        cudaGetLastError();
        pm_spec_free(st);
        return pm_fail_msg("pm_set_specialized: shared-memory request rejected");
    }
    st->kernel_e = nullptr;
    st->warps_e = 0;
    st->smem_e = 0;
    if (warps_energy > 0 && warps_energy <= 32 && cudaLibraryGetKernel(&st->kernel_e, st->lib, "pm_e_spec_kernel") == cudaSuccess) {
        st->warps_e = warps_energy;
        st->smem_e = (((size_t)(PM_ACOS_N + 1) * sizeof(double2) + 127) & ~(size_t)127) + (size_t)warps_energy * (128 + red_bytes + tile);
        if (st->smem_e > (size_t)h->max_smem ||
            cudaFuncSetAttribute((const void *)st->kernel_e, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem_e) != cudaSuccess)
            st->kernel_e = nullptr;
    } else {
        st->kernel_e = nullptr;
    }
    cudaGetLastError();
    st->kernel_g = nullptr;
    st->kernel_gr = nullptr;
    st->warps_g = 0;
    st->ngp = 0;
    st->smem_g = 0;
    if (warps_grad > 0 && warps_grad <= 32) {
        void *info = nullptr;
        size_t info_bytes = 0;
        int host_info[2] = {0, 0};
        if (cudaLibraryGetKernel(&st->kernel_g, st->lib, "pm_grad_spec_kernel") == cudaSuccess &&
            cudaLibraryGetKernel(&st->kernel_gr, st->lib, "pm_grad_spec_reduce") == cudaSuccess &&
            cudaLibraryGetGlobal(&info, &info_bytes, st->lib, "grad_info") == cudaSuccess && info_bytes == sizeof(host_info) &&
            cudaMemcpy(host_info, info, sizeof(host_info), cudaMemcpyDeviceToHost) == cudaSuccess &&
            host_info[1] == 2 * (h->prog.n_bonds + h->prog.n_angles + h->prog.n_torsions) + 3 * h->prog.n_pairs) {
            st->warps_g = warps_grad;
            st->ngp = host_info[0];
            st->smem_g = (((size_t)(PM_ACOS_N + 1) * sizeof(double2) + 127) & ~(size_t)127) +
                         (size_t)warps_grad * (128 + 2 * tile + (size_t)st->ngp * sizeof(double));
            if (st->smem_g > (size_t)h->max_smem ||
                cudaFuncSetAttribute((const void *)st->kernel_g, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem_g) != cudaSuccess)
                st->kernel_g = nullptr;
        } else {
            st->kernel_g = nullptr;
        }
        cudaGetLastError();
    }
    h->spec_state = st;
    h->spec_free = pm_spec_free;
    return pm_spec_metric_changed(h);
}

int pm_spec_launch_info(pm_handle h, int with_forces, int *block, int *smem_bytes, int *warps) {
    PmSpecState *st = (PmSpecState *)h->spec_state;
    if (!st || (!with_forces && !st->kernel_e)) return 0;
    if (block) *block = (with_forces ? st->warps : st->warps_e) * 32;
    if (smem_bytes) *smem_bytes = (int)(with_forces ? st->smem : st->smem_e);
    if (warps) *warps = with_forces ? st->warps : st->warps_e;
    return 1;
}

int pm_spec_metric_changed(pm_handle h) {
    PmSpecState *st = (PmSpecState *)h->spec_state;
    if (!st) return 0;
    const size_t nd = (size_t)h->prog.doff_pair + 3 * (size_t)h->prog.n_pairs;
    PMS_CUDA(cudaMemcpy(st->tab_dev + nd, h->d_metric, sizeof(double) * 9 * h->prog.n_atoms, cudaMemcpyDeviceToDevice));
    return 0;
}

int pm_spec_can_eval(pm_handle h, int with_forces) {
    PmSpecState *st = (PmSpecState *)h->spec_state;
    return st && (with_forces || st->kernel_e) ? 1 : 0;
}

int pm_spec_fused(pm_handle h) {
    PmSpecState *st = (PmSpecState *)h->spec_state;
    return st && st->fused ? 1 : 0;
}

// rows of five sums one launch of the generated kernel writes per parameter vector when it reduces (one per warp of the grid)
long pm_spec_rows(pm_handle h, int with_forces, long n_conf) {
    PmSpecState *st = (PmSpecState *)h->spec_state;
    if (!st) return 0;
    const int warps = with_forces ? st->warps : st->warps_e;
    if (warps < 1) return 0;
    const long n_tiles = (n_conf + 31) / 32;
    long grid = h->n_sm;
    if (grid * warps > n_tiles) grid = (n_tiles + warps - 1) / warps;
    if (grid < 1) grid = 1;
    return grid * warps;
}

// red.rows != nullptr (images with fused sums only): the kernel adds the five weighted sums of pm_reduce_objective over the
// conformations of each warp and writes [P][grid * warps][5]; emm / fres may then be null (nothing per conformation is written)
int pm_spec_eval(pm_handle h, const double *derived_dev, int n_vec, const double *xt, const double *fref_t, long n_conf,
                 double *emm, double *fres, double *fmm_t, int with_forces, PmRed red, cudaStream_t stream) {
    PmSpecState *st = (PmSpecState *)h->spec_state;
    const cudaKernel_t kernel = with_forces ? st->kernel : st->kernel_e;
    const int warps = with_forces ? st->warps : st->warps_e;
    const size_t smem = with_forces ? st->smem : st->smem_e;
    const size_t nd = (size_t)h->prog.doff_pair + 3 * (size_t)h->prog.n_pairs;
    long n_tiles = (n_conf + 31) / 32;
    const long tile_d = (long)3 * h->prog.n_atoms * 32;
    long grid = h->n_sm;
    if (grid * warps > n_tiles) grid = (n_tiles + warps - 1) / warps;
    if (grid < 1) grid = 1;
    if (red.rows && !st->fused) return pm_fail_msg("pm_spec_eval: this image has no fused residual sums");
    if (!red.rows && !emm) return pm_fail_msg("pm_spec_eval: no output requested");
    const double2 *acos_tab = h->d_acos;
    const long n_rows = grid * warps;
    for (int p = 0; p < n_vec; ++p) {
        // the parameters of this vector become constant-bank operands of the kernel
        PMS_CUDA(cudaMemcpyAsync(st->tab_dev, derived_dev + (size_t)p * h->prog.n_derived, sizeof(double) * nd, cudaMemcpyDeviceToDevice, stream));
        double *e_p = emm ? emm + (size_t)p * n_conf : nullptr, *r_p = fres ? fres + (size_t)p * n_conf : nullptr;
        double *f_p = fmm_t ? fmm_t + (size_t)p * n_tiles * tile_d : nullptr;
        const double *eref = red.eref, *wgt = red.w;
        double d_shift = red.d_shift;
        double *rows_p = red.rows ? red.rows + (size_t)p * n_rows * 5 : nullptr;
        void *args[] = {(void *)&acos_tab, (void *)&xt, (void *)&fref_t, (void *)&n_conf, (void *)&n_tiles, (void *)&e_p, (void *)&r_p, (void *)&f_p,
                        (void *)&eref, (void *)&wgt, (void *)&d_shift, (void *)&rows_p};   // (the last four: fused images only)
        PMS_CUDA(cudaLaunchKernel((const void *)kernel, dim3((unsigned)grid), dim3(warps * 32), args, smem, stream));
    }
    return 0;
}

// ---- generated analytic-gradient kernel ----------------------------------------------------------------------------
int pm_spec_can_grad(pm_handle h) {
    PmSpecState *st = (PmSpecState *)h->spec_state;
    return st && st->kernel_g ? 1 : 0;
}

long pm_spec_grad_scratch(pm_handle h) {
    PmSpecState *st = (PmSpecState *)h->spec_state;
    if (!st || !st->kernel_g) return 0;
    return (long)h->n_sm * st->warps_g * st->ngp;
}

int pm_spec_grad(pm_handle h, const double *derived_dev, const double *xt, const double *fmm_t, const double *fref_t, long n_conf,
                 const double *coef_e, const double *coef_f, double *gout, double *scratch, cudaStream_t stream) {
    PmSpecState *st = (PmSpecState *)h->spec_state;
    const size_t nd = (size_t)h->prog.doff_pair + 3 * (size_t)h->prog.n_pairs;
    long n_tiles = (n_conf + 31) / 32;
    long grid = h->n_sm;
    if (grid * st->warps_g > n_tiles) grid = (n_tiles + st->warps_g - 1) / st->warps_g;
    if (grid < 1) grid = 1;
    PMS_CUDA(cudaMemcpyAsync(st->tab_dev, derived_dev, sizeof(double) * nd, cudaMemcpyDeviceToDevice, stream));
    const double2 *acos_tab = h->d_acos;
    void *args[] = {(void *)&acos_tab, (void *)&xt, (void *)&fmm_t, (void *)&fref_t, (void *)&coef_e, (void *)&coef_f, (void *)&n_conf,
                    (void *)&n_tiles, (void *)&scratch};
    PMS_CUDA(cudaLaunchKernel((const void *)st->kernel_g, dim3((unsigned)grid), dim3(st->warps_g * 32), args, st->smem_g, stream));
    int n_rows = (int)grid * st->warps_g;
    void *rargs[] = {(void *)&scratch, (void *)&n_rows, (void *)&gout};
    PMS_CUDA(cudaLaunchKernel((const void *)st->kernel_gr, dim3((unsigned)((st->ngp + 127) / 128)), dim3(128), rargs, 0, stream));
    return 0;
}

// ---- one objective call of a captured single-vector graph ------------------------------------------------------------
// The host side of ObjectiveFunction.f(x) between "the slots row is known" and "the partial sums are on the host", as two C calls
// instead of five trips through the Python bindings of the CUDA runtime.  pm_graph_launch copies the row into the graph's pinned
// staging buffer and launches the instantiated graph (upload, pm_prepare_params, pm_eval_reduce, download); the caller does its
// host-only work (the regularisation term) while the device runs; pm_graph_wait waits for the stream and copies the sums out.
extern "C" int pm_graph_launch(void *graph_exec, void *stream, const double *in_host, double *pin_in, long n_in) {
    if (!graph_exec || !pin_in || (n_in > 0 && !in_host) || n_in < 0) return pm_fail_msg("pm_graph_launch: null argument");
    if (n_in > 0) memcpy(pin_in, in_host, sizeof(double) * (size_t)n_in);
    PMS_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
    return 0;
}

extern "C" int pm_graph_wait(void *stream, const double *pin_out, long n_out, double *out_host) {
    if (!pin_out || !out_host || n_out < 0) return pm_fail_msg("pm_graph_wait: null argument");
    PMS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    memcpy(out_host, pin_out, sizeof(double) * (size_t)n_out);
    return 0;
}
