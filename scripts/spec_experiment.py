#!/usr/bin/env python
"""Experiment: topology-specialised kernel (paramol_b200/csrc/specialize.py) vs the generic K1 on the headline workload.
Run on the GPU box: generates the source, compiles it with nvcc, checks emm / fres / forces against pm_eval and times both."""
import ctypes, os, subprocess, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
from paramol_b200.MM_engines.b200_engine import B200Engine
from paramol_b200.MM_engines.device import DeviceEnsemble
from paramol_b200.csrc import specialize
from paramol_b200.Utils.synthetic import make_ligand

HARNESS = r'''
#include <stdio.h>
extern "C" int spec_launch(const double *derived_host, const double *metric_host, const void *acos_dev, const double *xt, const double *fref,
                           long n_conf, double *emm, double *fres, double *fmm, int grid, int warps, int f_tile) {
    static SpecTab tab;
    for (int i = 0; i < SPEC_ND; ++i) tab.v[i] = derived_host[i];
    for (int i = 0; i < 9 * SPEC_N; ++i) tab.metric[i] = metric_host[i];
    const long n_tiles = (n_conf + 31) / 32;
    size_t smem = (((size_t)(PM_ACOS_N + 1) * 16 + 127) & ~(size_t)127) + (size_t)warps * (128 + (size_t)TILE_D * 8 * f_tile);  /* f_tile = number of tiles per warp */
    cudaFuncSetAttribute(pm_ef_spec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const double2 *at = (const double2 *)acos_dev;
    void *args[] = {&tab, &at, &xt, &fref, &n_conf, (void *)&n_tiles, &emm, &fres, &fmm};
    cudaError_t e = cudaLaunchKernel((const void *)pm_ef_spec_kernel, dim3(grid), dim3(warps * 32), args, smem, 0);
    if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
'''

def build(tag, src):
    path = os.path.join(REPO, "build", "spec_%s.cu" % tag)
    open(path, "w").write(src + HARNESS)
    lib = os.path.join(REPO, "build", "libspec_%s.so" % tag)
    t0 = time.time()
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                        "-Xptxas", "-v", "-I", os.path.join(REPO, "paramol_b200", "csrc"), path, "-o", lib], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    info = [l for l in r.stdout.splitlines() if "registers" in l or "spill" in l or "error" in l]
    print("  [%s] nvcc %.1f s | %s" % (tag, time.time() - t0, " | ".join(s.strip().replace("ptxas info    : ", "") for s in info[-2:])))
    if r.returncode != 0:
        print(r.stdout[-3000:]); return None
    return ctypes.CDLL(lib)

def main(which, S):
    golden = os.path.join(REPO, "tests", "golden")
    if which == "aniline":
        engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(golden, "aniline.prmtop"), crd_file=os.path.join(golden, "aniline.inpcrd"))
        x0 = np.asarray(engine.get_positions(), dtype=np.float64).reshape(-1, 3)
    else:
        mm, x0 = make_ligand(int(which)); engine = B200Engine(system=mm)
    topo_best = engine.device_topology()          # the generic kernel at its own best tile shape
    info_best = topo_best.launch_info(True)
    cpw_best = topo_best.conf_per_tile
    engine._topology = None if hasattr(engine, "_topology") else None
    os.environ["PM_CPW"] = "32"
    from paramol_b200.MM_engines.device import DeviceTopology
    topo = DeviceTopology(engine.system)
    assert topo.conf_per_tile == 32
    N = topo.n_atoms
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.from_numpy(x0).cuda()[None] + 0.005 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
    F = 50.0 * torch.randn((S,) + x0.shape, dtype=torch.float64, device="cuda", generator=g)
    ens = DeviceEnsemble(topo, X, F, torch.zeros(S, dtype=torch.float64, device="cuda"))
    metric = np.tile(np.eye(3).ravel(), (N, 1)) * np.linspace(0.5, 2.0, N)[:, None]
    topo.set_force_metric(metric)
    base = engine.current_slots()
    derived = topo.prepare(torch.from_numpy(base[None].copy()).cuda())
    emm, fres, fmm = ens.evaluate(derived, with_forces=True, want_fres=True, want_fmm=True)
    emm, fres, fmm = emm.clone(), fres.clone(), fmm.clone()
    def timeit(fn, n=10):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    ens_best = DeviceEnsemble(topo_best, X, F, torch.zeros(S, dtype=torch.float64, device="cuda"))
    topo_best.set_force_metric(metric)
    der_best = topo_best.prepare(torch.from_numpy(base[None].copy()).cuda())
    t_gen = timeit(lambda: ens_best.evaluate(der_best, with_forces=True, want_fres=True))
    print("%s N=%d S=%d generic kernel (cpw %d) %.4f ms  (%.1f Mconf/s) launch %s" % (which, N, S, cpw_best, t_gen, S / t_gen / 1e3, info_best))
    del ens_best
    k = topo._keep
    pairs = topo.pairs()
    der_host = derived[0].cpu().numpy().copy()
    acos = torch.tensor([[np.cos(q * np.pi / 256), np.sin(q * np.pi / 256)] for q in range(257)], dtype=torch.float64, device="cuda")
    tile_bytes = 3 * N * 32 * 8
    variants = [(False, False, fence, warps) for fence in (1, 2, 3) for warps in (4, 5, 6, 8, 10)]
    for (f_reg, x_reg, fence, warps) in variants:
        n_tiles_smem = (0 if f_reg else 1) + (0 if x_reg else 1)
        smem = 4224 + warps * (128 + tile_bytes * n_tiles_smem)
        if smem > 227 * 1024: continue
        tag = "%s_%s_%s_f%d_w%d" % (which, "freg" if f_reg else "ftile", "xreg" if x_reg else "xtile", fence, warps)
        lib = build(tag, specialize.generate(N, k["bond"], k["angle"], k["tors"], k["tper"], pairs, warps, f_in_registers=f_reg,
                                             x_in_registers=x_reg, fence_every=fence))
        if lib is None: continue
        e2 = torch.zeros_like(emm[0]); r2 = torch.zeros_like(fres[0]); f2 = torch.zeros_like(fmm[0])
        lib.spec_launch.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_long] + [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3
        def run(fm=None):
            rc = lib.spec_launch(der_host.ctypes.data, metric.ctypes.data, acos.data_ptr(), ens.coord_tiles.data_ptr(), ens.fref_tiles.data_ptr(), S,
                                 e2.data_ptr(), r2.data_ptr(), fm, 148, warps, n_tiles_smem)
            assert rc == 0
        run(f2.data_ptr()); torch.cuda.synchronize()
        de = (e2 - emm[0]).abs().max().item() / emm[0].abs().max().item()
        dr = (r2 - fres[0]).abs().max().item() / fres[0].abs().max().item()
        df = (topo.unpack(f2, S) - topo.unpack(fmm[0], S)).abs().max().item() / fmm[0].abs().max().item()
        t = timeit(lambda: run(None))
        print("  [%s] %.4f ms (%.1f Mconf/s, x%.2f)  rel err E %.1e res %.1e F %.1e" % (tag, t, S / t / 1e3, t_gen / t, de, dr, df), flush=True)

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "aniline", int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20)
