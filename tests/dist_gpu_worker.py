"""
Worker of tests/test_gpu_distributed.py: one process per GPU (NCCL).  Every rank evaluates the objective on its shard
through ObjectiveFunction.f / f_batch (one allreduce per call) and, separately, on ALL conformations on its own GPU; the
two must agree to 1e-12 relative, for uniform, Boltzmann and non-Boltzmann weights; so must the analytic gradient.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")

from paramol_b200.MM_engines.b200_engine import B200Engine  # noqa: E402
from paramol_b200.System.system import ParaMolSystem  # noqa: E402
from paramol_b200.Parameter_space.parameter_space import ParameterSpace  # noqa: E402
from paramol_b200.Objective_function.objective_function import ObjectiveFunction  # noqa: E402
from paramol_b200.Objective_function.Properties.energy_property import EnergyProperty  # noqa: E402
from paramol_b200.Objective_function.Properties.force_property import ForceProperty  # noqa: E402
from paramol_b200.Objective_function.Properties.regularization import Regularization  # noqa: E402
from paramol_b200.Utils.netcdf_lite import read_paramol_data  # noqa: E402
from paramol_b200.Utils.synthetic import perturbed_conformations  # noqa: E402


def build(X, E, F, shard, weighting):
    engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(GOLDEN, "aniline.prmtop"),
                        crd_file=os.path.join(GOLDEN, "aniline.inpcrd"))
    system = ParaMolSystem("aniline", engine, 14, ref_coordinates=X.copy(), ref_energies=E.copy(), ref_forces=F.copy())
    system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True, opt_charges=True, opt_lj=True)
    if shard:
        system.shard_for_rank()
    ps = ParameterSpace()
    _, x0 = ps.get_optimizable_parameters([system])
    e_prop, f_prop = EnergyProperty([system]), ForceProperty([system], term_type="components")
    e_prop.calculate_variance()
    f_prop.calculate_variance()
    ps.calculate_prior_widths("default")
    r_prop = Regularization(np.asarray(x0).copy(), ps.prior_widths, "L2")
    of = ObjectiveFunction(None, ps, [e_prop, f_prop, r_prop], [system], weighting_method=weighting, checkpoint_freq=10 ** 9)
    return of, np.asarray(x0)


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    X10, _, _ = read_paramol_data(os.path.join(GOLDEN, "aniline_10_struct.nc"))
    n_conf = 1003
    X = perturbed_conformations(X10, n_conf, sigma=0.004, seed=3)
    rng = np.random.default_rng(17)
    E = rng.normal(-43000.0, 8.0, n_conf)
    F = rng.normal(0, 300.0, (n_conf, 14, 3))
    for weighting in ("uniform", "boltzmann", "non_boltzmann"):
        of_shard, x0 = build(X, E, F, True, weighting)
        rngx = np.random.default_rng(1)
        trial = np.stack([x0 * (1.0 + 0.01 * rngx.uniform(-1, 1, x0.shape)) for _ in range(3)])
        got = [of_shard.f(x.copy()) for x in trial]
        got_batch = of_shard.f_batch(trial)
        values = torch.tensor(got + list(got_batch), dtype=torch.float64, device="cuda")
        ref = values.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(values, ref), "ranks disagree on the objective"
        if rank == 0:
            np.save(os.path.join(REPO, "gpurun_out", "dist_values_%s_w%d.npy" % (weighting, world)), values.cpu().numpy())
        if weighting != "non_boltzmann":     # analytic gradient: one more sum-allreduce (of the term gradient)
            v, g = of_shard.f_and_grad(trial[0].copy())
            assert abs(v - got[0]) <= 1e-12 * abs(v)
            gt = torch.from_numpy(g).cuda()
            ref_g = gt.clone()
            dist.broadcast(ref_g, src=0)
            assert torch.equal(gt, ref_g), "ranks disagree on the gradient"
            if rank == 0:
                np.save(os.path.join(REPO, "gpurun_out", "dist_grad_%s_w%d.npy" % (weighting, world)), g)
    # ---- energy-only batches (pm_energy_batch_reduce, the cross-GPU sum inside the row-sum kernel) and the LLS fit (one all-reduce
    # each of extrema, column sums, normal equations and refinement gradients): sharded == all conformations on one GPU
    from paramol_b200.MM_engines.least_square import LinearLeastSquare
    for shard in (True, False):
        import contextlib
        from paramol_b200.Utils import dist as pdist
        # the unsharded reference runs with every distributed helper switched to single-rank behaviour (both ranks compute it)
        with (contextlib.nullcontext() if shard else pdist.local_only()):
            engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(GOLDEN, "aniline.prmtop"),
                                crd_file=os.path.join(GOLDEN, "aniline.inpcrd"))
            system = ParaMolSystem("aniline", engine, 14, ref_coordinates=X.copy(), ref_energies=E.copy(), ref_forces=F.copy())
            system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True)
            if shard:
                system.shard_for_rank()
            ps = ParameterSpace()
            _, x0 = ps.get_optimizable_parameters([system])
            x0 = np.asarray(x0)
            e_prop = EnergyProperty([system])
            e_prop.calculate_variance()
            ps.calculate_prior_widths("default")
            of = ObjectiveFunction(None, ps, [e_prop, Regularization(x0.copy(), ps.prior_widths, "L2")], [system], weighting_method="uniform",
                                   checkpoint_freq=10 ** 9)
            rngx = np.random.default_rng(5)
            trial = np.stack([x0 * (1.0 + 0.01 * rngx.uniform(-1, 1, x0.shape)) for _ in range(20)])   # >= 16: the contraction route
            vals = np.asarray(of.f_batch(trial))
            ps.update_systems([system], x0)
            ps.get_optimizable_parameters([system], symmetry_constrained=False)
            lls = LinearLeastSquare(ps, True, method="L2", scaling_factor=1.0, hyperbolic_beta=0.01, weighting_method="uniform",
                                    weighting_temperature=300.0)
            lls.fit_parameters_lls([system])
            fitted = np.asarray(lls._parameters, dtype=np.float64)
        if rank == 0:
            tag = "shard" if shard else "full"
            np.save(os.path.join(REPO, "gpurun_out", "dist_ebatch_%s_w%d.npy" % (tag, world)), vals)
            np.save(os.path.join(REPO, "gpurun_out", "dist_lls_%s_w%d.npy" % (tag, world)), fitted)
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok" % rank)


if __name__ == "__main__":
    main()
