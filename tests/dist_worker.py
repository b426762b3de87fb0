"""
Worker of tests/test_distributed_gloo.py (launched with torch.distributed.run, world_size 2, gloo, CPU only).
Checks the host-side logic of the conformation-sharded objective: shards + one sum-allreduce of the partial rows must
reproduce the single-process objective.  The per-rank partial rows that the GPU reduction kernel would produce are
restated here with numpy from the CPU oracle's energies and forces.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")

from oracle import oracle as O  # noqa: E402
from paramol_b200.MM_engines.b200_engine import B200Engine  # noqa: E402
from paramol_b200.System.system import ParaMolSystem  # noqa: E402
from paramol_b200.Parameter_space.parameter_space import ParameterSpace  # noqa: E402
from paramol_b200.Objective_function.objective_function import ObjectiveFunction  # noqa: E402
from paramol_b200.Objective_function.Properties.energy_property import EnergyProperty  # noqa: E402
from paramol_b200.Objective_function.Properties.force_property import ForceProperty  # noqa: E402
from paramol_b200.MM_engines.least_square import LinearLeastSquare  # noqa: E402
from paramol_b200.Utils import dist as pdist  # noqa: E402
from paramol_b200.Utils.netcdf_lite import read_paramol_data  # noqa: E402
from paramol_b200.Utils.synthetic import perturbed_conformations  # noqa: E402


def make_system(X, E, F):
    engine = B200Engine(True, topology_format="AMBER", crd_format="AMBER", top_file=os.path.join(GOLDEN, "aniline.prmtop"),
                        crd_file=os.path.join(GOLDEN, "aniline.inpcrd"))
    system = ParaMolSystem("aniline", engine, 14, ref_coordinates=X.copy(), ref_energies=E.copy(), ref_forces=F.copy())
    system.force_field.create_force_field(opt_bonds=True, opt_angles=True, opt_torsions=True)
    return system


def partial_rows(emm, fmm, system, metric, d_shift, weights):
    d = np.asarray(system.ref_energies) - emm - d_shift
    diff = fmm - np.asarray(system.ref_forces)
    res = np.einsum("sna,nab,snb->s", diff, metric.reshape(-1, 3, 3), diff)
    w = np.ones(len(d)) if weights is None else weights
    return np.array([[d.sum(), (w * d).sum(), (w * d * d).sum(), w.sum(), (w * res).sum(), float(len(d)), 0.0, 0.0]])


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert world == 2 and pdist.world_size() == 2 and pdist.rank() == rank
    X10, E10, F10 = read_paramol_data(os.path.join(GOLDEN, "aniline_10_struct.nc"))
    n_conf = 101  # odd: uneven shards
    X = perturbed_conformations(X10, n_conf, sigma=0.004, seed=3)
    full = make_system(X, np.zeros(n_conf), np.zeros((n_conf, 14, 3)))
    E, F = O.energies_forces(full.engine.system, X)
    rng = np.random.default_rng(17)
    Eref = E + rng.normal(0, 5.0, n_conf) - 43000.0
    Fref = F + rng.normal(0, 50.0, F.shape)

    # shard bounds cover everything exactly once
    lo, hi = pdist.shard_bounds(n_conf)
    both = [pdist.shard_bounds(n_conf, 2, r) for r in range(2)]
    assert both[0][0] == 0 and both[0][1] == both[1][0] and both[1][1] == n_conf and (lo, hi) == both[rank]

    for weighting in ("uniform", "boltzmann"):
        for term_type in ("norm", "components"):
            # ---- single-process reference values from the oracle ----------------------------------------------
            want_e = O.energy_property(Eref, E, O.conformation_weights(weighting, n_conf, ref_energies=Eref))
            want_f = O.force_property(Fref, F, O.conformation_weights(weighting, n_conf, ref_energies=Eref), term_type)
            # ---- this rank's shard ---------------------------------------------------------------------------
            system = make_system(X, Eref, Fref)
            assert system.shard_for_rank() == (lo, hi) and system.n_structures == hi - lo
            assert system.n_structures_global() == n_conf
            e_prop, f_prop = EnergyProperty([system]), ForceProperty([system], term_type=term_type)
            e_prop.calculate_variance()
            f_prop.calculate_variance()
            assert abs(e_prop.variance[0] - np.var(Eref)) <= 1e-12 * np.var(Eref)
            assert np.allclose(f_prop.variance[0], O.force_variance(Fref), rtol=1e-12, atol=0)
            assert np.allclose(f_prop.inv_covariance[0], O.force_inverse_covariance(Fref), rtol=1e-9, atol=0)
            w_local = system.compute_conformations_weights(temperature=300.0, weighting_method=weighting)
            w_full = O.conformation_weights(weighting, n_conf, ref_energies=Eref)
            assert np.allclose(w_local, w_full[lo:hi], rtol=1e-12, atol=0)
            ps = ParameterSpace()
            ps.get_optimizable_parameters([system])
            of = ObjectiveFunction(None, ps, [e_prop, f_prop], [system], weighting_method=weighting, checkpoint_freq=10 ** 9)
            # partial rows as pm_reduce_objective would write them for this shard, then THE allreduce
            uniform = weighting == "uniform"
            rows = partial_rows(E[lo:hi], F[lo:hi], system, f_prop.metric(0), e_prop.mean_ref[0], None if uniform else w_local)
            t = torch.from_numpy(rows.copy())
            pdist.allreduce_sum_tensor_(t)
            vals = of.assemble_from_partials([t.numpy()], [1.0 / n_conf if uniform else 1.0], set_values=True)
            assert abs(e_prop.value - want_e) <= 1e-10 * abs(want_e), (e_prop.value, want_e)
            assert abs(f_prop.value - want_f) <= 1e-10 * abs(want_f), (f_prop.value, want_f)
            assert abs(vals[0] - (want_e + want_f)) <= 1e-10 * abs(want_e + want_f)
            # the reference-structured property formulas agree as well (distributed sums inside)
            got_e = e_prop.calculate_property([E[lo:hi]])
            got_f = f_prop.calculate_property([F[lo:hi]])
            assert abs(np.sum(got_e) - want_e) <= 1e-10 * abs(want_e)
            assert abs(np.sum(got_f) - want_f) <= 1e-10 * abs(want_f)

    # ---- TSQR gather of the LLS factors ----------------------------------------------------------------------------
    rng = np.random.default_rng(5)
    A = rng.normal(size=(n_conf, 9))
    A[:, 8] = A[:, 0] + A[:, 1]  # rank deficient on purpose: minimum-norm solutions must agree
    b = rng.normal(size=n_conf)
    Q, R = torch.linalg.qr(torch.from_numpy(A[lo:hi]), mode="reduced")
    R_all, q_all = LinearLeastSquare._gather_factors(R, Q.T @ torch.from_numpy(b[lo:hi]))
    got = np.linalg.lstsq(R_all, q_all, rcond=None)[0]
    want = np.linalg.lstsq(A, b, rcond=None)[0]
    assert np.abs(got - want).max() < 1e-10, np.abs(got - want).max()

    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok" % rank)


if __name__ == "__main__":
    main()
