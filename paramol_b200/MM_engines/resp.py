# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`RESP` -- restrained electrostatic potential charge fitting, with the public surface of the reference class
(``ParaMol/MM_engines/resp.py:19-613``): ``calculate_inverse_distances``, ``set_charges``, ``set_initial_charges``,
``set_symmetry_constraints``, ``set_not_optimize_atoms``, ``fit_resp_charges_explicitly``.

What changes underneath: the reference materialises ``inv_rij[m][j, i] = 1/|x_mj - g_mi|`` with a triple Python loop and
then forms ``A_aux[m] = R_m R_m^T`` / ``B_aux[m] = R_m V_m`` with a four-deep one.  Here the conformations, grids and
reference potentials are staged on the GPU once and ``pm_resp_assemble`` (include/paramol_b200.h) produces the weighted
sums ``sum_m w_m^2 R_m R_m^T`` and ``sum_m w_m^2 R_m V_m`` directly, generating the inverse distances on the fly.  The
Lagrange rows, the restraint derivative and the (N+c)x(N+c) solve stay on the host (they are O(N^2) work).

The ESP objective of the scipy solver path needs no per-call device pass either: with ``A1 = sum_m w_m R_m R_m^T``,
``B1 = sum_m w_m R_m V_m`` and ``C1 = sum_m w_m V_m.V_m`` it is ``C1 - 2 q.B1 + q^T A1 q`` for any charge vector
(:meth:`esp_objective`).  ``inv_rij`` is still available (computed on demand) for code that reads it.

Ragged grids (a different number of points per conformation) are padded with points placed 1e30 away, whose inverse
distance is 0 to machine precision.
"""
import copy
import ctypes

import numpy as np

from .. import _lib
from ..Utils import dist as _dist


def _torch():
    import torch
    return torch


class RESP:
    """
    Parameters
    ----------
    total_charge : int
    include_regularization : bool
    method : str
        "L2" or "hyperbolic".
    scaling_factor : float
    hyperbolic_beta : float
    weighting_method : str
        "UNIFORM", "MANUAL" or "BOLTZMANN".
    weighting_temperature : float or quantity
    """

    def __init__(self, total_charge, include_regularization, method, scaling_factor, hyperbolic_beta, weighting_method,
                 weighting_temperature, **kwargs):
        self._n_constraints = 1
        self._n_not_optimized = 0
        self._total_charge = total_charge
        self._symmetry_constraints = None
        self._not_optimized = None
        self.charges = None
        self.initial_charges = None
        self._inv_rij = None
        self._A = None
        self._B = None
        self._A_aux = None
        self._B_aux = None
        self._include_regularization = include_regularization
        self._regularization_type = method
        self._scaling_factor = scaling_factor
        self._hyperbolic_beta = hyperbolic_beta
        self._weighting_method = weighting_method
        self._weighting_temperature = weighting_temperature
        self._dev = None
        self._system = None
        self._sums = {}

    # ------------------------------------------------------------------------------------------
    # device staging
    # ------------------------------------------------------------------------------------------
    def calculate_inverse_distances(self, system):
        """
        Stages coordinates, ESP grids and reference potentials of ``system`` on the GPU (the reference computes and
        stores every inverse distance here, ``resp.py:99-134``; this class generates them inside the kernels).
        """
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.B200Error("no CUDA device is visible: the B200 path has no CPU fallback")
        n_conf = len(system.ref_coordinates)
        n_atoms = system.n_atoms
        sizes = [len(v) for v in system.ref_esp]
        g_max = max(sizes)
        coords = np.ascontiguousarray(np.asarray([np.asarray(x, dtype=np.float64) for x in system.ref_coordinates]))
        grid = np.full((n_conf, g_max, 3), 1.0e30)
        vref = np.zeros((n_conf, g_max))
        for m in range(n_conf):
            grid[m, :sizes[m]] = np.asarray(system.ref_esp_grid[m], dtype=np.float64)
            vref[m, :sizes[m]] = np.asarray(system.ref_esp[m], dtype=np.float64)
        dev = torch.device("cuda", torch.cuda.current_device())
        self._dev = dict(device=dev, n_conf=n_conf, n_atoms=n_atoms, n_grid=g_max, sizes=sizes,
                         coords=torch.from_numpy(coords).to(dev), grid=torch.from_numpy(grid).to(dev),
                         vref=torch.from_numpy(vref).to(dev))
        self._system = system
        self._sums = {}
        self._inv_rij = None
        return self

    @property
    def inv_rij(self):
        """list over conformations of [N, G_m] inverse distances (computed on demand; reference attribute)."""
        if self._inv_rij is None and self._dev is not None:
            torch = _torch()
            d = self._dev
            out = []
            for m in range(d["n_conf"]):
                diff = d["coords"][m][:, None, :] - d["grid"][m][None, :d["sizes"][m], :]
                out.append((1.0 / torch.linalg.norm(diff, dim=2)).cpu().numpy())
            self._inv_rij = out
        return self._inv_rij

    @inv_rij.setter
    def inv_rij(self, value):
        self._inv_rij = value

    def _assemble(self, weights):
        """(A [N,N], B [N], C) = sums over ALL ranks' conformations with per-conformation weights (numpy [S] or None)."""
        torch = _torch()
        d = self._dev
        assert d is not None, "calculate_inverse_distances was not called."
        L = _lib.lib()
        n = d["n_atoms"]
        out = torch.empty(n * n + n + 1, dtype=torch.float64, device=d["device"])
        scratch = torch.empty(int(L.pm_resp_scratch_doubles(n, d["n_conf"], d["n_grid"])), dtype=torch.float64, device=d["device"])
        wt = None if weights is None else torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float64)).to(d["device"])
        stream = ctypes.c_void_p(torch.cuda.current_stream(d["device"]).cuda_stream)
        _lib.check(L.pm_resp_assemble(n, d["coords"].data_ptr(), d["grid"].data_ptr(), d["vref"].data_ptr(),
                                      wt.data_ptr() if wt is not None else None, d["n_conf"], d["n_grid"], out.data_ptr(),
                                      scratch.data_ptr(), stream), "pm_resp_assemble")
        _dist.allreduce_sum_tensor_(out)
        host = out.cpu().numpy()
        return host[:n * n].reshape(n, n).copy(), host[n * n:n * n + n].copy(), float(host[n * n + n])

    def esp_ensemble(self):
        """MM electrostatic potential at every grid point of every conformation (list of arrays), current charges."""
        torch = _torch()
        d = self._dev
        L = _lib.lib()
        q = torch.from_numpy(np.ascontiguousarray(self.charges, dtype=np.float64)[:d["n_atoms"]]).to(d["device"])
        vmm = torch.empty((d["n_conf"], d["n_grid"]), dtype=torch.float64, device=d["device"])
        stream = ctypes.c_void_p(torch.cuda.current_stream(d["device"]).cuda_stream)
        _lib.check(L.pm_esp_potential(d["n_atoms"], d["coords"].data_ptr(), d["grid"].data_ptr(), q.data_ptr(), d["n_conf"],
                                      d["n_grid"], vmm.data_ptr(), stream), "pm_esp_potential")
        host = vmm.cpu().numpy()
        return [host[m, :d["sizes"][m]].copy() for m in range(d["n_conf"])]

    def esp_objective(self, charges, weights):
        """
        sum_m w_m sum_i (V_ref - V_mm)^2 for one charge vector [N] or a batch [P, N], from the assembled quadratic form
        (ESPProperty.calculate_property, ``esp_property.py:89-104``).  The sums are cached per weight vector.
        """
        w = np.ascontiguousarray(weights, dtype=np.float64)
        key = ("esp", w.tobytes())
        if key not in self._sums:
            self._sums = {key: self._assemble(w)}
        A, B, C = self._sums[key]
        q = np.atleast_2d(np.asarray(charges, dtype=np.float64))
        val = C - 2.0 * q @ B + np.einsum("pi,ij,pj->p", q, A, q)
        return val if np.ndim(charges) > 1 else float(val[0])

    def esp_objective_gradient(self, charges, weights):
        """d esp_objective / d charges = 2 (A q - B), numpy [N]."""
        w = np.ascontiguousarray(weights, dtype=np.float64)
        key = ("esp", w.tobytes())
        if key not in self._sums:
            self._sums = {key: self._assemble(w)}
        A, B, _ = self._sums[key]
        q = np.asarray(charges, dtype=np.float64).reshape(-1)
        return 2.0 * (0.5 * (A + A.T) @ q - B)

    # ------------------------------------------------------------------------------------------
    # charges / constraints (reference :136-379)
    # ------------------------------------------------------------------------------------------
    def set_charges(self, force_field):
        self.charges = []
        if "NonbondedForce" in force_field:
            for sub_force in force_field["NonbondedForce"]:
                for nonbonded_term in sub_force:
                    self.charges.append(nonbonded_term.parameters["charge"].value)
        return self.charges

    def set_initial_charges(self, force_field):
        self.initial_charges = []
        if "NonbondedForce" in force_field:
            for sub_force in force_field["NonbondedForce"]:
                for nonbonded_term in sub_force:
                    self.initial_charges.append(nonbonded_term.parameters["charge"].value)
        return self.initial_charges

    def set_symmetry_constraints(self, system, symmetry_constrained):
        """Pairs (i, j) of atoms whose charges are constrained to be equal (same symmetry group), reference :285-339."""
        self._symmetry_constraints = []
        if symmetry_constrained:
            ff = system.force_field.force_field['NonbondedForce']
            flat = [term.parameters['charge'] for sub in ff for term in sub]
            symmetry_groups = []
            for i, parameter_i in enumerate(flat):
                if parameter_i.symmetry_group != system.force_field.symmetry_group_default and \
                        parameter_i.symmetry_group not in symmetry_groups:
                    symmetry_groups.append(parameter_i.symmetry_group)
                    for j in range(i + 1, len(flat)):
                        if flat[j].symmetry_group == symmetry_groups[-1]:
                            self._symmetry_constraints.append([i, j])
                            self._n_constraints = self._n_constraints + 1
        return self._symmetry_constraints

    def set_not_optimize_atoms(self, system):
        self._not_optimized = {}
        i = 0
        for sub_force in system.force_field.force_field['NonbondedForce']:
            for term in sub_force:
                parameter = term.parameters['charge']
                if not parameter.optimize:
                    self._not_optimized[i] = parameter.value
                    self._n_not_optimized += 1
                i += 1
        return self._not_optimized

    # ------------------------------------------------------------------------------------------
    # explicit solution (reference :184-283, 381-524)
    # ------------------------------------------------------------------------------------------
    def fit_resp_charges_explicitly(self, system, rmsd_tol, max_iter):
        """Solves the Lagrange-constrained RESP normal equations, iterating the restraint derivative to self-consistency."""
        assert self._weighting_method.upper() in ["UNIFORM", "MANUAL", "BOLTZMANN"], \
            "RESP only accepts the following weighting methods:'UNIFORM', 'MANUAL' or 'BOLTZMANN'"
        if self._regularization_type.upper() == "HYPERBOLIC":
            print("Since HYPERBOLIC regularization was chosen, initial charges that will be optimized will be set to zero.")
            initial_charges_tmp = []
            for atom_id in range(len(self.initial_charges)):
                if atom_id in self._not_optimized:
                    initial_charges_tmp.append(self._not_optimized[atom_id])
                else:
                    initial_charges_tmp.append(0.)
            self.initial_charges = np.asarray(initial_charges_tmp)
        n_iter = 1
        system.compute_conformations_weights(temperature=self._weighting_temperature, weighting_method=self._weighting_method)
        old_charges = self.initial_charges
        self._calculate_a(system, initialize=True)
        self._calculate_b(system, initialize=True)
        self.charges = np.linalg.solve(self._A, self._B)
        new_charges = self.charges[:system.n_atoms]
        rmsd = np.sqrt(np.sum((old_charges - new_charges) ** 2) / system.n_atoms)
        old_charges = copy.deepcopy(new_charges)
        print("Initial net charge: {}".format(self._total_charge))
        print("\n{:20s} {:s}".format("Niter", "RMSD"))
        print("================================")
        print("{:<20d} {:.4e}".format(n_iter, rmsd))
        while n_iter < max_iter and rmsd > rmsd_tol:
            n_iter = n_iter + 1
            self._calculate_a(system, initialize=False)
            self._calculate_b(system, initialize=False)
            self.charges = np.linalg.solve(self._A, self._B)
            new_charges = copy.deepcopy(self.charges[:system.n_atoms])
            rmsd = np.sqrt(np.sum((old_charges - new_charges) ** 2) / system.n_atoms)
            old_charges = copy.deepcopy(new_charges)
            print("{:<20d} {:.4e}".format(n_iter, rmsd))
        print("================================")
        if n_iter < max_iter:
            print("Convergence was achieved.\n")
        else:
            print("Convergence was not achieved. Maximum number of iterations was reached.")
        print("\nFinal net charge: {:.4e}".format(np.sum(self.charges[:system.n_atoms])))
        return self.charges[:system.n_atoms]

    def _n_rows(self, system):
        return system.n_atoms + self._n_constraints + self._n_not_optimized

    def _calculate_a(self, system, initialize=False):
        """
        A = sum_m [ w_m^2 R_m R_m^T  (+) restraint diagonal ] with the Lagrange rows/columns.  As in the reference every
        conformation contributes the constraint entries and the restraint term, i.e. they enter S times (:412-461).
        """
        n = system.n_atoms
        n_global = system.n_structures_global()
        if initialize:
            w2 = np.asarray(system.weights, dtype=np.float64) ** 2 * np.ones(system.n_structures)
            data, rhs, _ = self._assemble(w2)
            size = self._n_rows(system)
            self._A_aux = np.zeros((size, size))
            self._A_aux[:n, :n] = data
            self._A_aux[n, 0:n] = 1.0 * n_global
            self._A_aux[0:n, n] = 1.0 * n_global
            for i in range(len(self._symmetry_constraints)):
                a, b = self._symmetry_constraints[i]
                self._A_aux[n + i + 1, a] = 1.0 * n_global
                self._A_aux[n + i + 1, b] = -1.0 * n_global
                self._A_aux[a, n + i + 1] = 1.0 * n_global
                self._A_aux[b, n + i + 1] = -1.0 * n_global
            for i, (atom_index, _charge_restraint) in enumerate(self._not_optimized.items()):
                self._A_aux[n + self._n_constraints + i, atom_index] = 1.0 * n_global
                self._A_aux[atom_index, n + self._n_constraints + i] = 1.0 * n_global
            self._B_data = rhs
            print("Calculated initial A matrix.")
        self._A = copy.deepcopy(self._A_aux)
        for j in range(n):
            self._A[j, j] = self._A[j, j] + n_global * self._calculate_regularization_derivative(j, self._scaling_factor,
                                                                                                self._hyperbolic_beta)
        return self._A

    def _calculate_b(self, system, initialize=False):
        n = system.n_atoms
        n_global = system.n_structures_global()
        if initialize:
            self._B_aux = np.zeros(self._n_rows(system))
            self._B_aux[:n] = self._B_data
            self._B_aux[n] = self._total_charge * n_global
            for i, (_atom_index, charge_restraint) in enumerate(self._not_optimized.items()):
                self._B_aux[n + self._n_constraints + i] = charge_restraint * n_global
            print("Calculated initial B matrix.")
        self._B = copy.deepcopy(self._B_aux)
        for j in range(n):
            self._B[j] = self._B[j] + n_global * self._calculate_regularization_derivative(
                j, self._scaling_factor, self._hyperbolic_beta) * self.initial_charges[j]
        return self._B

    def _calculate_regularization_derivative(self, at_idx, a=None, b=None):
        if self._include_regularization:
            if self._regularization_type.upper() == "L2":
                return self._regularization_derivative_l2(at_idx, a)
            elif self._regularization_type.upper() == "HYPERBOLIC":
                return self._regularization_derivative_hyperbolic(at_idx, a, b)
            else:
                raise NotImplementedError("Regularization {} scheme not implement.".format(self._regularization_type))
        return 0.0

    def _regularization_derivative_l2(self, at_idx, a):
        """-2 a (q0 - q) / q ; 0 when the division is invalid (the reference runs under np.seterr(all='raise'), :554-582)."""
        if a is None:
            a = self._scaling_factor
        with np.errstate(all='raise'):
            try:
                reg_deriv = - 2.0 * a * (self.initial_charges[at_idx] - self.charges[at_idx]) / self.charges[at_idx]
            except (FloatingPointError, ZeroDivisionError):
                reg_deriv = 0
        return reg_deriv

    def _regularization_derivative_hyperbolic(self, at_idx, a, b):
        if a is None:
            a = self._scaling_factor
        if b is None:
            b = self._hyperbolic_beta
        return a * (self.charges[at_idx] ** 2 + b ** 2) ** (-1 / 2.)
