# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`LinearLeastSquare` -- linear least-squares fit of the bonded force constants (and, through two-column
expansions, of equilibrium values and torsion phases), with the public surface of the reference class
(``ParaMol/MM_engines/least_square.py:7-693``): ``fit_parameters_lls``, ``_calculate_a``, ``_calculate_b``,
``_add_regularization``, ``_add_symmetries``, ``_reconstruct_parameters``.

What changes underneath:

* the design matrix (reference ``:469-649``: a Python loop over every conformation for every column) is two kernels --
  ``pm_lls_internal`` (bond lengths / angles / dihedrals with the conventions of ``ParaMol/Utils/geometry.py``) and
  ``pm_lls_columns`` (1/2 (q - q0)^2 and 1 + cos(n phi - delta)) -- on device-resident coordinates;
* the right-hand side (reference ``:651-692``) is one energy-only launch of the E+F kernel with the optimizable
  parameters zeroed;
* the tall least-squares problem is reduced ON THE DEVICE to the QR factors (R, Q^T b) of the weighted, centred, scaled
  [S, M] matrix (``_reduce_tall``: CholeskyQR2, Householder ``torch.linalg.qr`` when that breaks down; FP64 library
  calls); the regularisation and symmetry rows are appended to
  the M x M triangular factor and the small stacked system goes through ``np.linalg.lstsq`` exactly as the reference's
  full system does.  Since ||A x - b||^2 = ||R x - Q^T b||^2 + const for every x, both systems have the same
  (minimum-norm) minimiser; the solve itself never uses the normal equations, which would square the condition number.
* multi-GPU: every rank factorises its shard, the R factors are gathered (TSQR).
"""
import copy
import ctypes

import numpy as np

from .. import _lib
from ..Utils import dist as _dist


def _torch():
    import torch
    return torch


class LinearLeastSquare:
    """
    Parameters
    ----------
    parameter_space : :obj:`paramol_b200.Parameter_space.parameter_space.ParameterSpace`
    include_regularization : bool
    method : str
        Regularisation type (only "L2" enters the linear system, as in the reference).
    scaling_factor : float
    hyperbolic_beta : float
    weighting_method : str
        "UNIFORM", "BOLTZMANN" or "MANUAL".
    weighting_temperature : float or quantity
    """

    def __init__(self, parameter_space, include_regularization, method, scaling_factor, hyperbolic_beta, weighting_method,
                 weighting_temperature, **kwargs):
        self._parameter_space = parameter_space
        self._parameters = None
        self._n_parameters = None
        self._A = None
        self._B = None
        self._Aw = None
        self._Bw = None
        self._param_keys_list = None
        self._p0 = None
        self._initial_param_regularization = None
        self._param_symmetries_list = None
        self._include_regularization = include_regularization
        self._regularization_type = method
        self._scaling_factor = scaling_factor
        self._hyperbolic_beta = hyperbolic_beta
        self._weighting_method = weighting_method
        self._weighting_temperature = weighting_temperature
        self._A_dev = None
        self.mm_energies_zero = None
        self.reduction = None

    # ------------------------------------------------------------------------------------------
    #                                       PUBLIC METHODS
    # ------------------------------------------------------------------------------------------
    def fit_parameters_lls(self, systems, alpha_bond=0.05, alpha_angle=0.05):
        """Fits the optimizable bonded parameters of ``systems[0]`` (reference ``:68-146``); returns the parameter vector."""
        assert self._weighting_method.upper() != "NON_BOLTZMANN", \
            "LLS does not support {} weighting method.".format(self._weighting_method)
        torch = _torch()
        system = systems[0]
        A = self._calculate_a_device(system, alpha_bond, alpha_angle)
        self._n_parameters = A.shape[1]
        B = torch.from_numpy(self._calculate_b(system)).to(A.device)
        system.compute_conformations_weights(temperature=self._weighting_temperature, weighting_method=self._weighting_method,
                                             emm=None)
        var = self._global_variance(np.asarray(system.ref_energies, dtype=np.float64))
        row_scale = torch.from_numpy(np.sqrt(np.asarray(system.weights, dtype=np.float64) * np.ones(system.n_structures))
                                     / np.sqrt(var)).to(A.device)
        self._calculate_scaling_constants()
        col_scale = torch.from_numpy(1.0 / self._scaling_constants).to(A.device)
        A = A * row_scale[:, None] * col_scale[None, :]
        B = B * row_scale
        # device-side reduction of the tall problem: A = Q R
        R, qtb = self._reduce_tall(A, B)
        R_all, qtb_all = self._gather_factors(R, qtb)
        self._A = R_all
        self._B = qtb_all
        if self._include_regularization:
            self._A, self._B = self._add_regularization()
        self._add_symmetries(system)
        self._parameters = np.linalg.lstsq(self._A, self._B, rcond=None)[0]
        self._parameters = self._parameters / self._scaling_constants
        self._reconstruct_parameters(self._parameters)
        self._parameter_space.get_optimizable_parameters([system], symmetry_constrained=False)
        return self._parameter_space.optimizable_parameters_values

    # ------------------------------------------------------------------------------------------
    #                                       PRIVATE METHODS
    # ------------------------------------------------------------------------------------------
    #: CholeskyQR2 is accepted when || Q1^T Q1 - I ||_F of the first pass stays below this (kappa(A) <~ 1e7)
    CHOLESKY_QR_MAX_DEFECT = 0.25

    def _reduce_tall(self, A, B):
        """
        (R [M, M] upper triangular, Q^T b [M]) of the weighted, centred, scaled design matrix A [S, M], so that
        ``|| A x - b ||^2 = || R x - Q^T b ||^2 + const``: the small system the host ``lstsq`` then solves together with
        the regularisation and symmetry rows (reference ``least_square.py:68-146`` hands the full [S, M] matrix to
        ``np.linalg.lstsq``).

        CholeskyQR2 first -- two passes of (Gram matrix, Cholesky, triangular solve), three FP64 library GEMM-class calls
        per pass, orthogonal to rounding for kappa(A) up to about 1e7 and an order of magnitude faster than Householder
        QR on a 200k x 742 matrix.  The first pass measures its own failure: when the Cholesky factorisation breaks down
        (rank-deficient columns) or Q1 is too far from orthogonal, the Householder factorisation (``torch.linalg.qr``) is
        used instead.  ``self.reduction`` records which one ran.
        """
        torch = _torch()
        self.reduction = "householder"
        if A.shape[0] >= A.shape[1] and A.shape[1] > 0:
            try:
                L1 = torch.linalg.cholesky(A.transpose(0, 1) @ A)                       # A^T A = L1 L1^T, R1 = L1^T
                Q1 = torch.linalg.solve_triangular(L1.transpose(0, 1), A, upper=True, left=False)    # Q1 = A R1^-1
                G2 = Q1.transpose(0, 1) @ Q1
                eye = torch.eye(G2.shape[0], dtype=G2.dtype, device=G2.device)
                defect = float(torch.linalg.matrix_norm(G2 - eye))
                if np.isfinite(defect) and defect < self.CHOLESKY_QR_MAX_DEFECT:
                    L2 = torch.linalg.cholesky(G2)                                       # Q1 = Q R2, R2 = L2^T
                    R = L2.transpose(0, 1) @ L1.transpose(0, 1)
                    qtb = torch.linalg.solve_triangular(L2, (Q1.transpose(0, 1) @ B)[:, None], upper=False)[:, 0]
                    self.reduction = "cholesky_qr2"
                    return R, qtb
            except RuntimeError:      # torch.linalg.LinAlgError: the Gram matrix is not positive definite
                pass
        Q, R = torch.linalg.qr(A, mode="reduced")
        return R, Q.transpose(0, 1) @ B

    @staticmethod
    def _global_variance(values):
        if _dist.world_size() == 1:
            return float(np.var(values))
        n, s1 = _dist.allreduce_sum_numpy(np.array([float(len(values)), np.sum(values)]))
        mean = s1 / n
        return float(_dist.allreduce_sum_numpy(np.array([np.sum((values - mean) ** 2)]))[0] / n)

    @staticmethod
    def _gather_factors(R, qtb):
        torch = _torch()
        if _dist.world_size() == 1:
            return R.cpu().numpy(), qtb.cpu().numpy()
        import torch.distributed as dist
        cpu = dist.get_backend() == "gloo"
        r_loc = R.cpu().contiguous() if cpu else R.contiguous()
        q_loc = qtb.cpu().contiguous() if cpu else qtb.contiguous()
        rows = torch.tensor([r_loc.shape[0]], dtype=torch.int64, device=r_loc.device)
        all_rows = [torch.zeros_like(rows) for _ in range(dist.get_world_size())]
        dist.all_gather(all_rows, rows)
        max_rows = int(max(int(r.item()) for r in all_rows))
        pad_r = torch.zeros((max_rows, r_loc.shape[1]), dtype=r_loc.dtype, device=r_loc.device)
        pad_q = torch.zeros((max_rows,), dtype=q_loc.dtype, device=q_loc.device)
        pad_r[:r_loc.shape[0]] = r_loc
        pad_q[:q_loc.shape[0]] = q_loc
        rs = [torch.zeros_like(pad_r) for _ in range(dist.get_world_size())]
        qs = [torch.zeros_like(pad_q) for _ in range(dist.get_world_size())]
        dist.all_gather(rs, pad_r)
        dist.all_gather(qs, pad_q)
        R_all = np.vstack([r[:int(n.item())].cpu().numpy() for r, n in zip(rs, all_rows)])
        q_all = np.concatenate([q[:int(n.item())].cpu().numpy() for q, n in zip(qs, all_rows)])
        return R_all, q_all

    def _add_regularization(self):
        """Rows diag(alpha / prior_width) with right-hand side alpha * p_init (reference ``:253-291``)."""
        alpha = self._scaling_factor / self._scaling_constants
        self._calculate_prior_widths()
        A_reg = np.identity(self._n_parameters)
        for row in range(A_reg.shape[0]):
            A_reg[row, :] = (A_reg[row, :]) / self._prior_widths
        A_reg = A_reg * alpha
        self._A = np.vstack((self._A, A_reg))
        B_reg = alpha * self._initial_param_regularization
        self._B = np.concatenate((self._B, B_reg))
        return self._A, self._B

    def _add_symmetries(self, system):
        """Rows e_i - e_j for columns that share a symmetry group (reference ``:293-337``)."""
        n_symmetries = 0
        symm_covered = []
        A_symm = []
        for i in range(len(self._param_symmetries_list)):
            symm_i = self._param_symmetries_list[i]
            if symm_i in symm_covered or symm_i in ["X_x", "X_y", "X"]:
                continue
            for j in range(i + 1, len(self._param_symmetries_list)):
                if symm_i == self._param_symmetries_list[j]:
                    row = np.zeros((self._n_parameters))
                    row[i] = 1.0
                    row[j] = -1.0
                    A_symm.append(row)
                    n_symmetries += 1
            symm_covered.append(symm_i)
        if n_symmetries > 0:
            self._A = np.vstack((self._A, np.asarray(A_symm)))
            self._B = np.concatenate((self._B, np.zeros((n_symmetries))))
        return self._A, self._B

    def _calculate_prior_widths(self, method=None):
        self._prior_widths = []
        prior_widths_dict, _ = self._parameter_space.calculate_prior_widths(method=method)
        for i in range(self._n_parameters):
            self._prior_widths.append(prior_widths_dict[self._param_keys_list[i]])
        self._prior_widths = np.asarray(self._prior_widths)
        return self._prior_widths

    def _calculate_scaling_constants(self, method=None):
        self._scaling_constants = []
        scaling_constants_dict, _ = self._parameter_space.calculate_scaling_constants(method=method)
        for i in range(self._n_parameters):
            self._scaling_constants.append(scaling_constants_dict[self._param_keys_list[i]])
        self._scaling_constants = np.asarray(self._scaling_constants)
        return self._scaling_constants

    def _reconstruct_parameters(self, final_parameters):
        """k = k_x + k_y ; x0 = sum(k x0)/sum(k) ; torsion (k, phase) from the phasor sum (reference ``:386-467``)."""
        m = 0
        for parameter in self._parameter_space.optimizable_parameters:
            ff_term = parameter.ff_term
            pairs = {"bond_k": "bond_eq", "angle_k": "angle_eq", "torsion_k": "torsion_phase"}
            if parameter.param_key in pairs:
                partner = ff_term.parameters[pairs[parameter.param_key]]
                if partner.optimize:
                    k_xy = np.asarray(final_parameters[m:m + 2])
                    p_xy = np.asarray(self._p0[m:m + 2])
                    if parameter.param_key == "torsion_k":
                        phasor = k_xy[0] * np.exp(1j * p_xy[0]) + k_xy[1] * np.exp(1j * p_xy[1])
                        parameter.value = np.abs(phasor)
                        partner.value = np.angle(phasor)
                    else:
                        parameter.value = np.sum(k_xy)
                        partner.value = np.sum(k_xy * p_xy) / np.sum(k_xy)
                    m += 2
                else:
                    parameter.value = final_parameters[m]
                    m += 1
            elif parameter.param_key not in ["torsion_phase", "bond_eq", "angle_eq"]:
                raise NotImplementedError("Fitting of {} not implemented in LLS.".format(parameter.param_key))
        return

    # ---- design matrix ----------------------------------------------------------------------------------
    def _calculate_a_device(self, system, alpha_bond=None, alpha_angle=None):
        """Centred design matrix as a CUDA tensor [S, M]; fills the column bookkeeping lists."""
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.B200Error("no CUDA device is visible: the B200 path has no CPU fallback")
        L = _lib.lib()
        dev = torch.device("cuda", torch.cuda.current_device())
        self._initial_param_regularization = []
        self._param_keys_list = []
        self._p0 = []
        self._param_symmetries_list = []
        kinds, atoms, owners = [], [], []
        partner_key = {"bond_k": ("bond_eq", 0), "angle_k": ("angle_eq", 1), "torsion_k": ("torsion_phase", 2)}
        for parameter in self._parameter_space.optimizable_parameters:
            if parameter.param_key in partner_key:
                kinds.append(partner_key[parameter.param_key][1])
                at = list(parameter.ff_term.atoms) + [0, 0, 0]
                atoms.append(at[:4])
                owners.append(parameter)
            elif parameter.param_key not in ["torsion_phase", "bond_eq", "angle_eq"]:
                raise NotImplementedError("Fitting of {} not implemented in LLS.".format(parameter.param_key))
        assert owners, "No bonded force constant is optimizable."
        coords = torch.from_numpy(np.ascontiguousarray(system.ref_coordinates, dtype=np.float64)).to(dev)
        n_conf, n_atoms = int(coords.shape[0]), int(coords.shape[1])
        kind_t = torch.tensor(kinds, dtype=torch.int32, device=dev)
        atoms_t = torch.tensor(atoms, dtype=torch.int32, device=dev)
        q = torch.empty((n_conf, len(kinds)), dtype=torch.float64, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.pm_lls_internal(n_atoms, coords.data_ptr(), n_conf, len(kinds), kind_t.data_ptr(), atoms_t.data_ptr(),
                                     q.data_ptr(), stream), "pm_lls_internal")
        # observed extrema of the angles (two-column expansion of angle_eq), global over ranks
        q_min, q_max = q.amin(dim=0), q.amax(dim=0)
        if _dist.world_size() > 1:
            import torch.distributed as dist
            if dist.get_backend() == "gloo":
                q_min, q_max = q_min.cpu(), q_max.cpu()
            dist.all_reduce(q_min, op=dist.ReduceOp.MIN)
            dist.all_reduce(q_max, op=dist.ReduceOp.MAX)
        q_min, q_max = q_min.cpu().numpy(), q_max.cpu().numpy()
        col_term, col_kind, col_p0, col_per = [], [], [], []
        for t, parameter in enumerate(owners):
            ff_term = parameter.ff_term
            key, kind = partner_key[parameter.param_key]
            partner = ff_term.parameters[key]
            per = int(ff_term.parameters["torsion_periodicity"].value) if kind == 2 else 0
            if partner.optimize:
                if kind == 0:
                    centres = [partner.value * (1 - alpha_bond), partner.value * (1 + alpha_bond)]
                elif kind == 1:
                    centres = [float(q_min[t]), float(q_max[t])]
                else:
                    centres = [-np.pi / 4, np.pi / 4]
                tags = ["_x", "_y"]
            else:
                centres = [partner.value]
                tags = [""]
            for centre, tag in zip(centres, tags):
                col_term.append(t)
                col_kind.append(kind)
                col_p0.append(centre)
                col_per.append(per)
                self._p0.append(centre)
                self._param_keys_list.append(parameter.param_key)
                self._initial_param_regularization.append(parameter.value)
                self._param_symmetries_list.append(parameter.symmetry_group + tag)
        n_cols = len(col_term)
        A = torch.empty((n_conf, n_cols), dtype=torch.float64, device=dev)
        ct = torch.tensor(col_term, dtype=torch.int32, device=dev)
        ck = torch.tensor(col_kind, dtype=torch.int32, device=dev)
        cp = torch.tensor(col_p0, dtype=torch.float64, device=dev)
        cn = torch.tensor(col_per, dtype=torch.int32, device=dev)
        _lib.check(L.pm_lls_columns(q.data_ptr(), n_conf, len(kinds), n_cols, ct.data_ptr(), ck.data_ptr(), cp.data_ptr(),
                                    cn.data_ptr(), A.data_ptr(), stream), "pm_lls_columns")
        # column centring with the global mean
        sums = A.sum(dim=0)
        count = float(n_conf)
        if _dist.world_size() > 1:
            sums = _dist.allreduce_sum_tensor_(sums)
            count = float(_dist.allreduce_sum_numpy(np.array([count]))[0])
        A -= (sums / count)[None, :]
        self._initial_param_regularization = np.asarray(self._initial_param_regularization)
        self._A_dev = A
        return A

    def _calculate_a(self, system, alpha_bond=None, alpha_angle=None):
        """Centred design matrix as a numpy array (reference ``:469-649``)."""
        self._A = self._calculate_a_device(system, alpha_bond, alpha_angle).cpu().numpy()
        return self._A

    # ---- right-hand side -----------------------------------------------------------------------------------
    def _energies_with_zeroed_parameters(self, system):
        """MM energies with every optimizable parameter set to zero; the parameter space is restored afterwards."""
        old_parameter_values = copy.deepcopy(self._parameter_space.optimizable_parameters_values)
        zero_values = np.zeros(len(old_parameter_values))
        self._parameter_space.update_systems([system], zero_values)
        mm_energies_zero = system.get_energies_ensemble()
        self.mm_energies_zero = mm_energies_zero
        self._parameter_space.update_systems([system], np.asarray(old_parameter_values, dtype=np.float64))
        return mm_energies_zero

    def _calculate_b(self, system):
        """b = E_ref - E_mm(optimizable parameters zeroed), centred with the global mean (reference ``:651-692``)."""
        self._B = np.asarray(system.ref_energies, dtype=np.float64) - self._energies_with_zeroed_parameters(system)
        if _dist.world_size() > 1:
            n, s1 = _dist.allreduce_sum_numpy(np.array([float(len(self._B)), np.sum(self._B)]))
            self._B = self._B - s1 / n
        else:
            self._B = self._B - np.mean(self._B)
        return self._B
