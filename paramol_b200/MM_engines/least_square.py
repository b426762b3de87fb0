# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`LinearLeastSquare` -- linear least-squares fit of the bonded force constants (and, through two-column
expansions, of equilibrium values and torsion phases), with the public surface of the reference class
(``ParaMol/MM_engines/least_square.py:7-693``): ``fit_parameters_lls``, ``_calculate_a``, ``_calculate_b``,
``_add_regularization``, ``_add_symmetries``, ``_reconstruct_parameters``.

What changes underneath:

* the design matrix (reference ``:469-649``: a Python loop over every conformation for every column) is never stored: the
  fused kernels of ``csrc/pm_lls_kernels.cu`` generate its columns -- internal coordinates with the conventions of
  ``ParaMol/Utils/geometry.py``, then 1/2 (q - q0)^2 and 1 + cos(n phi - delta), centred, weighted and scaled -- from the
  conformation tiles that are already resident on the device (the same tiles the E+F kernels read) and contract them on the spot:
  ``pm_lls_colstats`` (column means, observed extrema of the angles), ``pm_lls_normal`` (G = [A b]^T W [A b] on the FP64
  tensor-core instruction) and ``pm_lls_residual`` (b - A x);
* the right-hand side (reference ``:651-692``) is one energy-only launch of the E+F kernel with the optimizable parameters zeroed;
* the (M x M) normal equations -- plus the regularisation and symmetry rows of the reference (``:253-337``) as E^T E -- are solved
  by Cholesky factorisation and polished by iterative refinement against the TRUE residual b - A x (two more passes over the
  tiles per step), which removes the squared condition number from the error of the solution as long as the factorisation is
  usable as a preconditioner (kappa(A) <~ 1e7); when Cholesky breaks down or the refinement does not contract, the fit falls back to
  the orthogonal route of round 1 (``_reduce_tall``: materialised matrix, CholeskyQR2 / Householder, host ``lstsq``: the
  minimum-norm solution of the reference for rank-deficient systems).  ``self.reduction`` records which route ran;
* multi-GPU: every rank contracts its shard; column sums, extrema, G and the refinement gradients are summed with one all-reduce
  each.
"""
import copy
import ctypes

import numpy as np

from .. import _lib
from ..Utils import dist as _dist


def _torch():
    import torch
    return torch


class LlsPlan:
    """``pm_lls`` handle bound to the resident conformation tiles of a :obj:`paramol_b200.MM_engines.device.DeviceEnsemble`."""

    def __init__(self, ensemble, kinds, atoms, pers, ncols):
        torch = _torch()
        self._L = _lib.lib()
        self.ens = ensemble
        self.device = ensemble.device
        self.n_conf = int(ensemble.n_conf)
        self.n_terms = len(kinds)
        arr = [np.ascontiguousarray(a, dtype=np.int32) for a in (kinds, np.asarray(atoms).reshape(-1, 4), pers, ncols)]
        h = ctypes.c_void_p()
        topo = ensemble.topology
        _lib.check(self._L.pm_lls_create(self.device.index, int(topo.n_atoms), int(topo.conf_per_tile), self.n_terms,
                                         arr[0].ctypes.data, arr[1].ctypes.data, arr[2].ctypes.data, arr[3].ctypes.data, ctypes.byref(h)),
                   "pm_lls_create")
        self._h = h
        self.n_cols = int(self._L.pm_lls_n_cols(h))
        self.means = None
        self._scratch = torch.empty(int(self._L.pm_lls_scratch_doubles(h, self.n_conf)), dtype=torch.float64, device=self.device)

    def __del__(self):
        try:
            if self._h:
                self._L.pm_lls_destroy(self._h)
                self._h = None
        except Exception:  # interpreter shutdown
            pass

    def _stream(self):
        return ctypes.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def set_columns(self, p0, mean, scale):
        a = [None if v is None else np.ascontiguousarray(v, dtype=np.float64) for v in (p0, mean, scale)]
        assert all(v is None or v.shape == (self.n_cols,) for v in a)
        _lib.check(self._L.pm_lls_set_columns(self._h, *[None if v is None else v.ctypes.data for v in a]), "pm_lls_set_columns")

    def colstats(self, e, w, want_minmax=False):
        """(y [M] = sum_s w_s e_s (a_sj - mean_j), min/max [T, 2] or None) over this rank's conformations (CUDA tensors)."""
        torch = _torch()
        y = torch.empty(self.n_cols, dtype=torch.float64, device=self.device)
        mm = torch.empty((self.n_terms, 2), dtype=torch.float64, device=self.device) if want_minmax else None
        _lib.check(self._L.pm_lls_colstats(self._h, self.ens.coord_tiles.data_ptr(), self.n_conf, None if e is None else e.data_ptr(),
                                           None if w is None else w.data_ptr(), y.data_ptr(), None if mm is None else mm.data_ptr(),
                                           self._scratch.data_ptr(), self._stream()), "pm_lls_colstats")
        return y, mm

    def residual(self, z, b):
        torch = _torch()
        e = torch.empty(self.n_conf, dtype=torch.float64, device=self.device)
        _lib.check(self._L.pm_lls_residual(self._h, self.ens.coord_tiles.data_ptr(), self.n_conf, z.data_ptr(),
                                           None if b is None else b.data_ptr(), e.data_ptr(), self._stream()), "pm_lls_residual")
        return e

    def normal(self, row_scale, b):
        torch = _torch()
        n = self.n_cols + 1
        g = torch.empty((n, n), dtype=torch.float64, device=self.device)
        _lib.check(self._L.pm_lls_normal(self._h, self.ens.coord_tiles.data_ptr(), self.n_conf,
                                         None if row_scale is None else row_scale.data_ptr(), None if b is None else b.data_ptr(),
                                         g.data_ptr(), self._scratch.data_ptr(), self._stream()), "pm_lls_normal")
        return g


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
        self.refinement_steps = 0
        self.timings = {}
        self._plan = None

    # ------------------------------------------------------------------------------------------
    #                                       PUBLIC METHODS
    # ------------------------------------------------------------------------------------------
    def fit_parameters_lls(self, systems, alpha_bond=0.05, alpha_angle=0.05):
        """Fits the optimizable bonded parameters of ``systems[0]`` (reference ``:68-146``); returns the parameter vector."""
        assert self._weighting_method.upper() != "NON_BOLTZMANN", \
            "LLS does not support {} weighting method.".format(self._weighting_method)
        import time
        torch = _torch()
        system = systems[0]
        self.timings = {}
        t0 = time.perf_counter()

        def lap(name):
            nonlocal t0
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            self.timings[name] = self.timings.get(name, 0.0) + (t1 - t0)
            t0 = t1

        plan = self._device_plan(system, alpha_bond, alpha_angle)          # term table, centres, column means
        lap("plan_and_means_s")
        self._n_parameters = plan.n_cols
        dev = plan.device
        B = torch.from_numpy(self._calculate_b(system)).to(dev)
        lap("rhs_s")
        system.compute_conformations_weights(temperature=self._weighting_temperature, weighting_method=self._weighting_method,
                                             emm=None)
        var = self._global_variance(np.asarray(system.ref_energies, dtype=np.float64))
        row_scale = torch.from_numpy(np.sqrt(np.asarray(system.weights, dtype=np.float64) * np.ones(system.n_structures))
                                     / np.sqrt(var)).to(dev)
        self._calculate_scaling_constants()
        col_scale = 1.0 / self._scaling_constants
        plan.set_columns(self._p0, plan.means, col_scale)
        G = plan.normal(row_scale, B)                                        # [(M+1), (M+1)] = [A b]^T W [A b], this rank's shard
        _dist.allreduce_sum_tensor_(G)                                       # the one data-path collective of the fit
        G = G.cpu().numpy()
        lap("normal_equations_s")
        M = self._n_parameters
        # regularisation and symmetry rows of the reference (E x = f) enter the normal equations as E^T E and E^T f; E is a diagonal
        # block plus rows e_i - e_j, so it is kept sparse (the dense rows are only built for the orthogonal fallback)
        E, f = self._extra_rows_sparse(system)
        N = np.array(G[:M, :M], copy=True)
        ete = (E.T @ E).tocoo()                       # a diagonal plus a handful of off-diagonal entries
        np.add.at(N, (ete.row, ete.col), ete.data)
        rhs = G[:M, M] + E.T @ f
        x = self._solve_refined(plan, N, rhs, E, f, row_scale, B, col_scale)
        lap("solve_and_refine_s")
        if x is None:
            # accuracy fallback: the orthogonal reduction of the materialised matrix and the reference's own lstsq
            A = self._calculate_a_device(system, alpha_bond, alpha_angle)
            A = A * row_scale[:, None] * torch.from_numpy(col_scale).to(dev)[None, :]
            R, qtb = self._reduce_tall(A, B * row_scale)
            R_all, qtb_all = self._gather_factors(R, qtb)
            self._A = np.vstack((R_all, E.toarray()))
            self._B = np.concatenate((qtb_all, f))
            x = np.linalg.lstsq(self._A, self._B, rcond=None)[0]
            lap("fallback_s")
        else:
            self._A, self._B = N, rhs
        self._parameters = x / self._scaling_constants
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

    # ---- fused path ---------------------------------------------------------------------------------------
    #: iterative refinement: accepted when a correction is below this fraction of the solution ...
    REFINE_TOL = 1e-13
    #: ... within this many steps, every step contracting at least by REFINE_MIN_CONTRACTION
    REFINE_MAX_STEPS = 4
    REFINE_MIN_CONTRACTION = 0.1

    def _solve_refined(self, plan, N, rhs, E, f, row_scale, B, col_scale):
        """
        Solution of min || diag(row_scale) (A x - b) ||^2 + || E x - f ||^2 from the normal matrix N = A^T W A + E^T E: Cholesky
        solve, then iterative refinement with the residual recomputed from the tiles (``pm_lls_residual`` + ``pm_lls_colstats``), so
        the error of the solution is governed by kappa(A), not kappa(A)^2.  Returns None when the factorisation fails or the
        refinement does not contract (ill-conditioned or rank-deficient systems: the caller takes the orthogonal route).
        """
        import scipy.linalg
        torch = _torch()
        self.reduction = "householder"
        self.refinement_steps = 0
        if not np.all(np.isfinite(N)) or not np.all(np.isfinite(rhs)):
            return None
        try:
            c = scipy.linalg.cho_factor(N, lower=True, check_finite=False)
        except (np.linalg.LinAlgError, scipy.linalg.LinAlgError, ValueError):
            return None
        x = scipy.linalg.cho_solve(c, rhs, check_finite=False)
        w2 = row_scale * row_scale
        cs = torch.from_numpy(np.ascontiguousarray(col_scale)).to(row_scale.device)
        last = None
        for step in range(self.REFINE_MAX_STEPS):
            z = torch.from_numpy(np.ascontiguousarray(x * col_scale)).to(row_scale.device)   # parameter units
            e = plan.residual(z, B)
            y = plan.colstats(e, w2)[0] * cs
            _dist.allreduce_sum_tensor_(y)
            grad = y.cpu().numpy() + E.T @ (f - E @ x)
            dx = scipy.linalg.cho_solve(c, grad, check_finite=False)
            size = float(np.abs(dx).max()) / max(float(np.abs(x).max()), 1e-300)
            if not np.isfinite(size) or (last is not None and size > self.REFINE_MIN_CONTRACTION * last and size > self.REFINE_TOL):
                return None
            x = x + dx
            self.refinement_steps = step + 1
            if size <= self.REFINE_TOL:
                self.reduction = "normal_equations_refined"
                return x
            last = size
        return None

    def _extra_rows_sparse(self, system):
        """The rows ``_add_regularization`` (diag(alpha / prior width), rhs alpha * p_init; reference ``:253-291``) and
        ``_add_symmetries`` (e_i - e_j, rhs 0; ``:293-337``) append to the system, as (scipy.sparse CSR [K, M], f [K])."""
        import scipy.sparse
        M = self._n_parameters
        rows, cols, vals, f = [], [], [], []
        k = 0
        if self._include_regularization:
            alpha = self._scaling_factor / self._scaling_constants
            self._calculate_prior_widths()
            rows += list(range(M))
            cols += list(range(M))
            vals += list(alpha / self._prior_widths)
            f += list(alpha * self._initial_param_regularization)
            k = M
        # one row per (first column of a symmetry group, later column of the same group), in the reference's order
        groups = {}
        for j, tag in enumerate(self._param_symmetries_list):
            groups.setdefault(tag, []).append(j)
        for tag, members in groups.items():          # dicts keep first-occurrence order
            if tag in ("X_x", "X_y", "X"):
                continue
            for j in members[1:]:
                rows += [k, k]
                cols += [members[0], j]
                vals += [1.0, -1.0]
                f.append(0.0)
                k += 1
        E = scipy.sparse.csr_matrix((np.asarray(vals, dtype=np.float64), (np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64))),
                                    shape=(k, M))
        return E, np.asarray(f, dtype=np.float64)

    def _term_table(self, alpha_bond, alpha_angle):
        """(kinds, atoms, periodicities, columns per term, owners) of the optimizable force constants, in column order."""
        kinds, atoms, pers, ncols, owners = [], [], [], [], []
        partner_key = {"bond_k": ("bond_eq", 0), "angle_k": ("angle_eq", 1), "torsion_k": ("torsion_phase", 2)}
        for parameter in self._parameter_space.optimizable_parameters:
            if parameter.param_key in partner_key:
                key, kind = partner_key[parameter.param_key]
                kinds.append(kind)
                at = list(parameter.ff_term.atoms) + [0, 0, 0]
                atoms.append(at[:4])
                pers.append(int(parameter.ff_term.parameters["torsion_periodicity"].value) if kind == 2 else 0)
                ncols.append(2 if parameter.ff_term.parameters[key].optimize else 1)
                owners.append(parameter)
            elif parameter.param_key not in ["torsion_phase", "bond_eq", "angle_eq"]:
                raise NotImplementedError("Fitting of {} not implemented in LLS.".format(parameter.param_key))
        assert owners, "No bonded force constant is optimizable."
        return kinds, atoms, pers, ncols, owners

    def _column_centres(self, owners, kinds, q_min, q_max, alpha_bond):
        """Fills the per-column bookkeeping lists of the reference (``:486-640``) and returns the centres p0 [M]."""
        self._initial_param_regularization = []
        self._param_keys_list = []
        self._p0 = []
        self._param_symmetries_list = []
        partner_key = {"bond_k": "bond_eq", "angle_k": "angle_eq", "torsion_k": "torsion_phase"}
        for t, parameter in enumerate(owners):
            partner = parameter.ff_term.parameters[partner_key[parameter.param_key]]
            kind = kinds[t]
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
                self._p0.append(centre)
                self._param_keys_list.append(parameter.param_key)
                self._initial_param_regularization.append(parameter.value)
                self._param_symmetries_list.append(parameter.symmetry_group + tag)
        self._initial_param_regularization = np.asarray(self._initial_param_regularization)
        return np.asarray(self._p0, dtype=np.float64)

    def _device_plan(self, system, alpha_bond, alpha_angle):
        """``pm_lls`` plan of the fit on the resident tiles of ``system``: terms, centres (angles: the GLOBAL observed extrema) and
        the global column means."""
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.B200Error("no CUDA device is visible: the B200 path has no CPU fallback")
        kinds, atoms, pers, ncols, owners = self._term_table(alpha_bond, alpha_angle)
        ens = system.device_ensemble()
        plan = LlsPlan(ens, kinds, atoms, pers, ncols)
        # pass 1: observed extrema of the internal coordinates (only the angle expansion needs them)
        q_min = q_max = np.zeros(len(kinds))
        if any(k == 1 and n == 2 for k, n in zip(kinds, ncols)):
            plan.set_columns(np.zeros(plan.n_cols), None, None)
            _, mm = plan.colstats(None, None, want_minmax=True)
            lo, hi = mm[:, 0].contiguous(), mm[:, 1].contiguous()
            _dist.allreduce_min_tensor_(lo)
            hi = -_dist.allreduce_min_tensor_(-hi)
            q_min, q_max = lo.cpu().numpy(), hi.cpu().numpy()
        p0 = self._column_centres(owners, kinds, q_min, q_max, alpha_bond)
        # pass 2: column sums -> global means
        plan.set_columns(p0, None, None)
        sums = plan.colstats(None, None)[0]
        count = float(ens.n_conf)
        if _dist.world_size() > 1:
            sums = _dist.allreduce_sum_tensor_(sums)
            count = float(_dist.allreduce_sum_numpy(np.array([count]))[0])
        plan.means = sums.cpu().numpy() / count
        self._plan = plan
        return plan

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
