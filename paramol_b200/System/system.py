# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`ParaMolSystem` holds the reference (QM) data of one molecule and exposes the ensemble methods the objective
function is built on; same attributes and methods as the reference class for this path
(``ParaMol/System/system.py:32-752``): ``ref_coordinates``, ``ref_energies``, ``ref_forces``, ``ref_esp``,
``ref_esp_grid``, ``weights``, ``wham_weights``, ``n_structures``, ``n_atoms``, ``engine``, ``resp_engine``,
``force_field``, ``n_cpus``; ``compute_conformations_weights``, ``wham_reweighing``, ``get_energies_ensemble``, ``get_forces_ensemble``,
``get_esp_ensemble``, ``energy_statistics``, ``force_statistics``, ``filter_conformations``,
``append_data_to_system``, ``read_data``, ``write_data``.

What changes underneath: the per-conformation ``engine.get_potential_energy`` / ``engine.get_forces`` loops
(reference ``:324-354``) become one kernel launch over a device-resident :obj:`DeviceEnsemble`.

Multi-GPU: when ``torch.distributed`` is initialised each rank's system holds a contiguous SHARD of the
conformations (``shard_for_rank``); everything that is a mean/sum over conformations (weights normalisation,
variances) is made global with a small allreduce (``paramol_b200.Utils.dist``).
"""
import logging

import numpy as np

from ..Force_field.force_field import ForceField
from ..MM_engines.b200_engine import B200Engine
from ..Utils import dist as _dist

# CODATA 2018 (the values simtk.unit carries in the OpenMM version the reference pins)
_KB_J_PER_K = 1.380649e-23
_NA_PER_MOL = 6.02214076e23


def _kelvin(temperature):
    t = getattr(temperature, "_value", temperature)
    return float(t)


class ParaMolSystem:
    """
    Parameters
    ----------
    name : str
    engine : :obj:`paramol_b200.MM_engines.b200_engine.B200Engine`
    n_atoms : int
    resp_engine : :obj:`paramol_b200.MM_engines.resp.RESP`, optional
    ref_coordinates : array [S,N,3] (nm), ref_energies : array [S] (kJ/mol), ref_forces : array [S,N,3] (kJ/mol/nm)
    ref_esp : list of arrays [G_m], ref_esp_grid : list of arrays [G_m,3]
    n_cpus : int
        Kept for signature compatibility (the reference's fork-pool width); unused on the B200 path.
    """

    def __init__(self, name, engine, n_atoms, resp_engine=None, ref_coordinates=None, ref_energies=None, ref_forces=None,
                 ref_esp=None, ref_esp_grid=None, n_cpus=1):
        self.name = name
        self.n_atoms = n_atoms
        self.wham_weights = 1.0
        self.weights = 1.0
        self._was_manual_weighting_set = False
        self.n_structures = 0
        self.ref_coordinates = ref_coordinates
        self.ref_energies = ref_energies
        self.ref_forces = ref_forces
        self.ref_esp = ref_esp
        self.ref_esp_grid = ref_esp_grid
        self.interface = None
        self.engine = engine
        self.resp_engine = resp_engine
        self.n_cpus = n_cpus
        self.qm_engine = None
        if self.ref_coordinates is not None:
            self.n_structures = len(self.ref_coordinates)
        # the reference insists on `type(engine) is OpenMMEngine` (:117-121); here the B200 engine plays that role
        if isinstance(self.engine, B200Engine):
            self.force_field = ForceField(engine)
        else:
            self.force_field = None
            raise NotImplementedError("Currently only B200Engine is supported.")
        self._ensemble = None
        self._ensemble_key = None
        logging.info("Created system named {}.".format(self.name))

    # ------------------------------------------------------------------------------------------
    # device residency
    # ------------------------------------------------------------------------------------------
    def _data_key(self):
        def k(a):
            return None if a is None else (id(a), len(a))
        return (k(self.ref_coordinates), k(self.ref_forces), k(self.ref_energies), id(self.engine._topo),
                getattr(self, "_data_version", 0))

    def invalidate_device_data(self):
        """
        Call after editing ``ref_coordinates`` / ``ref_energies`` / ``ref_forces`` / ``weights`` / ``wham_weights`` IN PLACE
        (``ref_energies -= ref_energies.min()``, ``ref_forces[:] = ...``).  The device-resident copies (and everything
        :obj:`ObjectiveFunction` derives from them: energy shift, static weights, captured CUDA graphs) are keyed on the identity
        of the host arrays, so REPLACING an array is detected, mutating it is not -- the reference re-reads the host arrays on
        every call and has no such rule.  Every method of this class that mutates reference data calls this itself.
        """
        self._data_version = getattr(self, "_data_version", 0) + 1
        self._ensemble = None
        self._ensemble_key = None

    def device_ensemble(self, pinned=False):
        """The device-resident copy of this system's conformations (rebuilt when the reference arrays are replaced)."""
        key = self._data_key()
        if self._ensemble is None or self._ensemble_key != key:
            from ..MM_engines.device import DeviceEnsemble
            assert self.ref_coordinates is not None, "Reference coordinates were not set."
            topo = self.engine.device_topology()
            self._ensemble = DeviceEnsemble(topo, np.asarray(self.ref_coordinates, dtype=np.float64),
                                            None if self.ref_forces is None else np.asarray(self.ref_forces, dtype=np.float64),
                                            None if self.ref_energies is None else np.asarray(self.ref_energies, dtype=np.float64),
                                            pinned=pinned)
            self.n_structures = self._ensemble.n_conf
            self._ensemble_key = self._data_key()
        return self._ensemble

    def shard_for_rank(self, rank=None, world_size=None):
        """Keep only this rank's contiguous shard of the conformations (call on every rank with the same data)."""
        lo, hi = _dist.shard_bounds(len(self.ref_coordinates), world_size, rank)
        self.ref_coordinates = np.asarray(self.ref_coordinates)[lo:hi]
        if self.ref_energies is not None:
            self.ref_energies = np.asarray(self.ref_energies)[lo:hi]
        if self.ref_forces is not None:
            self.ref_forces = np.asarray(self.ref_forces)[lo:hi]
        if self.ref_esp is not None:
            self.ref_esp = list(self.ref_esp)[lo:hi]
        if self.ref_esp_grid is not None:
            self.ref_esp_grid = list(self.ref_esp_grid)[lo:hi]
        self.n_structures = len(self.ref_coordinates)
        return lo, hi

    def n_structures_global(self):
        return int(round(float(_dist.allreduce_sum_numpy(np.array([float(self.n_structures)]))[0])))

    # ------------------------------------------------------------------------------------------
    # weights (reference :165-248)
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def beta(temperature):
        """1 / (kB T NA) in mol/kJ."""
        return 1.0 / (_KB_J_PER_K * _kelvin(temperature) * _NA_PER_MOL / 1000.0)

    def compute_conformations_weights(self, temperature=None, emm=None, weighting_method="UNIFORM", manual_weights_array=None):
        """
        UNIFORM 1/S; BOLTZMANN softmax(-beta (E_ref - mean)); NON_BOLTZMANN softmax(-beta (d - mean d)), d = E_mm - E_ref;
        MANUAL user array.  Overflow/invalid falls back to uniform weights exactly as the reference does with
        ``np.seterr(all='raise')`` (:193, 215-221, 233-239).  Sums and means are global over ranks.
        """
        method = weighting_method.upper()
        n_global = self.n_structures_global()
        if method == "UNIFORM":
            self.weights = np.ones(self.n_structures) / n_global
        elif method in ("NON_BOLTZMANN", "BOLTZMANN"):
            assert temperature is not None, "Temperature was not chosen."
            assert self.ref_energies is not None, "Reference energies were not set."
            if method == "NON_BOLTZMANN":
                assert emm is not None, "Non-Boltzmann weighting was chosen but MM energies were not provided."
                emm_sys = np.asarray(emm, dtype=np.float64)
                if emm_sys.ndim == 2:  # the objective function hands over the per-system list
                    emm_sys = emm_sys[0]
                diff = emm_sys - np.asarray(self.ref_energies)
            else:
                diff = np.asarray(self.ref_energies, dtype=np.float64)
            beta = self.beta(temperature)
            with np.errstate(all='raise'):
                try:
                    mean = float(_dist.allreduce_sum_numpy(np.array([np.sum(diff)]))[0]) / n_global
                    w = np.exp(-beta * (diff - mean))
                    total = float(_dist.allreduce_sum_numpy(np.array([np.sum(w)]))[0])
                    self.weights = w / total
                except FloatingPointError:
                    self.weights = np.ones(self.n_structures) / n_global
        elif method == "MANUAL":
            if not self._was_manual_weighting_set:
                assert manual_weights_array is not None, "No weights array was provided."
                assert len(manual_weights_array) == self.n_structures, \
                    "Weights array provided has different number of weights than the number of structures."
                self.weights = np.asarray(manual_weights_array)
                self._was_manual_weighting_set = True
        else:
            raise NotImplementedError("Weighting method {} is not implemented.".format(weighting_method))
        return self.weights

    def wham_reweighing(self, parameters_generation):
        """
        WHAM weights of the conformations over the parameter generations of an adaptive parametrisation
        (reference ``:250-322``).  The reference runs one full energy pass over the ensemble per generation; here the
        slot rows of all generations are built with the object path and evaluated by ONE energy-only launch (P = number
        of generations).  The fixed-point iteration on the [n_gen, S] weight matrices is the reference's, on the host.
        """
        import torch
        n_gen = len(parameters_generation)
        current_gen = n_gen - 1
        rmsd = 999999.0
        threshold = 1e-6
        if n_gen < 2:
            return np.ones(self.n_structures)
        wham_weights = np.ones((n_gen, self.n_structures)) / self.n_structures
        A = np.ones(n_gen) / n_gen
        rows = []
        for j in range(n_gen):
            self.force_field.update_force_field(parameters_generation[j])
            self.engine.set_bonded_parameters(self.force_field.force_field_optimizable)
            self.engine.set_nonbonded_parameters(self.force_field.force_field_optimizable)
            rows.append(self.engine.current_slots())
        ens = self.device_ensemble()
        derived = ens.topology.prepare(torch.from_numpy(np.ascontiguousarray(np.stack(rows))).to(ens.device))
        if derived.shape[0] >= 16 and ens.energy_batch_supported():
            # many generations: geometry once per conformation, energies by the batch contraction (pm_energy_batch_reduce)
            _, emm = ens.energy_batch_reduce(derived, want_emm=True)
        else:
            emm, _, _ = ens.evaluate(derived, with_forces=False)
        exp_energies = np.exp(-0.5 * emm.cpu().numpy())
        weights = exp_energies / np.sum(exp_energies, axis=1, keepdims=True)
        prev_weight = weights[current_gen, :]
        final_weights = np.zeros((n_gen, self.n_structures))
        G = current_gen
        while rmsd > threshold:
            final_weights = np.zeros((n_gen, self.n_structures))
            for G in range(n_gen):
                for j in range(G + 1):
                    final_weights[G, :] = final_weights[G, :] + A[j] * (weights[G, :] / weights[j, :])
                wham_weights[G, :] = final_weights[G, :]
                A[G] = np.sum(wham_weights[G, :])
                A = A / np.sum(A)
            rmsd = np.sqrt(np.sum((final_weights[current_gen, :] - prev_weight) ** 2) / self.n_structures)
            prev_weight = final_weights[current_gen, :]
        self.wham_weights = final_weights[G, :]
        self.force_field.update_force_field(parameters_generation[current_gen])
        self.engine.set_bonded_parameters(self.force_field.force_field_optimizable)
        self.engine.set_nonbonded_parameters(self.force_field.force_field_optimizable)
        return self.wham_weights

    # ------------------------------------------------------------------------------------------
    # ensembles (reference :324-382)
    # ------------------------------------------------------------------------------------------
    def _derived_current(self, ens):
        import torch
        slots = torch.from_numpy(self.engine.current_slots()[None, :]).to(ens.device)
        return ens.topology.prepare(slots)

    def get_forces_ensemble(self):
        """MM forces of every conformation, numpy [S,N,3] (kJ/mol/nm)."""
        ens = self.device_ensemble()
        _, _, fmm = ens.evaluate(self._derived_current(ens), with_forces=True, want_fres=False, want_fmm=True)
        return ens.forces_from_tiles(fmm)[0].cpu().numpy()

    def get_energies_ensemble(self):
        """MM energies of every conformation, numpy [S] (kJ/mol)."""
        ens = self.device_ensemble()
        emm, _, _ = ens.evaluate(self._derived_current(ens), with_forces=False)
        return emm[0].cpu().numpy()

    def get_esp_ensemble(self):
        """MM electrostatic potential at the grid points of every conformation (list of arrays; atomic units)."""
        assert self.resp_engine is not None
        return self.resp_engine.esp_ensemble()

    # ------------------------------------------------------------------------------------------
    # statistics / data handling (reference :384-565)
    # ------------------------------------------------------------------------------------------
    def energy_statistics(self, emm):
        assert self.ref_energies is not None
        emm = np.concatenate(np.asarray(emm))
        diff = np.sum(np.abs(self.ref_energies - emm - np.mean(self.ref_energies - emm))) / self.n_structures
        rms = np.sqrt(np.mean((self.ref_energies - np.mean(self.ref_energies)) ** 2))
        return diff, rms, diff / rms

    def force_statistics(self, fmm):
        fmm = np.asarray(fmm)
        num = np.sum(np.sqrt(np.sum((self.ref_forces - fmm) ** 2, 2))) / (3 * self.n_atoms * self.n_structures)
        denom = np.sqrt(np.mean(np.sum(np.sum(self.ref_forces ** 2, 2), 1) / (3 * self.n_atoms)))
        return num, denom, num / denom

    def filter_conformations(self, energy_threshold):
        """Keeps the conformations whose reference energy is within ``energy_threshold`` of the minimum."""
        forces_tmp, energies_tmp, coordinates_tmp = [], [], []
        min_energy = np.min(self.ref_energies)
        for i in range(self.n_structures):
            if (self.ref_energies[i] - min_energy) <= energy_threshold:
                coordinates_tmp.append(self.ref_coordinates[i])
                energies_tmp.append(self.ref_energies[i])
                forces_tmp.append(self.ref_forces[i])
        self.ref_coordinates = coordinates_tmp
        self.ref_energies = energies_tmp
        self.ref_forces = forces_tmp
        self.n_structures = len(self.ref_coordinates)
        return self.ref_coordinates, self.ref_energies, self.ref_forces

    def append_data_to_system(self, conformations_list, qm_energies_list, qm_forces_list):
        """Appends conformations and their QM labels to the reference data (reference :461-511)."""
        def cat(old, new):
            new = np.asarray(new)
            if old is None or len(old) == 0:
                return new
            return np.concatenate((np.asarray(old), new))
        self.ref_coordinates = cat(self.ref_coordinates, conformations_list)
        self.ref_energies = cat(self.ref_energies, qm_energies_list)
        self.ref_forces = cat(self.ref_forces, qm_forces_list)
        self.n_structures = len(self.ref_coordinates)
        return self.ref_coordinates, self.ref_energies, self.ref_forces

    def convert_system_ref_arrays_to_list(self):
        """Reference data as lists (reference :513-565)."""
        out = []
        for attr in ("ref_forces", "ref_energies", "ref_coordinates"):
            value = getattr(self, attr)
            if value is None:
                value = []
            elif isinstance(value, np.ndarray):
                value = value.tolist()
            elif not isinstance(value, list):
                raise NotImplementedError("Data type {} is not supported.".format(type(value)))
            setattr(self, attr, value)
            out.append(value)
        return tuple(out)

    def read_data(self, input_file_name=None, append=False):
        """Reads reference coordinates / forces / energies from a ParaMol NetCDF file (reference :567-633)."""
        from ..Utils.netcdf_lite import read_nc
        if input_file_name is None:
            input_file_name = '{}_parmol.nc'.format(self.name)
        print("\nReading {} file for system {}.".format(input_file_name, self.name))
        variables = read_nc(input_file_name)
        for var, attr in (("reference_coordinates", "ref_coordinates"), ("reference_forces", "ref_forces"),
                          ("reference_energies", "ref_energies")):
            if var in variables:
                data = np.asarray(variables[var], dtype=np.float64)
                if append and getattr(self, attr) is not None:
                    setattr(self, attr, np.concatenate((getattr(self, attr), data)))
                else:
                    setattr(self, attr, data)
                if len(data) == 0:
                    setattr(self, attr, None)
                if attr == "ref_coordinates" and self.ref_coordinates is not None:
                    self.n_structures = len(self.ref_coordinates)
            else:
                print("{} does not contain {} data.".format(input_file_name, var))
        print("Data of system {} was read from file {}".format(self.name, input_file_name))

    def write_data(self, output_file_name=None):
        """Writes the reference data to a ParaMol data file (reference :635-686)."""
        from ..Utils.netcdf_lite import write_paramol_raw
        if output_file_name is None:
            output_file_name = '{}_paramol.nc'.format(self.name)
        print("\nWriting {} file for system {}.".format(output_file_name, self.name))
        variables = {}
        for var, attr in (("reference_coordinates", "ref_coordinates"), ("reference_forces", "ref_forces"),
                          ("reference_energies", "ref_energies")):
            value = getattr(self, attr)
            if value is not None and len(value) != 0:
                variables[var] = np.asarray(value, dtype=np.float64)
        write_paramol_raw(output_file_name, variables)
        print("Data of system {} was written to file {}".format(self.name, output_file_name))
