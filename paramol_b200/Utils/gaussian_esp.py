# -*- coding: utf-8 -*-
"""
Description
-----------
Reader for Gaussian ESP-fit log files, mirroring the interface of the reference's
``ParaMol.Utils.gaussian_esp.GaussianESP`` (``ParaMol/Utils/gaussian_esp.py:25-117``):
atomic centres and fit centres are converted from angstrom to bohr (1.8897259886), ESP values stay
in atomic units and the last ``SCF Done`` energy is converted to kJ/mol (2625.5002).
"""
import logging
import numpy as np


class GaussianESP:
    angstrom_to_au = 1.8897259886
    hartree_to_kj_mol = 2625.5002

    def __init__(self):
        self.conformations = None
        self.grids = None
        self.esps = None
        self.energies = None

    @staticmethod
    def gaussian_read_log(log_file):
        """
        Returns
        -------
        conformation, grid, esp, energy
        """
        a2b = GaussianESP.angstrom_to_au
        conformation = []
        grid = []
        esp = []
        energy = None
        logging.info("Reading Gaussian file {}.".format(log_file))
        with open(log_file, "r") as gaussian_file:
            for line in gaussian_file:
                if "Fit Center" in line:
                    s = line.split()
                    grid.append([float(s[6]) * a2b, float(s[7]) * a2b, float(s[8]) * a2b])
                elif "Atomic Center" in line:
                    s = line.split()
                    conformation.append([float(s[5]) * a2b, float(s[6]) * a2b, float(s[7]) * a2b])
                elif "Fit    " in line:
                    esp.append(float(line.split()[2]))
                elif "SCF Done:" in line:
                    energy = float(line.split()[4]) * GaussianESP.hartree_to_kj_mol
        return conformation, grid, esp, energy

    def read_log_files(self, files_names):
        """
        Returns
        -------
        conformations, grids, esps, energies : lists with one entry per log file
        """
        self.conformations = []
        self.grids = []
        self.esps = []
        self.energies = []
        if type(files_names) is str:
            files_names = [files_names]
        for gaussian_file in files_names:
            conformation, grid, esp, energy = self.gaussian_read_log(gaussian_file)
            assert len(grid) == len(esp), "Number of grid points and ESP values does not coincide."
            self.conformations.append(np.asarray(conformation))
            self.grids.append(np.asarray(grid))
            self.esps.append(np.asarray(esp))
            self.energies.append(energy)
        return self.conformations, self.grids, self.esps, self.energies
