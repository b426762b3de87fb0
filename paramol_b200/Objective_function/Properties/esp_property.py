# -*- coding: utf-8 -*-
"""
Description
-----------
:obj:`ESPProperty`: sum over conformations of w_m sum_i (V_ref - V_mm)^2 (reference
``ParaMol/Objective_function/Properties/esp_property.py:68-128``).  Two reference quirks are kept: the variance
is computed but not applied (:102) and ``value`` is overwritten per system rather than summed (:104).
"""
import numpy as np

from .property import Property


class ESPProperty(Property):
    def __init__(self, systems=[], weight=None):
        self.name = 'ESP'
        self.value = None
        self.weight = weight
        self.systems = systems
        self.variance = []
        self.units = "kilojoule_per_mole"

    def add_system(self, system):
        self.systems.append(system)
        return self.systems

    def calculate_property(self, esp_data):
        self.value = 0
        for system, esp, var in zip(self.systems, esp_data, self.variance):
            tmp_value = np.zeros((system.n_structures))
            for m in range(system.n_structures):
                diff = np.asarray(system.ref_esp[m]) - np.asarray(esp[m])
                tmp_value[m] = np.sum(diff * diff)
            self.value = np.sum(tmp_value * system.weights, axis=0)
        return self.value

    def calculate_variance(self):
        self.variance = []
        for system in self.systems:
            assert system.ref_esp is not None, "ERROR: It is not possible to calculate the variance, data was not set."
            for m in range(system.n_structures):
                self.variance.append(np.var(system.ref_esp[m]))
        self.variance = np.asarray(self.variance)
        return self.variance
