# -*- coding: utf-8 -*-
"""Base class of the objective-function properties (reference ``ParaMol/Objective_function/Properties/property.py:10``)."""


class Property:
    def __init__(self):
        self.name = None
        self.value = None
        self.weight = None

    def calculate_variance(self):
        pass
