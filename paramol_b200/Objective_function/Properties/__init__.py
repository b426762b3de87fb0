__all__ = ['property', 'energy_property', 'force_property', 'esp_property', 'regularization']
