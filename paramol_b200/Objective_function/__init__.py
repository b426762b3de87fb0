__all__ = ['objective_function', 'Properties']
