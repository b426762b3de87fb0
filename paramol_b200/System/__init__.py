__all__ = ['system']
