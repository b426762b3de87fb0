__all__ = ['parameter_space']
