from .coma import COMACritic

__all__ = ["COMACritic"]
