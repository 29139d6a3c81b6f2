from .ik_solver import BatchIKResult, IKResult, IKSolver, JacobianIKController  # noqa: F401
