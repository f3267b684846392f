from .optimiser import AdamOptimizer, TrainOp  # noqa: F401
from .loop import save_results, train  # noqa: F401
