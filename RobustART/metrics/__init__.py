"""Result-file metrics of the RobustART benchmark (reference: RobustART/metrics/).  Only the two evaluators the adversarial
solvers' outputs feed -- AR and WCAR -- live here; they read the result lines `robustart_b200.resultfile` writes
(byte-compatible with the reference's ImageNetDataset.dump)."""
from .AR_evaluator import AdvRobustEvaluator  # noqa: F401
from .WCAR_evaluator import WorstCaseAdvRobustEvaluator  # noqa: F401
