"""Result-file metrics of the RobustART benchmark (reference: RobustART/metrics/).  The evaluators the solvers' outputs feed -- AR, WCAR and the ImageNet-C
file evaluator -- live here; they read the result lines `robustart_b200.resultfile` writes
(byte-compatible with the reference's ImageNetDataset.dump)."""
from .AR_evaluator import AdvRobustEvaluator  # noqa: F401
from .WCAR_evaluator import WorstCaseAdvRobustEvaluator  # noqa: F401
from .imagenetc_evaluator import ImageNetCEvaluator  # noqa: F401
