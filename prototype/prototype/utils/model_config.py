# name -> model config (subset of the reference's model_config.py:3-288 that has a B200 kernel path)
from robustart_b200.solver import model_name_dict  # noqa: F401
