from .add_noise_utils import noise_list, default_config, function_dict  # noqa: F401
