"""AddNoise: the plugin surface of RobustART.noise, backed by B200 kernels.

Interface parity with the reference's RobustART/noise/add_noise.py:5-42 (constructor, set_config,
add_noise; same noise names, same config keys and defaults).  Deliberate behavioural fixes, all
documented in DESIGN.md:
  * the noise type is validated BEFORE the default config is looked up (reference :13-14 raises
    KeyError instead of its own assertion message);
  * every instance owns a copy of its config (reference :13,24 mutates the module-level defaults);
  * a file path IS accepted for imagenet-c (reference :36-38 has the assertion inverted).
"""
import copy

from RobustART.noise.utils.add_noise_utils import noise_list, default_config, function_dict


class AddNoise(object):
    """Add noise to one image or a batch.  noise_type is one of `noise_list`."""

    def __init__(self, noise_type):
        assert noise_type in noise_list, f'Add noise only support for {noise_list}'
        self.noise_type = noise_type
        self.config = copy.copy(default_config[noise_type])

    def set_config(self, **kwargs):
        """Override entries of this noise's config; unknown keys are rejected (reference :21-22)."""
        unknown = set(kwargs) - set(self.config)
        assert not unknown, f'Key Error! Unexpect Keys {unknown}'
        self.config.update(kwargs)
        print(f'Config for {self.noise_type} Noise')
        print({k: (type(v).__name__ if hasattr(v, 'forward') or hasattr(v, 'model') else v)
               for k, v in self.config.items()})

    def add_noise(self, image, label=None):
        """imagenet-c / imagenet-s: `image` is a path or a uint8 (n,h,w,3) batch; returns the corrupted
        array.  Adversarial noise: `image` is a float32 CUDA tensor (n,3,h,w) in [0,1], `label` int64 (n)."""
        if isinstance(image, str):
            assert self.noise_type in ('imagenet-s', 'imagenet-c'), \
                'Only imagenet-s and imagenet-c support image path input'
        if self.noise_type in ('imagenet-s', 'imagenet-c'):
            return function_dict[self.noise_type](image, **self.config)
        return function_dict[self.noise_type](image, label, **self.config)
