"""Host-side logic of the plugin surface (no GPU): registry, config handling, error behaviour."""
import numpy as np
import pytest


def test_registry_matches_reference_surface():
    from RobustART.noise.utils.add_noise_utils import noise_list, default_config, function_dict
    assert noise_list == ['imagenet-s', 'imagenet-c', 'pgd_linf', 'pgd_l2', 'fgsm', 'autoattack_linf', 'mim_linf', 'pgd_l1']
    assert set(default_config) == set(noise_list) == set(function_dict)
    assert default_config['pgd_linf'] == {'f_model': None, 'eps': 8 / 255, 'rel_stepsize': 3 / 40, 'steps': 20}
    assert default_config['mim_linf']['step_size'] == 0.002 and default_config['imagenet-c']['severity'] == 1


def test_addnoise_config_rules(capsys):
    from RobustART.noise import AddNoise
    with pytest.raises(AssertionError):
        AddNoise('no-such-noise')
    a, b = AddNoise('imagenet-c'), AddNoise('imagenet-c')
    a.set_config(severity=4, corruption_name='fog')
    assert b.config['severity'] == 1          # the reference leaks this through the shared default dict
    with pytest.raises(AssertionError):
        a.set_config(sevrity=2)
    assert 'Config for imagenet-c Noise' in capsys.readouterr().out
    with pytest.raises(ValueError):
        AddNoise('imagenet-c').add_noise(np.zeros((1, 224, 224, 3), np.uint8))   # neither name nor number
    with pytest.raises(AssertionError):
        AddNoise('pgd_linf').add_noise('some/path.jpg')


def test_no_cpu_fallback():
    import torch
    from robustart_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(TypeError):
        ops.corrupt_u8(torch.zeros((1, 224, 224, 3), dtype=torch.uint8), 'gaussian_noise', 1)
