"""Run the bodies of the gated GPU tests (tests/test_resize_cv_gpu.py, test_token_grad_gpu.py, test_spatter_water_gpu.py) on the HOST,
through robustart_b200.ops and the emulated libraries of tests/test_kernel_emulation_cpu.py: a check of the TEST code (tolerances,
shapes, argument plumbing) before GPU minutes are spent on it.  Not collected by pytest; `python tests/emu/run_gated_gpu_tests_on_host.py`
(about two minutes).  The full-size ViT / Mixer end-to-end test is left out (minutes of emulation per model)."""
import contextlib, os, sys, pathlib, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ.update(B200R_NATIVE_TOKEN_GRAD='1', B200R_SPATTER_WATER='1', B200R_CV_RESIZE='1')
import torch
torch.set_num_threads(1)
import test_kernel_emulation_cpu as E
from robustart_b200 import _lib, ops
emu = E._EmuLibs(pathlib.Path(tempfile.mkdtemp()))
fac = E._Facade(emu, ["resize_cv", "token_layers", "token_backward", "layers", "backward_layers", "loss_metrics", "corrupt"])
_lib._lib = fac
ops._need_cuda = lambda t, dtype, name: None
ops._stream = lambda: None
torch.cuda.device = lambda d: contextlib.nullcontext()
torch.cuda.current_device = lambda: 0
torch.cuda.synchronize = lambda *a: None
cpu = torch.device('cpu')
_orig_device = torch.device
import RobustART.noise.utils.add_noise_utils as U
class _T:  # torch proxy for add_noise_utils: torch.device('cuda', i) -> cpu
    def __getattr__(self, k): return getattr(torch, k)
    @staticmethod
    def device(*a): return cpu
U.torch = _T()
def run(name, fn, *a, **k):
    t = time.time()
    try:
        fn(*a, **k); print("PASS %-60s %.1fs" % (name, time.time() - t))
    except Exception as e:
        import traceback; traceback.print_exc(); print("FAIL", name, repr(e)[:300])
import test_resize_cv_gpu as RC
run("resize_cv (375,500,256,256)", RC.test_resize_cv_matches_cv2, cpu, 375, 500, 256, 256)
run("resize_cv (512,512,256,256)", RC.test_resize_cv_matches_cv2, cpu, 512, 512, 256, 256)
run("imagenet_s opencv types", RC.test_imagenet_s_opencv_types, cpu, pathlib.Path(tempfile.mkdtemp()))
import test_token_grad_gpu as TG
run("layernorm_bwd", TG.test_layernorm_bwd, cpu)
run("act gelu_erf", TG.test_activation_forward_and_backward, cpu, "gelu_erf")
run("attention_bwd (2,50,4)", TG.test_attention_bwd, cpu, 2, 50, 4)
run("patch_scatter", TG.test_patch_scatter_is_the_transpose_of_patch_gather, cpu)
for a in ("gelu_tanh", "tanh"):
    run("act " + a, TG.test_activation_forward_and_backward, cpu, a)
run("attention_bwd (1,1,2)", TG.test_attention_bwd, cpu, 1, 1, 2)
import test_spatter_water_gpu as SW
run("spatter water sev 2", SW.test_spatter_water_branch, cpu, 2)
