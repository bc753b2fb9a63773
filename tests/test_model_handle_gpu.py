"""C-ABI model handles (include/b200r.h b200r_model_*, csrc/model_handle.cu): the C++ layer sequencing must issue the
same launches as robustart_b200/nets.ResNet -- bit-identical logits and input gradients --, hold the golden logits of the
reference's classes, and be drivable from plain C (tests/c/test_model_handle.c, compiled here with gcc)."""
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

from util import synth_images

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "resnet_logits.npz"))


@pytest.mark.parametrize("arch,passes", [("resnet18", 3), ("resnet50", 3), ("resnet18", 16), ("resnet50", 16)])
def test_handle_matches_python_sequencing_and_golden(cuda, arch, passes):
    from robustart_b200 import nets
    from robustart_b200.handle import ModelHandle
    sd = nets.random_state_dict(nets.resnet_spec(arch), 0)
    ref = nets.build_model(arch, sd, device=cuda, passes=passes)
    hm = ModelHandle(arch, {"module." + k: v for k, v in sd.items()}, cuda, passes)     # prefixes are stripped like the solvers do
    images = torch.from_numpy(synth_images(4, seed=7)).to(cuda)
    a, b = hm(images), ref(images)
    assert torch.equal(a, b)
    assert np.abs(a.cpu().numpy() - GOLD[arch]).max() < (1e-3 if passes == 3 else 2e-3)
    x01 = images.permute(0, 3, 1, 2).float().div(255).contiguous()
    y = torch.tensor([1, 2, 3, 4], device=cuda)
    la, vjp = hm.forward_vjp(x01)
    lb, saved = ref.forward_saved(x01)
    assert torch.equal(la, lb)
    from robustart_b200 import ops
    _, d = ops.ce_loss_grad(la, y)
    ga, gb = vjp(d), ref.input_grad(d, saved)
    assert torch.isfinite(ga).all() and torch.equal(ga, gb)
    # a second, larger batch re-sizes the arena; an inference forward invalidates the saved state
    big = torch.from_numpy(synth_images(9, seed=3)).to(cuda)
    assert torch.equal(hm(big), ref(big))
    with pytest.raises(ValueError):
        vjp(torch.zeros(9, 1000, device=cuda))
    hm.close()


def test_handle_as_attack_source(cuda):
    """The handle plugs into the attack loops (attacks.forward_vjp duck-types on .forward_vjp)."""
    from robustart_b200 import attacks, nets
    from robustart_b200.handle import ModelHandle
    sd = nets.random_state_dict(nets.resnet_spec("resnet18"), 1)
    hm = ModelHandle("resnet18", sd, cuda, 3)
    ref = attacks.NativeModel(nets.build_model("resnet18", sd, device=cuda, passes=3), use_graphs=False)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(3, 3, 224, 224, generator=g).to(cuda)
    y = torch.randint(0, 1000, (3,), generator=g).to(cuda)
    u = torch.rand(3, 3, 224, 224, generator=g).to(cuda)
    a = attacks.pgd_linf(x, y, hm, 4 / 255, 3 / 40, 3, start_uniform=u)
    b = attacks.pgd_linf(x, y, ref, 4 / 255, 3 / 40, 3, start_uniform=u)
    assert torch.equal(a, b)


def test_eval_pipeline_on_the_handle(cuda):
    """CorruptEvalPipeline (corrupt -> graphed forward -> counters) runs on the C++ handle exactly as on nets.ResNet."""
    from robustart_b200 import nets
    from robustart_b200.evalpipe import CorruptEvalPipeline
    from robustart_b200.handle import ModelHandle
    sd = nets.random_state_dict(nets.resnet_spec("resnet18"), 0)
    a = CorruptEvalPipeline(ModelHandle("resnet18", sd, cuda, 3), batch=8, seed=3)
    b = CorruptEvalPipeline(nets.build_model("resnet18", sd, device=cuda, passes=3), batch=8, seed=3)
    assert a.model.launches_per_forward() == b.model.launches_per_forward()
    g = torch.Generator().manual_seed(1)
    for step in range(3):
        imgs = torch.from_numpy(synth_images(8, seed=30 + step)).to(cuda)
        labels = torch.randint(0, 1000, (8,), generator=g).to(cuda)
        la = a.step_device(imgs, labels, "gaussian_noise", 1 + step).clone()
        lb = b.step_device(imgs, labels, "gaussian_noise", 1 + step).clone()
        assert torch.equal(la, lb)
    assert a.counters.tolist() == b.counters.tolist() and a.counters.tolist()[2] == 24


def test_handle_from_plain_c(cuda, tmp_path):
    from robustart_b200 import nets, ops
    arch, passes, n, h, w = "resnet18", 3, 2, 64, 96
    sd = nets.random_state_dict(nets.resnet_spec(arch), 0)
    ref = nets.build_model(arch, sd, device=cuda, passes=passes)
    images = torch.from_numpy(synth_images(n, seed=5, h=h, w=w)).to(cuda)
    logits_u8 = ref(images)
    x01 = images.permute(0, 3, 1, 2).float().div(255).contiguous()
    logits_f32, saved = ref.forward_saved(x01)
    _, d = ops.ce_loss_grad(logits_f32, torch.tensor([3, 5], device=cuda))
    dx = ref.input_grad(d, saved)
    keep = [(k, v.float().contiguous()) for k, v in sd.items() if not k.endswith("num_batches_tracked")]
    dump = tmp_path / "dump.bin"
    with open(dump, "wb") as f:
        f.write(struct.pack("<7i", 0, passes, n, h, w, 1000, len(keep)))
        for k, v in keep:
            f.write(struct.pack("<i", len(k)) + k.encode() + struct.pack("<q", v.numel()) + v.numpy().tobytes())
        f.write(images.cpu().numpy().tobytes())
        for t in (x01, logits_u8, d, logits_f32, dx):
            f.write(t.float().cpu().numpy().tobytes())
    exe = tmp_path / "t_handle"
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    lib = os.path.join(ROOT, "robustart_b200", "lib")
    subprocess.run(["gcc", os.path.join(ROOT, "tests", "c", "test_model_handle.c"), "-I", os.path.join(ROOT, "include"), "-I", cuda_home + "/include",
                    "-L", lib, "-lb200robust", "-L", cuda_home + "/lib64", "-lcudart", "-lm", "-o", str(exe)], check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=lib + ":" + cuda_home + "/lib64:" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([str(exe), str(dump)], capture_output=True, text=True, env=env)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "PASS" in r.stdout, (r.stdout, r.stderr)


def test_allreduce_counts_single_rank_communicator(cuda):
    """b200r_allreduce_counts on a real ncclComm_t (one rank, created through NCCL's own C API): the counters come back unchanged
    (a sum over one rank) and the launch is stream-ordered.  The N > 1 path is the same call; bench.py exercises torch's communicator."""
    import ctypes as C
    from robustart_b200 import _lib
    import torch.cuda.nccl  # noqa: F401  (loads libnccl into the process)
    nccl = None
    for name in ("libnccl.so.2", "libnccl.so"):
        try:
            nccl = C.CDLL(name, mode=C.RTLD_GLOBAL)
            break
        except OSError:
            continue
    if nccl is None:
        import glob
        import site
        cands = [p for sp in site.getsitepackages() for p in glob.glob(os.path.join(sp, "nvidia", "nccl", "lib", "libnccl.so*"))]
        if not cands:
            pytest.skip("libnccl not found")
        nccl = C.CDLL(cands[0], mode=C.RTLD_GLOBAL)
    uid = (C.c_char * 128)()
    assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    comm = C.c_void_p()

    class _Uid(C.Structure):
        _fields_ = [("internal", C.c_char * 128)]
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _Uid, C.c_int]
    u = _Uid()
    C.memmove(C.byref(u), uid, 128)
    with torch.cuda.device(cuda):
        assert nccl.ncclCommInitRank(C.byref(comm), 1, u, 0) == 0
        counts = torch.tensor([17, 42, 256], dtype=torch.int64, device=cuda)
        _lib.check(_lib.load().b200r_allreduce_counts(comm, counts.data_ptr(), 3, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert counts.tolist() == [17, 42, 256]
        nccl.ncclCommDestroy.argtypes = [C.c_void_p]
        nccl.ncclCommDestroy(comm)


def test_vit_handle_matches_python_sequencing(cuda):
    """ViT-B/16 behind the C-ABI (b200r_model_create(B200R_ARCH_VIT_B16, ...)): logits from uint8 and float input, and the input
    gradient, bit-identical to nets.ViT's launch sequence (vision_transformer.py:44-349)."""
    from robustart_b200 import nets, ops
    from robustart_b200.handle import ModelHandle
    sd = nets.random_token_state_dict(nets.vit_spec(), 0)
    ref = nets.build_model("vit_b16_224", sd, device=cuda)
    hm = ModelHandle("vit_b16_224", {"module." + k: v for k, v in sd.items()}, cuda, 3)
    assert hm.num_classes == ref.num_classes
    images = torch.from_numpy(synth_images(3, seed=5)).to(cuda)
    assert torch.equal(hm(images), ref(images))
    x01 = images.permute(0, 3, 1, 2).float().div(255).contiguous()
    la, vjp = hm.forward_vjp(x01)
    lb, saved = ref.forward_saved(x01)
    assert torch.equal(la, lb)
    _, d = ops.ce_loss_grad(la, torch.tensor([7, 1, 3], device=cuda))
    ga, gb = vjp(d), ref.input_grad(d, saved)
    assert torch.isfinite(ga).all() and ga.abs().max().item() > 0
    assert torch.equal(ga, gb)
    big = torch.from_numpy(synth_images(5, seed=6)).to(cuda)              # a larger batch re-sizes the arena
    assert torch.equal(hm(big), ref(big))
    with pytest.raises(ValueError):
        hm(torch.zeros(2, 160, 160, 3, dtype=torch.uint8, device=cuda))  # 101 tokens against a 197-row position embedding
    hm.close()


def test_mixer_handle_matches_python_sequencing(cuda):
    """MLP-Mixer-B/16 behind the C-ABI (B200R_ARCH_MIXER_B16): forward from uint8 / float input and the input gradient, bit-identical
    to nets.Mixer's launch sequence (vit/mlp_mixer.py:7-159)."""
    from robustart_b200 import nets, ops
    from robustart_b200.handle import ModelHandle
    sd = nets.random_token_state_dict(nets.mixer_spec(), 0)
    ref = nets.build_model("mixer_b16_224", sd, device=cuda)
    hm = ModelHandle("mixer_b16_224", sd, cuda, 3)
    images = torch.from_numpy(synth_images(3, seed=8)).to(cuda)
    assert torch.equal(hm(images), ref(images))
    x01 = images.permute(0, 3, 1, 2).float().div(255).contiguous()
    la, vjp = hm.forward_vjp(x01)
    lb, saved = ref.forward_saved(x01)
    assert torch.equal(la, lb)
    _, d = ops.ce_loss_grad(la, torch.tensor([7, 1, 3], device=cuda))
    ga, gb = vjp(d), ref.input_grad(d, saved)
    assert torch.isfinite(ga).all() and ga.abs().max().item() > 0
    assert torch.equal(ga, gb)
    hm.close()


@pytest.mark.parametrize("arch", ["mobilenet_v2", "efficientnet_b0"])
def test_mobile_inference_handles_match_python_sequencing(cuda, arch):
    """MobileNetV2 / EfficientNet-B0 inference handles (B200R_ARCH_MOBILENET_V2 / _EFFICIENTNET_B0; the ImageNet-C sweep of BASELINE
    configs[3] is evaluation only): logits bit-identical to nets.MobileNetV2 / nets.EfficientNetB0; the gradient calls are refused."""
    from robustart_b200 import nets
    from robustart_b200.handle import ModelHandle
    sd = nets.random_state_dict(nets._MOBILE_ARCHS[arch][1](), 0)
    ref = nets.build_model(arch, sd, device=cuda)
    hm = ModelHandle(arch, sd, cuda, 3)
    assert hm.num_classes == ref.num_classes
    for n, seed in ((3, 5), (6, 9)):                                     # the second, larger batch re-sizes the arena
        images = torch.from_numpy(synth_images(n, seed=seed)).to(cuda)
        assert torch.equal(hm(images), ref(images))
    with pytest.raises(NotImplementedError):
        hm.forward_vjp(torch.rand(2, 3, 224, 224, device=cuda))
    hm.close()
