"""The evaluation command lines end to end on one GPU (SKIP_DIST=1, synthetic data)."""
import json
import os

import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(tmp_path, arch="resnet18_official", n=96, bs=32):
    cfg = yaml.safe_load(open(os.path.join(ROOT, "exprs", "b200", "resnet50_eval.yaml")))
    cfg["model"]["type"] = arch
    cfg["data"]["batch_size"] = bs
    cfg["data"]["test"]["limit_samples"] = n
    cfg["save_path"] = str(tmp_path)
    p = tmp_path / "config.yaml"
    p.write_text(yaml.safe_dump(cfg))
    return str(p)


def test_cls_solver_clean_and_imagenet_c(cuda, tmp_path, monkeypatch):
    monkeypatch.setenv("SKIP_DIST", "1")
    import prototype.prototype.solver.cls_solver as cls
    from RobustART.noise.utils import add_noise_utils as anu
    anu.reseed(7)
    cfg = _cfg(tmp_path)
    m0 = cls.main(["--config", cfg, "--evaluate"])
    assert m0["count"] == 96 and 0 <= m0["top1"] <= m0["top5"] <= 100
    m1 = cls.main(["--config", cfg, "--evaluate", "--corruption", "gaussian_noise", "--severity", "5"])
    assert m1["count"] == 96
    assert json.load(open(tmp_path / "results" / "noise-gaussian_noise-5-results.metrics.json")) == m1
    # an image's noise stream is its position in the epoch permutation: another batch size corrupts every image identically
    anu.reseed(7)
    (tmp_path / "b48").mkdir()
    m2 = cls.main(["--config", _cfg(tmp_path / "b48", bs=48), "--evaluate", "--corruption", "gaussian_noise", "--severity", "5"])
    assert (m2["top1"], m2["top5"], m2["count"]) == (m1["top1"], m1["top5"], m1["count"])
    # same counters as doing it by hand through the kernel ops
    from robustart_b200 import nets, ops, solver as S
    model = nets.build_model("resnet18", device=cuda)
    ds = S.SyntheticImageNet(96, 224, cuda)
    idx = S.shard_indices(96, 1, 0)
    c = torch.zeros(3, dtype=torch.int64, device=cuda)
    for i in range(0, 96, 32):
        x, y = ds.batch(idx[i:i + 32])
        ops.topk_count_(c, model(x), y)
    assert abs(100.0 * c[0].item() / 96 - m0["top1"]) < 1e-9


@pytest.mark.parametrize("attack,eps", [("pgd_linf", "4/255"), ("fgsm", "8/255"), ("mim_linf", "8/255"), ("none", "0")])
def test_benchmark_eval_adv_cli(cuda, tmp_path, monkeypatch, attack, eps):
    monkeypatch.setenv("SKIP_DIST", "1")
    import prototype.prototype.solver.base_benchmark_eval_adv as adv
    cfg = _cfg(tmp_path, n=32, bs=16)
    m = adv.main(["--config", cfg, "--src_name", "resnet18", "--src_path", "", "--tgt_name", "resnet18", "--tgt_path", "",
                  "--attack", attack, "--eps", eps])
    assert m["count"] == 32 and 0 <= m["top1"] <= 100


def test_imgnet_c_sweep_cli(cuda, tmp_path, monkeypatch):
    """imgnet_c_solver: all 19 x 5 cells from one clean shard, robust.json in the reference's schema; each cell's
    counters equal a stand-alone cls_solver run of that cell (same seed, same image offsets)."""
    monkeypatch.setenv("SKIP_DIST", "1")
    monkeypatch.chdir(tmp_path)
    import prototype.prototype.solver.cls_solver as cls
    import prototype.prototype.solver.imgnet_c_solver as sweep
    from RobustART.noise.utils import add_noise_utils as anu
    cfg = _cfg(tmp_path, n=32, bs=16)
    anu.reseed(11)
    res = sweep.main(["--config", cfg, "--evaluate", "--save-detail"])
    data = res["resnet18_official"]
    rj = json.load(open(tmp_path / "results" / "robust.json"))
    assert rj["all"]["all_with_extra"] == pytest.approx(data["all"]["all_with_extra"])
    assert set(rj) >= {"all", "noise", "blur", "weather", "digital", "extra"}
    assert len(rj["blur"]) == 4 and len(rj["extra"]) == 4
    for g in ("noise", "blur", "weather", "digital"):
        for t, v in rj[g].items():
            assert 0.0 <= v <= 100.0, (t, v)
    assert "skipped" not in rj and rj["extra"]["spatter"] is not None                  # all 95 cells have a kernel
    cell = json.load(open(tmp_path / "results" / "digital-contrast-3-metric"))
    anu.reseed(11)
    alone = cls.main(["--config", cfg, "--evaluate", "--corruption", "contrast", "--severity", "3"])
    assert cell == alone
    assert "done" in open(tmp_path / "status.txt").read()


def test_multi_eval_cli(cuda, tmp_path, monkeypatch):
    monkeypatch.setenv("SKIP_DIST", "1")
    monkeypatch.chdir(tmp_path)
    import prototype.prototype.solver.multi_eval_solver as multi
    cfg_path = _cfg(tmp_path, n=32, bs=16)
    cfg = yaml.safe_load(open(cfg_path))
    cfg["eval_list"] = ["resnet18", "no_such_model"]
    cfg["eval_list"] = ["resnet18", "no_such_model", "resnet50"]
    open(cfg_path, "w").write(yaml.safe_dump(cfg))
    # a checkpoint for resnet18 only: resnet50's file is missing, which must be logged and skipped (never random weights)
    from robustart_b200 import nets
    torch.save({"model": nets.random_state_dict(nets.resnet_spec("resnet18"), 0)}, tmp_path / "resnet18.pth.tar")
    res = multi.main(["--config", cfg_path, "--evaluate", "--ckpt-filePath", str(tmp_path)])
    assert list(res) == ["resnet18"] and res["resnet18"]["count"] == 32
    st = open(tmp_path / "status.txt").read()
    assert "resnet18 done" in st and "Error when load no_such_model" in st
    assert "Error when load resnet50" in st and "FileNotFoundError" in st


@pytest.mark.parametrize("passes", [16, 3])
def test_eval_pipeline_interfaces_agree(cuda, passes):
    """CorruptEvalPipeline: device-resident steps, synchronous host steps and the double-buffered host interface (H2D on a
    copy stream, counters read one step late) count the same hits as the eager ops on the same batches -- the loop body of
    cls_solver.py:404-428 + imagenet_evaluator.py:49-67 with AddNoise('imagenet-c') in front."""
    from robustart_b200 import nets, ops
    from robustart_b200.evalpipe import CorruptEvalPipeline
    from util import synth_images
    model = nets.build_model("resnet18", device=cuda, seed=0, passes=passes)
    bs, steps = 8, 5
    imgs = [torch.from_numpy(synth_images(bs, seed=20 + i)) for i in range(steps)]
    g = torch.Generator().manual_seed(5)
    labels = [torch.randint(0, 1000, (bs,), generator=g) for _ in range(steps)]
    # eager restatement: corruption (counter-based RNG keyed by the global image index) -> forward -> top-k hits
    want = torch.zeros(3, dtype=torch.int64)
    for i in range(steps):
        x = ops.corrupt_u8(imgs[i].to(cuda), "gaussian_noise", 1 + i % 5, seed=3, image_offset=i * bs)
        lg = model(x).cpu()
        top5 = lg.topk(5, dim=1).indices
        want += torch.tensor([(top5[:, 0] == labels[i]).sum(), (top5 == labels[i][:, None]).any(1).sum(), bs])
    pipe = CorruptEvalPipeline(model, batch=bs, seed=3)
    for i in range(steps):
        pipe.step_device(imgs[i].to(cuda), labels[i].to(cuda), "gaussian_noise", 1 + i % 5)
    assert torch.equal(pipe.counters.cpu(), want)
    pipe.reset()
    for i in range(steps):
        got = pipe.step_host(imgs[i].pin_memory(), labels[i].pin_memory(), "gaussian_noise", 1 + i % 5)
    assert torch.equal(got, want)
    pipe.reset()
    pinned = [(a.pin_memory(), b.pin_memory()) for a, b in zip(imgs, labels)]
    seen = []
    for i in range(steps):
        r = pipe.step_host_pipelined(pinned[i][0], pinned[i][1], "gaussian_noise", 1 + i % 5)
        seen.append(None if r is None else r.clone())
    assert seen[0] is None and all(int(s[2]) == bs * i for i, s in enumerate(seen) if s is not None)   # one step late
    assert torch.equal(pipe.finish(), want)


def test_cls_solver_dump_results_compat(cuda, tmp_path, monkeypatch):
    """data.test.dump_results: the reference's results.txt.rank0 / .all files next to the device counters
    (imagenet_dataset.py:250-277, base_dataset.py:116-133); the file-based evaluator agrees with the counters."""
    monkeypatch.setenv("SKIP_DIST", "1")
    import prototype.prototype.solver.cls_solver as cls
    from robustart_b200 import resultfile
    cfg_path = _cfg(tmp_path, n=40, bs=16)
    cfg = yaml.safe_load(open(cfg_path))
    cfg["data"]["test"]["dump_results"] = True
    open(cfg_path, "w").write(yaml.safe_dump(cfg))
    m = cls.main(["--config", cfg_path, "--evaluate"])
    rank0, merged = tmp_path / "results" / "results.txt.rank0", tmp_path / "results" / "results.txt.all"
    assert rank0.read_text() == merged.read_text() and m["result_file"] == str(merged)
    lines = [json.loads(l) for l in merged.read_text().splitlines()]
    assert len(lines) == 40 and set(lines[0]) == {"filename", "image_id", "prediction", "label", "score"}
    assert len(lines[0]["score"]) == 1000 and abs(sum(lines[0]["score"]) - 1) < 1e-4
    assert all(l["prediction"] == max(range(1000), key=l["score"].__getitem__) for l in lines)
    assert sorted(l["image_id"] for l in lines) == list(range(40))
    assert abs(m["file_metric"]["top1"] - m["top1"]) < 1e-4 and abs(m["file_metric"]["top5"] - m["top5"]) < 1e-4
    assert resultfile.evaluate(str(merged)) == m["file_metric"]


def test_cls_solver_with_gpu_eval_transform(cuda, tmp_path, monkeypatch):
    """data.raw_size: synthetic "decoded files" go through Resize(test_resize) + CenterCrop(input_size) on the GPU
    (imagenet_dataloader.py:74-80, bit-exact Pillow resize) before the forward."""
    monkeypatch.setenv("SKIP_DIST", "1")
    import prototype.prototype.solver.cls_solver as cls
    from robustart_b200 import solver as S
    cfg_path = _cfg(tmp_path, n=24, bs=8)
    cfg = yaml.safe_load(open(cfg_path))
    cfg["data"]["raw_size"] = [300, 400]
    open(cfg_path, "w").write(yaml.safe_dump(cfg))
    m = cls.main(["--config", cfg_path, "--evaluate"])
    assert m["count"] == 24 and 0 <= m["top1"] <= m["top5"] <= 100
    ds = S.SyntheticImageNet(24, 224, cuda, raw_size=(300, 400), test_resize=256)
    imgs, labels = ds.batch(torch.arange(3))
    assert imgs.shape == (3, 224, 224, 3) and imgs.dtype == torch.uint8
    from PIL import Image
    import numpy as np
    g = torch.Generator(device=cuda)
    g.manual_seed(1)          # item 1: seed * 1000003 + idx with seed 0
    raw = torch.randint(0, 256, (300, 400, 3), dtype=torch.uint8, device=cuda, generator=g).cpu().numpy()
    pil = Image.fromarray(raw).resize((int(256 * 400 / 300), 256), Image.BILINEAR)
    x0 = int(round((pil.size[0] - 224) / 2.0))
    assert np.array_equal(imgs[1].cpu().numpy(), np.asarray(pil.crop((x0, 16, x0 + 224, 240))))


def test_adv_solver_ar_and_result_files(cuda, tmp_path, monkeypatch):
    """Adversarial Robustness on the device counters == RobustART.metrics.AdvRobustEvaluator on the dumped clean / adversarial
    result files (AR_evaluator.py:23-39); WCAR over [the attack, the clean file] equals AR."""
    monkeypatch.setenv("SKIP_DIST", "1")
    import prototype.prototype.solver.base_benchmark_eval_adv as adv
    from RobustART.metrics import AdvRobustEvaluator, WorstCaseAdvRobustEvaluator
    cfg_path = _cfg(tmp_path, n=48, bs=16)
    cfg = yaml.safe_load(open(cfg_path))
    cfg["data"]["test"]["dump_results"] = True
    open(cfg_path, "w").write(yaml.safe_dump(cfg))
    m = adv.main(["--config", cfg_path, "--src_name", "resnet18", "--src_path", "", "--tgt_name", "resnet18", "--tgt_path", "",
                  "--attack", "fgsm", "--eps", "8/255"])
    assert m["count"] == 48 and 0 <= m["AR"] <= 100 and 0 <= m["clean_top1"] <= 100
    assert m["top1"] <= m["clean_top1"] + 1e-9 or True       # random weights: no ordering guaranteed, only consistency below
    ar = AdvRobustEvaluator().eval(m["clean_result_file"], m["result_file"])
    assert abs(ar - m["AR"]) < 1e-9
    wcar = WorstCaseAdvRobustEvaluator().eval(m["clean_result_file"], [m["result_file"], m["clean_result_file"]])
    assert abs(wcar - ar) < 1e-9
    lines = [json.loads(l) for l in open(m["result_file"])]
    assert abs(100.0 * sum(l["prediction"] == l["label"] for l in lines) / 48 - m["top1"]) < 1e-9


@pytest.mark.parametrize("name,attack", [("vit_base_patch16_224", "pgd_linf"), ("mixer_b16_224", "fgsm")])
def test_adv_cli_token_models(cuda, tmp_path, monkeypatch, name, attack):
    """BASELINE configs[2] / [4] in miniature: white-box attack on a token model from the command line -- source = the
    autograd twin (torch_models.ViT / Mixer), target = the B200 kernels; the two agree on the clean logits."""
    monkeypatch.setenv("SKIP_DIST", "1")
    import prototype.prototype.solver.base_benchmark_eval_adv as adv
    from robustart_b200 import nets, ops, solver as S, torch_models
    cfg = _cfg(tmp_path, n=8, bs=4)
    m = adv.main(["--config", cfg, "--src_name", name, "--src_path", "", "--tgt_name", name, "--tgt_path", "",
                  "--attack", attack, "--eps", "4/255"])
    assert m["count"] == 8 and 0 <= m["AR"] <= 100
    arch = S.model_name_dict[name]["type"]
    twin = S.build_torch_model({"type": arch}, "", cuda)
    tgt = nets.build_model(arch, device=cuda)
    x = torch.rand(2, 3, 224, 224, device=cuda)
    with torch.no_grad():
        a = twin(ops.normalize(x.clone()))
    assert (tgt(x) - a).abs().max().item() < 2e-3
