"""Host-side logic of the plugin surface (no GPU): registry, config handling, error behaviour."""
import numpy as np
import pytest


def test_registry_matches_reference_surface():
    from RobustART.noise.utils.add_noise_utils import noise_list, default_config, function_dict
    assert noise_list == ['imagenet-s', 'imagenet-c', 'pgd_linf', 'pgd_l2', 'fgsm', 'autoattack_linf', 'mim_linf', 'pgd_l1']
    assert set(default_config) == set(noise_list) == set(function_dict)
    assert default_config['pgd_linf'] == {'f_model': None, 'eps': 8 / 255, 'rel_stepsize': 3 / 40, 'steps': 20}
    assert default_config['mim_linf']['step_size'] == 0.002 and default_config['imagenet-c']['severity'] == 1


def test_plugin_surface_matches_reference_sources():
    """noise_list / default_config / function_dict keys, the corruption id table and the model name table against
    tests/golden/plugin_surface.json, extracted from the reference's SOURCES by tests/golden/make_golden_surface.py
    (add_noise_utils.py:7-18,41-50; imagenet_c/__init__.py:5-8; utils/model_config.py)."""
    import json
    import os
    from RobustART.noise.utils.add_noise_utils import noise_list, default_config, function_dict
    from robustart_b200 import ops, solver as S
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "plugin_surface.json")))
    assert noise_list == g["noise_list"] and list(function_dict) == g["function_dict_keys"]
    assert {k: dict(v) for k, v in default_config.items()} == g["default_config"]
    assert list(ops.CORRUPTION_NAMES) == g["corruption_tuple"]            # ids are positions in corruption_tuple
    for name, cfg in S.model_name_dict.items():
        assert g["model_name_dict_types"][name] == cfg["type"], name
    # ImageNet-S 'train' transform: the random resized crop box (imagenet_s_gen.py:199-239), same draws from `random`
    import random
    from RobustART.noise.utils.add_noise_utils import random_resized_crop_params
    gp = g["get_params"]
    random.seed(gp["seed"])
    got = [list(random_resized_crop_params(h, w)) for h, w in gp["shapes"]]
    assert got == gp["params"]
    assert any(p[2] == h and p[3] != w or p[3] == w and p[2] != h for p, (h, w) in zip(gp["params"], gp["shapes"]))   # fallback branch hit


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


def test_result_file_mode_matches_reference_dump(tmp_path):
    """robustart_b200.resultfile against text produced by the reference's own ImageNetDataset.dump / BaseDataset.merge /
    ImageNetEvaluator.eval (tests/golden/make_golden_results.py): byte-identical lines (incl. the float("%.8f" % s) rounding
    traps: ties at the 9th decimal, the 1e-4 repr switch, denormals), same merged file, same top-k metrics."""
    import json
    import os
    import numpy as np
    from robustart_b200 import resultfile as R
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "result_lines.json")))
    scores = np.frombuffer(bytes.fromhex(g["scores_f32_bytes_hex"]), dtype=np.float32).reshape(g["shape"])
    pred, label = np.array(g["prediction"]), np.array(g["label"])
    fn, ids = ["val/a_%d.JPEG" % i for i in range(4)], [7, 8, 9, 10]
    assert R.format_lines(pred, label, scores, fn, ids) == g["text_pytorch"]
    assert R.format_lines(pred, label, scores) == g["text_dali"]
    # elementwise definition of the rounding
    r8 = R.round8(scores)
    assert all(float("%.8f" % s) == v for s, v in zip(scores.ravel().tolist(), r8.ravel().tolist()))
    with R.ResultWriter(str(tmp_path), 0) as w0, R.ResultWriter(str(tmp_path), 1) as w1:
        w0.write_batch(pred[:3], label[:3], scores[:3], fn[:3], ids[:3])
        w1.write_batch(pred[3:], label[3:], scores[3:], fn[3:], ids[3:])
    merged = R.merge(os.path.join(str(tmp_path), "results.txt.rank"), 2)
    assert os.path.basename(merged) == "results.txt.all" and open(merged).read() == g["merged"]
    assert R.evaluate(merged) == g["metric"]
    # the ImageNet-C file evaluator on the same file (imagenetc_evaluator.py:49-77): same numbers + the "metric" file
    from RobustART.metrics import ImageNetCEvaluator
    noise = os.path.join(str(tmp_path), "noise-fog-3-results.txt.all")
    open(noise, "w").write(g["merged"])
    m = ImageNetCEvaluator().eval(noise)
    assert m.metric == g["metric"] and m.v == g["metric"]["top1"]
    assert json.load(open(os.path.join(str(tmp_path), "noise-fog-3-metric"))) == g["metric"]


def test_ar_wcar_evaluators(tmp_path):
    """RobustART.metrics.{AdvRobustEvaluator, WorstCaseAdvRobustEvaluator} on hand-made result files
    (AR_evaluator.py:23-39, WCAR_evaluator.py:23-44): AR = kept / correct-before, WCAR = kept under every attack."""
    import json
    from RobustART.metrics import AdvRobustEvaluator, WorstCaseAdvRobustEvaluator

    def write(name, preds, labels, dali):
        p = tmp_path / name
        with open(p, "w") as f:
            for i, (a, b) in enumerate(zip(preds, labels)):
                d = {"prediction": a, "label": b, "score": [0.5, 0.5]}
                if not dali:
                    d = {"filename": "f%d.JPEG" % i, "image_id": i, **d}
                f.write(json.dumps(d) + "\n")
        return str(p)

    labels = [0, 1, 2, 3, 4, 5]
    clean = write("clean", [0, 1, 2, 3, 9, 9], labels, dali=False)       # 4 correct before
    a1 = write("a1", [0, 1, 7, 7, 4, 9], labels, dali=True)               # keeps 0, 1 (image 4 was wrong before: not counted)
    a2 = write("a2", [0, 8, 2, 7, 9, 9], labels, dali=False)              # keeps 0, 2
    assert AdvRobustEvaluator().eval(clean, a1) == 50.0
    assert AdvRobustEvaluator().eval(clean, a2) == 50.0
    assert WorstCaseAdvRobustEvaluator().eval(clean, [a1, a2]) == 25.0    # only image 0 survives both


def test_token_autograd_twins_match_reference_logits():
    """torch_models.ViT / Mixer (the autograd SOURCE models of attacks on BASELINE configs[2] / [4]) reproduce the golden logits
    of the reference's own classes (tests/golden/make_golden_models.py --tokens) on CPU, and give input gradients."""
    import os
    import numpy as np
    import torch
    from robustart_b200 import nets, torch_models
    from util import synth_images
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "token_logits.npz"))
    imgs = torch.from_numpy(synth_images(2, seed=9))
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    x = ((imgs.permute(0, 3, 1, 2).float().div(255) - mean) / std).requires_grad_(True)
    for arch, spec in (("vit_b16_224", nets.vit_spec()), ("mixer_b16_224", nets.mixer_spec())):
        m = torch_models.build(arch, nets.random_token_state_dict(spec, 0))
        y = m(x)
        assert np.abs(y.detach().numpy() - gold[arch]).max() < 1e-5, arch
        (g,) = torch.autograd.grad(torch.nn.functional.cross_entropy(y, torch.tensor([1, 2])), x)
        assert g.shape == x.shape and torch.isfinite(g).all() and g.abs().max() > 0
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "mobile_logits.npz"))
    imgs = torch.from_numpy(synth_images(2, seed=11))
    x = ((imgs.permute(0, 3, 1, 2).float().div(255) - mean) / std).requires_grad_(True)
    for arch in ("mobilenet_v2", "efficientnet_b0"):
        m = torch_models.build(arch, nets.random_state_dict(nets._MOBILE_ARCHS[arch][1](), 0))
        y = m(x)
        assert np.abs(y.detach().numpy() - gold[arch]).max() < 1e-6, arch
        # the synthetic weights attenuate the signal layer by layer (the float32 input gradient underflows to denormals):
        # differentiate the float64 copy
        xd = x.detach().double().requires_grad_(True)
        (g,) = torch.autograd.grad(m.double()(xd).square().sum(), xd)
        assert torch.isfinite(g).all() and g.abs().max() > 0
    from robustart_b200 import solver as S
    for name in ("mobilenet_v2_x1_0", "efficientnet_b0", "vit_base_patch16_224", "mixer_b16_224", "resnet18"):
        assert isinstance(S.build_torch_model({"type": S.model_name_dict[name]["type"]}, "", "cpu"), torch.nn.Module)


def test_file_dataset_host_logic(tmp_path, monkeypatch):
    """solver.FileImageNet (`read_from: fs`): meta-file parsing, PIL decode, per-image transform, labels, filenames
    (imagenet_dataset.py:55-67, imagenet_dataloader.py:74-80).  The GPU resize launch is replaced by the same Pillow call it is
    bit-exact with (tests/test_resize_gpu.py) so the host logic runs without a device."""
    import numpy as np
    import torch
    from PIL import Image
    from robustart_b200 import ops, solver as S

    def pil_transform(d, resize, crop, filter="bilinear"):
        arr = d[0].cpu().numpy()
        h, w = arr.shape[:2]
        oh, ow = (resize, int(resize * w / h)) if h <= w else (int(resize * h / w), resize)
        im = Image.fromarray(arr).resize((ow, oh), Image.BILINEAR)
        y0, x0 = int(round((oh - crop) / 2.0)), int(round((ow - crop) / 2.0))
        return torch.from_numpy(np.asarray(im.crop((x0, y0, x0 + crop, y0 + crop))).copy())[None]

    monkeypatch.setattr(ops, "resize_center_crop_u8", pil_transform)
    rs = np.random.RandomState(0)
    (tmp_path / "val").mkdir()
    sizes = [(90, 120), (150, 100), (64, 64), (80, 200)]
    lines = []
    for i, (h, w) in enumerate(sizes):
        Image.fromarray(rs.randint(0, 256, (h, w, 3)).astype(np.uint8)).save(tmp_path / "val" / ("img%d.png" % i))
        lines.append("img%d.png %d" % (i, 100 + i))
    (tmp_path / "meta.txt").write_text("\n".join(lines) + "\n")
    ds = S.FileImageNet(str(tmp_path / "val"), str(tmp_path / "meta.txt"), 56, "cpu", test_resize=64)
    assert ds.n == 4 and ds.filename(2) == "img2.png"
    imgs, labels = ds.batch(torch.tensor([3, 0]))
    assert imgs.shape == (2, 56, 56, 3) and imgs.dtype == torch.uint8 and labels.tolist() == [103, 100]
    want = pil_transform(torch.from_numpy(np.array(Image.open(tmp_path / "val" / "img3.png").convert("RGB")))[None], 64, 56)[0]
    assert torch.equal(imgs[0], want)
    assert S.FileImageNet(str(tmp_path / "val"), str(tmp_path / "meta.txt"), 56, "cpu", limit=3).n == 3
    # image_reader.type: opencv (image_reader.py:21-31: cv2.imdecode + BGR->RGB); ffmpeg is refused
    cv2 = pytest.importorskip("cv2")
    ds_cv = S.FileImageNet(str(tmp_path / "val"), str(tmp_path / "meta.txt"), 56, "cpu", test_resize=64, reader="opencv")
    assert torch.equal(ds_cv.batch(torch.tensor([3, 0]))[0], imgs)                       # PNG: both decoders give the same pixels
    jpg = str(tmp_path / "val" / "a.jpg")
    Image.fromarray(rs.randint(0, 256, (48, 64, 3)).astype(np.uint8)).save(jpg, quality=80)
    want_cv = cv2.cvtColor(cv2.imdecode(np.fromfile(jpg, dtype=np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
    assert np.array_equal(S.FileImageNet.decode(jpg, "opencv"), want_cv)
    assert np.array_equal(S.FileImageNet.decode(jpg, "pil"), np.array(Image.open(jpg).convert("RGB")))
    with pytest.raises(NotImplementedError):
        S.FileImageNet(str(tmp_path / "val"), str(tmp_path / "meta.txt"), 56, "cpu", reader="ffmpeg")


def test_robust_json_matches_reference_merge_eval_res(tmp_path):
    """resultfile.merge_imagenet_c_metrics and the arithmetic EvalSolver.evaluate_imagenet_c uses (mean top-1 error per type, means
    over types) against robust.json written by the reference's own ImageNetCDataset.merge_eval_res
    (tests/golden/make_golden_robust_json.py): same keys, same numbers."""
    import json
    import os
    from robustart_b200 import resultfile, solver as S
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "robust_json.json")))
    for name, m in g["metrics"].items():
        json.dump(m, open(tmp_path / name, "w"))
    got = resultfile.merge_imagenet_c_metrics(str(tmp_path))
    assert json.load(open(tmp_path / "robust.json")) == got
    want = g["robust"]
    assert list(got) == list(want) and all(list(got[k]) == list(want[k]) for k in want)
    for k in want:
        for t in want[k]:
            assert abs(got[k][t] - want[k][t]) < 1e-9, (k, t)
    assert resultfile.IMAGENET_C_GROUPS == S.EvalSolver.IMAGENET_C_GROUPS
    # the solver's in-loop formula: sum(errs) / len(errs) per type, then the same over types
    errs = {t: [100.0 - g["metrics"]["%s-%s-%d-metric" % (grp, t, s)]["top1"] for s in range(1, 6)]
            for grp, ts in S.EvalSolver.IMAGENET_C_GROUPS.items() for t in ts}
    per_type = {t: sum(e) / len(e) for t, e in errs.items()}
    assert abs(sum(per_type.values()) / len(per_type) - want["all"]["all_with_extra"]) < 1e-9
    wo = [per_type[t] for grp, ts in S.EvalSolver.IMAGENET_C_GROUPS.items() if grp != "extra" for t in ts]
    assert abs(sum(wo) / len(wo) - want["all"]["all_without_extra"]) < 1e-9


def test_round8_property():
    """resultfile.round8 == float("%.8f" % s) elementwise (the reference's score formatting, imagenet_dataset.py:262) on random
    float32 bit patterns in [0, 1], on values sitting exactly on / next to the 9th-decimal ties, and on softmax-like tails."""
    from robustart_b200 import resultfile as R
    rs = np.random.RandomState(1)
    bits = rs.randint(0, 0x3F800000, size=200000, dtype=np.int64).astype(np.uint32)        # every float32 in [0, 1)
    vals = [bits.view(np.float32)]
    ties = (np.arange(1, 4000, dtype=np.float64) * 2 + 1) * 0.5e-8                          # k + 0.5 in units of 1e-8
    vals.append(ties.astype(np.float32))
    vals.append(np.nextafter(ties.astype(np.float32), np.float32(1)))
    vals.append(np.nextafter(ties.astype(np.float32), np.float32(0)))
    vals.append(np.exp(rs.uniform(-40, 0, 50000)).astype(np.float32))
    v = np.concatenate(vals)
    got = R.round8(v)
    want = np.array([float("%.8f" % s) for s in v.tolist()])
    assert np.array_equal(got, want)
    assert R.round8(np.float32(0.5)).item() == 0.5 and R.round8(np.array([[1.0, 0.0]], np.float32)).tolist() == [[1.0, 0.0]]


def test_missing_checkpoint_raises(tmp_path):
    """a non-empty path that does not exist must raise (the reference's torch.load does), never fall back to
    synthetic weights; no path / 'synthetic' is the explicit opt-in."""
    import torch
    from robustart_b200 import solver as S
    assert S.load_checkpoint(None) is None and S.load_checkpoint("") is None and S.load_checkpoint("synthetic") is None
    with pytest.raises(FileNotFoundError):
        S.load_checkpoint(str(tmp_path / "nope.pth.tar"))
    torch.save({"model": {"a": torch.ones(2)}}, tmp_path / "ok.pth.tar")
    assert "model" in S.load_checkpoint(str(tmp_path / "ok.pth.tar"))


def test_noise_stream_ids_are_positions_in_the_permutation():
    """solver.stream_base + position: every image gets a unique RNG stream that does not depend on the number of ranks (ADVICE r1:
    permuted shard indices used as contiguous offsets let different images share a Philox stream)."""
    import math
    from robustart_b200.solver import shard_indices
    n, ref = 103, None
    for world in (1, 2, 3, 8):
        seen = {}
        for r in range(world):
            base = int(math.ceil(n / world)) * r
            for j, i in enumerate(shard_indices(n, world, r).tolist()):
                seen[i] = base + j
        assert sorted(seen.values()) == list(range(n))
        assert ref is None or seen == ref
        ref = seen
