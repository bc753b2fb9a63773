"""Evaluation solvers of the hot path: the reference's `cls_solver --evaluate` (clean / ImageNet-C) and
`*_benchmark_eval_adv` (adversarial) command lines, re-hosted on the B200 pipeline.

Kept from the reference (prototype/prototype/solver/cls_solver.py:352-457,460-481;
benchmark_eval_adv.py:191-299): CLI flags, YAML schema (model / model_src / model_tgt, data.*, saver.pretrain.*),
`--eps` as a Python expression, env-var rank discovery (dist.py:21-54, linklink/__init__.py:21-34), the
DistributedSampler(round_up=False) sharding rule (sampler.py:8-52), checkpoint key handling.
Replaced: the per-image JSON dump + file merge + re-parse (imagenet_dataset.py:250-277, base_dataset.py:116-133,
imagenet_evaluator.py:49-67) by device counters and ONE all-reduce; the CPU DataLoader by device-resident
synthetic batches (`data.read_from: fake|synthetic`) -- JPEG decode / DALI / memcached readers are outside
the hot path and raise NotImplementedError.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, Optional

import torch
import yaml

from . import nets, ops


class AttrDict(dict):
    """Tiny EasyDict stand-in (the reference depends on `easydict`, misc.py:15)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        elif isinstance(v, list):
            v = [AttrDict(x) if isinstance(x, dict) else x for x in v]
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def parse_config(config_file: str) -> AttrDict:  # misc.py:67-72
    with open(config_file) as f:
        return AttrDict(yaml.load(f, Loader=yaml.FullLoader))


# name -> model config, the subset of prototype/prototype/utils/model_config.py:3-288 with a B200 kernel path
model_name_dict = {
    "resnet18": {"type": "resnet18_official", "kwargs": {"bn": {"use_sync_bn": False, "kwargs": {}}}},
    "resnet34": {"type": "resnet34_official", "kwargs": {"bn": {"use_sync_bn": False, "kwargs": {}}}},
    "resnet50": {"type": "resnet50_official", "kwargs": {"bn": {"use_sync_bn": False, "kwargs": {}}}},
    "resnet101": {"type": "resnet101_official", "kwargs": {"bn": {"use_sync_bn": False, "kwargs": {}}}},
    "mobilenet_v2_x1_0": {"type": "mobilenet_v2", "kwargs": {"scale": 1.0, "bn": {"use_sync_bn": False, "kwargs": {}}}},
    "efficientnet_b0": {"type": "efficientnet_b0", "kwargs": {"bn": {"use_sync_bn": False, "kwargs": {}}}},
    "vit_base_patch16_224": {"type": "vit_b16_224", "kwargs": {"drop_path": 0.0, "dropout": 0.0, "attention_dropout": 0.0,
                                                               "qkv_bias": True, "representation_size": 768}},
    "mixer_b16_224": {"type": "mixer_b16_224", "kwargs": {"drop_path": 0.0, "drop_path_rate": 0.0}},
}


# ------------------------------------------------------------------------------------------------
# distributed bring-up (dist.py:21-54): one process per GPU, NCCL; SKIP_DIST=1 -> single process
# ------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self, rank=0, world_size=1, local_rank=0, initialized=False):
        self.rank, self.world_size, self.local_rank, self.initialized = rank, world_size, local_rank, initialized


def dist_init(backend: Optional[str] = None) -> Dist:
    import torch.distributed as dist
    if os.environ.get("SKIP_DIST", "0") == "1":
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        return Dist()
    if "SLURM_PROCID" in os.environ and "RANK" not in os.environ:
        rank, world = int(os.environ["SLURM_PROCID"]), int(os.environ["SLURM_NTASKS"])
    else:
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank % max(torch.cuda.device_count(), 1))))
    if world == 1:
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        return Dist(rank, world, local)
    backend = backend or os.environ.get("B200R_DIST_BACKEND") or ("nccl" if torch.cuda.is_available() else "gloo")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend, rank=rank, world_size=world)
    return Dist(rank, world, local, True)


def shard_indices(n_items: int, world_size: int, rank: int, epoch: int = 0):
    """DistributedSampler(round_up=False).__iter__ (sampler.py:36-46): one seeded permutation, rank r takes
    [r*ceil(N/W), (r+1)*ceil(N/W)), the last rank the remainder; nothing is padded or duplicated."""
    num_samples = int(math.ceil(n_items * 1.0 / world_size))
    g = torch.Generator()
    g.manual_seed(epoch)
    indices = torch.randperm(n_items, generator=g)
    return indices[num_samples * rank: num_samples * rank + num_samples]


def reduce_counters(counters: torch.Tensor, d: Dist) -> torch.Tensor:
    """The single collective of an evaluation: int64 [.., 3] hit counters summed over ranks."""
    if d.initialized and d.world_size > 1:
        import torch.distributed as dist
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


# ------------------------------------------------------------------------------------------------
class SyntheticImageNet:
    """Device-resident stand-in for the validation set (`read_from: fake` reuses one decoded file for
    every sample in the reference, base_dataset.py:73-80; here every index gets its own seeded image).
    Items are generated on the fly from their global index, so any sharding sees the same data."""

    def __init__(self, n_items: int, input_size: int, device, seed: int = 0, raw_size=None, test_resize: int = 256):
        self.n, self.size, self.device, self.seed = n_items, input_size, device, seed
        # raw_size = (h, w): items are "decoded files" of that size and go through the eval transform on the GPU --
        # Resize(test_resize) + CenterCrop(input_size), imagenet_dataloader.py:74-80 -- instead of being generated at input_size
        self.raw_size = tuple(int(v) for v in raw_size) if raw_size else None
        self.test_resize = int(test_resize)

    def batch(self, indices: torch.Tensor):
        n = len(indices)
        h, w = self.raw_size or (self.size, self.size)
        imgs = torch.empty((n, h, w, 3), dtype=torch.uint8, device=self.device)
        labels = torch.empty(n, dtype=torch.int64, device=self.device)
        g = torch.Generator(device=self.device)
        for j, idx in enumerate(indices.tolist()):
            g.manual_seed(self.seed * 1000003 + idx)
            imgs[j] = torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device=self.device, generator=g)
            labels[j] = idx % 1000
        if self.raw_size:
            imgs = ops.resize_center_crop_u8(imgs, self.test_resize, self.size)
        return imgs, labels


class FileImageNet:
    """`read_from: fs` (imagenet_dataset.py:55-67, base_dataset.py:55-72): `meta_file` lines "filename label" under `root_dir`.
    Files are decoded on the host with PIL (`image_reader.type: pil`, image_reader.py:11-18 -- file parsing is outside the
    hot path, as in the reference) and go through the eval transform on the GPU: Resize(test_resize) + CenterCrop(input_size)
    as ONE bit-exact Pillow-resize launch per image (imagenet_dataloader.py:74-80; ops.resize_center_crop_u8)."""

    def __init__(self, root_dir: str, meta_file: str, input_size: int, device, test_resize: int = 256, limit: Optional[int] = None,
                 reader: str = "pil"):
        if reader not in ("pil", "opencv"):
            raise NotImplementedError("image_reader.type=%r: pil and opencv are implemented (image_reader.py:11-31); ffmpeg is not" % reader)
        self.reader = reader
        self.root, self.size, self.device, self.test_resize = root_dir, input_size, device, int(test_resize)
        self.metas = []
        with open(meta_file) as f:
            for line in f:
                if line.strip():
                    filename, label = line.rstrip().split()
                    self.metas.append((filename, int(label)))
        if limit is not None:
            self.metas = self.metas[:int(limit)]
        self.n = len(self.metas)

    def filename(self, idx: int) -> str:
        return self.metas[idx][0]

    @staticmethod
    def decode(path: str, reader: str = "pil"):
        """pil_loader / opencv_loader of image_reader.py:11-31: a uint8 RGB [h, w, 3] array (host side, like the reference)."""
        import numpy as np
        if reader == "opencv":
            import cv2
            return cv2.cvtColor(cv2.imdecode(np.fromfile(path, dtype=np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
        from PIL import Image
        with Image.open(path) as im:
            return np.array(im.convert("RGB"))

    def batch(self, indices: torch.Tensor):
        imgs, labels = [], []
        for idx in indices.tolist():
            name, label = self.metas[idx]
            arr = self.decode(os.path.join(self.root, name), self.reader)
            d = torch.from_numpy(arr)[None].to(self.device)
            imgs.append(ops.resize_center_crop_u8(d, self.test_resize, self.size)[0])
            labels.append(label)
        return torch.stack(imgs), torch.tensor(labels, dtype=torch.int64, device=self.device)


def load_checkpoint(path: Optional[str]) -> Optional[Dict[str, torch.Tensor]]:
    """None (-> synthetic weights) only when NO path is given or it is the literal 'synthetic'; a path that does not exist
    raises like the reference's torch.load does (benchmark_eval_adv.py:75-79), so a mistyped --src_path can never be
    evaluated as a random-weight model."""
    if not path or path == "synthetic":
        return None
    if not os.path.exists(path):
        raise FileNotFoundError("checkpoint %r does not exist (pass no path or 'synthetic' for synthetic weights)" % path)
    return torch.load(path, map_location="cpu")


def build_b200_model(model_cfg, ckpt_path, device):
    """model_entry (model/__init__.py:292-332) for the families with a kernel path."""
    mtype = model_cfg["type"]
    sd = load_checkpoint(ckpt_path)
    return nets.build_model(mtype, sd, device=device)


def build_source_model(model_cfg, ckpt_path, device):
    """The SOURCE model of an attack (needs input gradients): ResNets run forward + dgrad on the B200 kernels
    (attacks.NativeModel); set B200R_SOURCE_AUTOGRAD=1 to get the torch.autograd twin instead."""
    from .attacks import NativeModel
    arch = nets.ARCH_ALIASES.get(model_cfg["type"], model_cfg["type"])
    if arch in nets._RESNET_CFG and os.environ.get("B200R_SOURCE_AUTOGRAD", "0") != "1":
        return NativeModel(build_b200_model(model_cfg, ckpt_path, device))
    if (arch in nets._TOKEN_ARCHS or arch in nets._MOBILE_ARCHS) and os.environ.get("B200R_SOURCE_AUTOGRAD", "0") != "1":
        # forward + input gradient of ViT / Mixer on our kernels (token_backward.cu + dgrad GEMMs; 4.2x / 1.65x the autograd twin
        # on the PGD-Linf loop when first measured, profiles/r2_pgd_token.json)
        return NativeModel(build_b200_model(model_cfg, ckpt_path, device))
    return build_torch_model(model_cfg, ckpt_path, device)


def build_torch_model(model_cfg, ckpt_path, device):
    """An autograd-capable nn.Module twin for the source model of an attack: the ResNet family, the token models
    (ViT-B/16, MLP-Mixer-B/16 -- BASELINE configs[2] / [4]) and the mobile families (MobileNetV2, EfficientNet-B0);
    same state_dict keys as the reference's classes."""
    from . import torch_models
    arch = nets.ARCH_ALIASES.get(model_cfg["type"], model_cfg["type"])
    sd = load_checkpoint(ckpt_path)
    if isinstance(sd, dict) and "model" in sd and isinstance(sd["model"], dict):
        sd = sd["model"]
    if arch in torch_models._TOKEN:
        if sd is None:      # synthetic weights: the same ones nets.build_model(arch) generates for the kernel model
            sd = (nets.random_token_state_dict(nets._TOKEN_ARCHS[arch][1](), 0) if arch in nets._TOKEN_ARCHS
                  else nets.random_state_dict(nets._MOBILE_ARCHS[arch][1](), 0))
        return torch_models.build(arch, nets._strip_prefix(sd)).to(device).eval()
    if arch not in torch_models._CFG:
        raise NotImplementedError("source model %r: the ResNet, ViT, MLP-Mixer, MobileNetV2 and EfficientNet-B0 families have "
                                  "autograd twins; use it as the target (forward-only) model or pass your own nn.Module to AddNoise" % arch)
    if sd is None:
        sd = nets.random_state_dict(nets.resnet_spec(arch), 0)
    return torch_models.build(arch, nets._strip_prefix(sd)).to(device).eval()


# ------------------------------------------------------------------------------------------------
class EvalSolver:
    def __init__(self, config: AttrDict, prefix: str = "", dist_info: Optional[Dist] = None):
        self.config = config
        self.dist = dist_info or dist_init()
        if not torch.cuda.is_available():
            raise RuntimeError("the evaluation solvers need a CUDA device (as the reference does, cls_solver.py:108)")
        self.device = torch.device("cuda", torch.cuda.current_device())
        data = config.data
        read_from = data.get("read_from", "fake")
        if read_from not in ("fake", "synthetic", "fs", "file"):
            raise NotImplementedError("data.read_from=%r: only `fake` (synthetic tensors) and `fs` (files decoded with PIL on the "
                                      "host, transformed on the GPU) are implemented; memcached / petrel / DALI readers are "
                                      "outside the B200 hot path" % read_from)
        self.batch_size = int(data.get("batch_size", 64))
        self.input_size = int(data.get("input_size", 224))
        test = data.get("test", {})
        if read_from in ("fs", "file"):
            self.dataset = FileImageNet(test["root_dir"], test["meta_file"], self.input_size, self.device,
                                        test_resize=int(data.get("test_resize", 256)), limit=test.get("limit_samples"),
                                        reader=(test.get("image_reader") or {}).get("type", "pil"))
            n_items = self.dataset.n
        else:
            n_items = int(test.get("limit_samples", data.get("num_samples", 50000)))
            self.dataset = SyntheticImageNet(n_items, self.input_size, self.device, raw_size=data.get("raw_size"),
                                             test_resize=int(data.get("test_resize", 256)))
        self.indices = shard_indices(n_items, self.dist.world_size, self.dist.rank)
        # RNG stream of an image = its POSITION in the epoch's permutation (this rank's slice starts at rank * ceil(N / W)): unique per
        # image, and the same whatever the batch size or the number of ranks -- any sharding corrupts an image with the same noise
        self.stream_base = int(math.ceil(n_items * 1.0 / self.dist.world_size)) * self.dist.rank
        self.result_path = os.path.join(config.get("save_path", "."), prefix, "results")
        if self.dist.rank == 0:
            os.makedirs(self.result_path, exist_ok=True)

    def _filenames(self, ids):
        fn = getattr(self.dataset, "filename", None)
        return [fn(i) if fn else "synthetic/%08d.JPEG" % i for i in ids]

    def _batches(self):
        for i in range(0, len(self.indices), self.batch_size):
            yield self.dataset.batch(self.indices[i:i + self.batch_size])

    def _finish(self, counters, tag="results"):
        reduce_counters(counters, self.dist)
        c = counters.tolist()
        metric = {"top1": 100.0 * c[0] / max(c[2], 1), "top5": 100.0 * c[1] / max(c[2], 1), "count": c[2]}
        if self.dist.rank == 0:
            with open(os.path.join(self.result_path, tag + ".metrics.json"), "w") as f:
                json.dump(metric, f, indent=2)
            print(json.dumps(metric, indent=2))
        return metric

    # cls_solver.py:352-457
    def evaluate(self, model, corruption=None, severity=1):
        """Counters on the device + one all-reduce.  With `data.test.dump_results: true` the reference's result files are
        written as well (resultfile.py: `<tag>.txt.rank{r}` per-image lines with the softmax scores, merged `.all`,
        file-based top-k next to the counter metric) -- the opt-in compatibility mode for downstream parsers."""
        from RobustART.noise.utils import add_noise_utils as anu
        counters = torch.zeros(3, dtype=torch.int64, device=self.device)
        tag = "results" if corruption is None else "noise-%s-%d-results" % (corruption, severity)
        dump = bool(self.config.data.get("test", {}).get("dump_results", False))
        writer = None
        if dump:
            from . import resultfile
            writer = resultfile.ResultWriter(self.result_path, self.dist.rank, stem=tag + ".txt")
            pred = torch.empty(self.batch_size, dtype=torch.int64, device=self.device)
        done = 0
        for imgs, labels in self._batches():
            n = imgs.shape[0]
            if corruption is not None:
                ops.corrupt_u8(imgs, corruption, severity, seed=anu._seed(), image_offset=self.stream_base + done, out=imgs)
            logits = model(imgs)
            if writer is None:
                ops.topk_count_(counters, logits, labels)
            else:
                ops.topk_count_(counters, logits, labels, pred[:n])
                ids = self.indices[done:done + n].tolist()
                writer.write_batch(pred[:n].cpu().numpy(), labels.cpu().numpy(), ops.softmax(logits).cpu().numpy(),
                                   self._filenames(ids), ids)
            done += n
        metric = self._finish(counters, tag)
        if writer is not None:
            from . import resultfile
            writer.close()
            if self.dist.initialized and self.dist.world_size > 1:
                import torch.distributed as dist
                dist.barrier()
            if self.dist.rank == 0:
                merged = resultfile.merge(os.path.join(self.result_path, tag + ".txt.rank"), self.dist.world_size)
                metric["result_file"] = merged
                metric["file_metric"] = resultfile.evaluate(merged)
        return metric

    # imgnet_c_solver.py:374-470 + datasets/imagnetc.py:165-218
    IMAGENET_C_GROUPS = {
        "noise": ["gaussian_noise", "shot_noise", "impulse_noise"],
        "blur": ["defocus_blur", "glass_blur", "motion_blur", "zoom_blur"],
        "weather": ["snow", "frost", "fog", "brightness"],
        "digital": ["contrast", "elastic_transform", "pixelate", "jpeg_compression"],
        "extra": ["speckle_noise", "spatter", "gaussian_blur", "saturate"],
    }

    def evaluate_imagenet_c(self, model, groups=None, severities=(1, 2, 3, 4, 5)):
        """The reference's ImageNet-C sweep: every (corruption, severity) cell over this rank's shard.  The reference
        reads 95 pre-corrupted copies of the validation set and shards the 95 *metric files* over ranks; here every
        rank corrupts its own image shard on the GPU (the clean batch is generated / loaded once per batch and
        reused for all cells), counters live in one int64 [cells, 3] tensor and ONE all-reduce ends the sweep.
        Writes `{group}-{type}-{sev}-metric` files and `robust.json` (mean top-1 error per type, 'all_with_extra',
        'all_without_extra') exactly as merge_eval_res does."""
        from RobustART.noise.utils import add_noise_utils as anu
        groups = groups or self.IMAGENET_C_GROUPS
        cells = [(g, t, s) for g in groups for t in groups[g] for s in severities]
        counters = torch.zeros((len(cells), 3), dtype=torch.int64, device=self.device)
        skipped = {}
        done = 0
        work = None
        for imgs, labels in self._batches():
            if work is None or work.shape != imgs.shape:
                work = torch.empty_like(imgs)
            for ci, (g, t, s) in enumerate(cells):
                if (t, s) in skipped:
                    continue
                try:
                    ops.corrupt_u8(imgs, t, s, seed=anu._seed(), image_offset=self.stream_base + done, out=work)
                except NotImplementedError as e:                 # a cell without a kernel yet is reported, not faked
                    skipped[(t, s)] = str(e)
                    continue
                ops.topk_count_(counters[ci], model(work), labels)
            done += imgs.shape[0]
        reduce_counters(counters, self.dist)
        table = counters.tolist()
        all_data = {"all": {}}
        avg, avg_wo = [], []
        for g in groups:
            all_data[g] = {}
            for t in groups[g]:
                errs = []
                for s in severities:
                    c = table[cells.index((g, t, s))]
                    if (t, s) in skipped or c[2] == 0:
                        continue
                    m = {"top1": 100.0 * c[0] / c[2], "top5": 100.0 * c[1] / c[2], "count": c[2]}
                    errs.append(100.0 - m["top1"])
                    if self.dist.rank == 0:
                        with open(os.path.join(self.result_path, "%s-%s-%d-metric" % (g, t, s)), "w") as f:
                            json.dump(m, f)
                # a type with a skipped severity has no number comparable with merge_eval_res (imagnetc.py:166-218): None, and
                # the two averages below become None as well
                partial = any((t, s) in skipped for s in severities)
                all_data[g][t] = sum(errs) / len(errs) if (errs and not partial) else None
                if errs and not partial:
                    avg.append(all_data[g][t])
                    if g != "extra":
                        avg_wo.append(all_data[g][t])
        skipped_types = {t for t, _ in skipped}
        extra_types = set(groups.get("extra", []))
        all_data["all"]["all_with_extra"] = sum(avg) / len(avg) if (avg and not skipped_types) else None
        all_data["all"]["all_without_extra"] = sum(avg_wo) / len(avg_wo) if (avg_wo and not (skipped_types - extra_types)) else None
        if skipped:
            all_data["skipped"] = {"%s-%d" % k: v for k, v in skipped.items()}
        if self.dist.rank == 0:
            with open(os.path.join(self.result_path, "robust.json"), "w") as f:
                json.dump(all_data, f, indent=1)
            print(json.dumps(all_data["all"]))
        return all_data

    # benchmark_eval_adv.py:191-254
    def evaluate_adv(self, model_src, model_tgt, attack="none", eps=0.0):
        """benchmark_eval_adv.py:191-254.  Besides top-1 / top-5 after the attack the device counters carry the benchmark's
        Adversarial Robustness (RobustART/metrics/AR_evaluator.py:23-39): correct before AND after the attack over correct
        before -- one extra clean forward of the target model per batch, no result files.  `data.test.dump_results: true`
        also writes `results.txt.*` (after the attack) and `clean-results.txt.*` in the reference's line format, which
        RobustART.metrics.AdvRobustEvaluator / WorstCaseAdvRobustEvaluator read."""
        from RobustART.noise import AddNoise
        from .attacks import NativeModel, PyTorchModel
        counters = torch.zeros(3, dtype=torch.int64, device=self.device)
        ar = torch.zeros(2, dtype=torch.int64, device=self.device)       # [correct before, correct before and after]
        gen = None
        if attack in ("autoattack_linf", "mim_linf", "pgd_l1"):
            gen = AddNoise(attack)
            gen.set_config(model=model_src, eps=eps)
        elif attack != "none":
            f_model = model_src if isinstance(model_src, NativeModel) else PyTorchModel(
                model_src, bounds=(0, 1), preprocessing=dict(mean=ops.IMAGENET_MEAN, std=ops.IMAGENET_STD, axis=-3))
            gen = AddNoise(attack)
            gen.set_config(f_model=f_model, eps=eps)
        dump = bool(self.config.data.get("test", {}).get("dump_results", False))
        w_adv = w_clean = None
        if dump:
            from . import resultfile
            w_adv = resultfile.ResultWriter(self.result_path, self.dist.rank, stem="results.txt")
            w_clean = resultfile.ResultWriter(self.result_path, self.dist.rank, stem="clean-results.txt")
        pred_c = torch.empty(self.batch_size, dtype=torch.int64, device=self.device)
        pred_a = torch.empty(self.batch_size, dtype=torch.int64, device=self.device)
        scratch = torch.zeros(3, dtype=torch.int64, device=self.device)
        done = 0
        for imgs, labels in self._batches():
            n = imgs.shape[0]
            x01 = ops.normalize(ops.u8nhwc_to_f32nchw(imgs), "inv")   # the loader normalises, the solver undoes it (:229)
            logits_c = model_tgt(x01)
            ops.topk_count_(scratch, logits_c, labels, pred_c[:n])
            if dump:
                ids = self.indices[done:done + n].tolist()
                names = self._filenames(ids)
                w_clean.write_batch(pred_c[:n].cpu().numpy(), labels.cpu().numpy(), ops.softmax(logits_c).cpu().numpy(), names, ids)
            if gen is not None:
                x01 = gen.add_noise(x01, labels).contiguous()
            logits = model_tgt(x01)                                   # normalisation fused into the stem gather
            ops.topk_count_(counters, logits, labels, pred_a[:n])
            ok_c = pred_c[:n] == labels
            ar += torch.stack([ok_c.sum(), (ok_c & (pred_a[:n] == labels)).sum()])
            if dump:
                w_adv.write_batch(pred_a[:n].cpu().numpy(), labels.cpu().numpy(), ops.softmax(logits).cpu().numpy(), names, ids)
            done += n
        reduce_counters(ar, self.dist)
        metric = self._finish(counters, "results")
        a = ar.tolist()
        metric["clean_top1"] = 100.0 * a[0] / max(metric["count"], 1)
        metric["AR"] = 100.0 * a[1] / max(a[0], 1)
        if dump:
            from . import resultfile
            w_adv.close(); w_clean.close()
            if self.dist.initialized and self.dist.world_size > 1:
                import torch.distributed as dist
                dist.barrier()
            if self.dist.rank == 0:
                metric["result_file"] = resultfile.merge(os.path.join(self.result_path, "results.txt.rank"), self.dist.world_size)
                metric["clean_result_file"] = resultfile.merge(os.path.join(self.result_path, "clean-results.txt.rank"), self.dist.world_size)
        if self.dist.rank == 0:
            with open(os.path.join(self.result_path, "results.metrics.json"), "w") as f:
                json.dump(metric, f, indent=2)
        return metric
