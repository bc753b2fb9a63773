"""Reference restatement of the eval reduction -- TEST INFRASTRUCTURE ONLY.

accuracy(): prototype/prototype/utils/misc.py:441-455
ImageNetEvaluator.eval(): prototype/prototype/data/metrics/imagenet_evaluator.py:49-67
DistributedSampler(round_up=False): prototype/prototype/data/sampler.py:8-52
"""
import math

import torch


def accuracy(output, target, topk=(1,)):
    maxk = max(topk)
    batch_size = target.size(0)
    _, pred = output.topk(maxk, 1, True, True)
    pred = pred.t()
    correct = pred.eq(target.reshape(1, -1).expand_as(pred))
    res = []
    for k in topk:
        correct_k = correct[:k].reshape(-1).float().sum(0, keepdim=True)
        res.append(correct_k.mul_(100.0 / batch_size))
    return res


def topk_hits(output, target, ks=(1, 5)):
    """integer hit counts (the quantity our counters hold)"""
    maxk = max(ks)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.eq(target.view(-1, 1))
    return [int(correct[:, :k].any(dim=1).sum()) for k in ks]


def sampler_indices(n_items, world_size, rank, epoch=0):
    """sampler.py:20-46 with round_up=False."""
    num_samples = int(math.ceil(n_items * 1.0 / world_size))
    g = torch.Generator()
    g.manual_seed(epoch)
    indices = torch.randperm(n_items, generator=g).tolist()
    offset = num_samples * rank
    return indices[offset:offset + num_samples]
