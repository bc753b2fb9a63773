"""Reference restatement of the eval reduction -- TEST INFRASTRUCTURE ONLY.

accuracy(): prototype/prototype/utils/misc.py:441-455
ImageNetEvaluator.eval(): prototype/prototype/data/metrics/imagenet_evaluator.py:49-67
DistributedSampler(round_up=False): prototype/prototype/data/sampler.py:8-52
"""
import math

import torch


def accuracy(output, target, topk=(1,)):
    """precision@k in percent, one 1-element float tensor per k (misc.py:441-455): a sample counts for k when its label is
    among the first k indices torch.topk returns (largest, sorted) -- ties are resolved by torch.topk, as in the reference."""
    order = output.topk(max(topk), dim=1, largest=True, sorted=True).indices
    hit = order == target.reshape(-1, 1)
    n = target.shape[0]
    return [hit[:, :k].any(dim=1).float().sum().reshape(1) * (100.0 / n) for k in topk]


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
