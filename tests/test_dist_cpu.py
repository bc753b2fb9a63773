"""Multi-rank host logic on CPU (gloo, world_size 2): sharding rule + the single counter all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), B200R_DIST_BACKEND="gloo")
    os.environ.pop("SKIP_DIST", None)
    from robustart_b200 import solver as S
    d = S.dist_init("gloo")
    idx = S.shard_indices(n_items, d.world_size, d.rank)
    # stand-in for the device counters: "hits" are a deterministic function of the index
    counters = torch.tensor([int((idx % 7 == 0).sum()), int((idx % 3 == 0).sum()), len(idx)], dtype=torch.int64)
    S.reduce_counters(counters, d)
    if rank == 0:
        torch.save({"counters": counters, "len0": len(idx)}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_counter_allreduce(tmp_path):
    n = 1001
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
    r = torch.load(out)
    allidx = torch.arange(n)
    assert r["counters"].tolist() == [int((allidx % 7 == 0).sum()), int((allidx % 3 == 0).sum()), n]
    assert r["len0"] == 501


def test_shard_rule_matches_reference_sampler():
    from robustart_b200 import solver as S
    from oracle import metrics as OM
    for n, w in [(50000, 8), (50001, 8), (10, 4), (7, 1)]:
        for r in range(w):
            assert S.shard_indices(n, w, r).tolist() == OM.sampler_indices(n, w, r)


def test_skip_dist_env(monkeypatch):
    from robustart_b200 import solver as S
    monkeypatch.setenv("SKIP_DIST", "1")
    d = S.dist_init()
    assert (d.rank, d.world_size, d.initialized) == (0, 1, False)


def test_config_and_cli_surface(tmp_path):
    from robustart_b200 import solver as S
    import prototype.prototype.solver.benchmark_eval_adv as adv
    import prototype.prototype.solver.cls_solver as cls
    cfg = S.parse_config(os.path.join(os.path.dirname(os.path.dirname(__file__)), "exprs", "b200", "resnet50_eval.yaml"))
    assert cfg.model.type == "resnet50_official" and cfg.data.test.evaluator.kwargs.topk == [1, 5]
    assert S.model_name_dict["resnet50"]["type"] == "resnet50_official"
    import glob
    cfgs = sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(__file__)), "exprs", "b200", "config*.yaml")))
    assert len(cfgs) == 5                          # one per BASELINE config
    for p in cfgs:
        c = S.parse_config(p)
        assert c.data.read_from == "fake" and c.model.type in [v["type"] for v in S.model_name_dict.values()]
    with pytest.raises(SystemExit):
        adv.main(["--config", "x.yaml"])          # the reference's required flags are required here too
    with pytest.raises(SystemExit):
        cls.main(["--config", "x.yaml"])          # training is not part of the hot path


def _rf_worker(rank, world, port, n_items, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), B200R_DIST_BACKEND="gloo")
    os.environ.pop("SKIP_DIST", None)
    import numpy as np
    from robustart_b200 import resultfile, solver as S
    d = S.dist_init("gloo")
    idx = S.shard_indices(n_items, d.world_size, d.rank).numpy()
    rs = np.random.RandomState(0)
    scores_all = rs.dirichlet(np.ones(10), size=n_items).astype(np.float32)        # same table on every rank
    with resultfile.ResultWriter(out_dir, d.rank) as w:
        for i in range(0, len(idx), 4):
            j = idx[i:i + 4]
            w.write_batch(scores_all[j].argmax(1), j % 10, scores_all[j], ["f%05d.JPEG" % k for k in j], j.tolist())
    dist.barrier()
    if rank == 0:
        resultfile.merge(os.path.join(out_dir, "results.txt.rank"), d.world_size)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_result_files_merge(tmp_path):
    """Result-file mode across ranks: every rank dumps its shard (base_dataset.py:116-133 layout), rank 0 concatenates the
    rank files in rank order, the file evaluator sees every image exactly once."""
    import json
    import numpy as np
    from robustart_b200 import resultfile, solver as S
    n = 37
    mp.spawn(_rf_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    merged = open(tmp_path / "results.txt.all").read()
    assert merged == open(tmp_path / "results.txt.rank0").read() + open(tmp_path / "results.txt.rank1").read()
    lines = [json.loads(l) for l in merged.splitlines()]
    order = S.shard_indices(n, 2, 0).tolist() + S.shard_indices(n, 2, 1).tolist()
    assert [l["image_id"] for l in lines] == order and sorted(order) == list(range(n))
    scores = np.random.RandomState(0).dirichlet(np.ones(10), size=n).astype(np.float32)
    top1 = 100.0 * np.mean([scores[i].argmax() == i % 10 for i in range(n)])
    m = resultfile.evaluate(str(tmp_path / "results.txt.all"), topk=(1, 5))
    assert abs(m["top1"] - top1) < 1e-4
