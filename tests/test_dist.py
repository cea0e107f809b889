"""World-size-2 test of the sharding / exchange logic on CPU (gloo).  The kernels are replaced by the kernel-logic
emulator; what is under test is pbnn_b200/dist.py: contiguous sharding, world-size independent keys, and the pooled
predictive moments (the single exchange step of SURVEY.md 8e)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, out_dir):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from emu_ops import EmuOps
    from oracle import pbnn_oracle as o
    from pbnn_b200 import dist as pdist
    from pbnn_b200.ops import FlatSpec
    ops = EmuOps()
    C, layers = 5, (8, 12, 1)  # 5 chains over 2 ranks: ragged shards 3 + 2
    X, y, Xt, _ = o.synthetic_regression(120, 8, 0, M=40)
    fs, sp = FlatSpec(layers, 0, 0.1), o.MlpSpec(layers, 0, 0.1)
    lo, hi = pdist.shard_range(C)
    keys = pdist.shard_chain_keys(ops, o.PRNGKey(3), C)
    all_keys = o.split(o.PRNGKey(3), C)
    assert np.array_equal(keys.numpy().view(np.uint32), all_keys[lo:hi])
    th0 = o.init_positions(sp, all_keys)[lo:hi]
    res = ops.sgld(fs, X, y, th0, keys, 16, 1e-5, 4)
    mean, var, count = pdist.predictive_moments_sharded(ops, fs, res["theta"], Xt)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), theta=res["theta"].numpy(), mean=mean.numpy(), var=var.numpy(),
             count=count, lo=lo, hi=hi)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partition():
    from pbnn_b200.dist import shard_range
    for C in (0, 1, 5, 256, 8192):
        for W in (1, 2, 3, 8):
            r = [shard_range(C, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == C
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_two_rank_sharding_matches_single_process(tmp_path, emu_ops):
    from oracle import pbnn_oracle as o
    from pbnn_b200.ops import FlatSpec
    port = 29500 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    r = [np.load(tmp_path / f"r{k}.npz") for k in range(2)]
    assert (int(r[0]["lo"]), int(r[0]["hi"]), int(r[1]["lo"]), int(r[1]["hi"])) == (0, 3, 3, 5)
    # single-process run of all 5 chains
    C, layers = 5, (8, 12, 1)
    X, y, Xt, _ = o.synthetic_regression(120, 8, 0, M=40)
    fs, sp = FlatSpec(layers, 0, 0.1), o.MlpSpec(layers, 0, 0.1)
    keys = o.split(o.PRNGKey(3), C)
    ref = emu_ops.sgld(fs, X, y, o.init_positions(sp, keys), keys, 16, 1e-5, 4)
    theta = np.concatenate([r[0]["theta"], r[1]["theta"]])
    assert np.array_equal(theta, ref["theta"].numpy())  # chains identical for any world size
    one = emu_ops.predict(fs, ref["theta"], Xt, want_preds=False, want_moments=True)
    for k in range(2):  # every rank holds the same pooled result
        assert int(r[k]["count"]) == C
        assert np.allclose(r[k]["mean"], one["mean"].numpy(), rtol=1e-6, atol=1e-7)
        assert np.allclose(r[k]["var"], one["var"].numpy(), rtol=1e-4, atol=1e-7)
    assert np.array_equal(r[0]["mean"], r[1]["mean"])
