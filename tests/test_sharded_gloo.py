"""CPU, world_size = 2, gloo: the multi-GPU exchange step (all-gather of packed per-shard candidates) and the
row partitioning, checked against the oracle's merge of the same candidates."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG_NAME, ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import importlib
    pkg = importlib.import_module(PKG_NAME)
    sh = importlib.import_module(PKG_NAME + ".sharded")
    orc = importlib.import_module("oracle.flat_oracle")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, D, Q, k = 1001, 24, 9, 6
        xb = np.random.default_rng(1234).standard_normal((N, D)).astype(np.float32)
        xq = np.random.default_rng(5678).standard_normal((Q, D)).astype(np.float32)
        s, e = pkg.shard_bounds(N, world, rank)
        local = orc.FlatIndexOracle(D, orc.METRIC_IP)
        local.add(xb[s:e])
        Dl, Il = local.search(xq, k)                     # per-shard candidates (IP: key == distance)
        key = torch.from_numpy(Dl)
        gid = torch.from_numpy(Il + s)
        lab = torch.from_numpy(((Il + s) % 2).astype(np.float32))
        gkey, ggid, glab = sh.exchange_candidates(key, gid, lab)
        assert gkey.shape == (Q, world, k) and ggid.dtype == torch.int64
        assert torch.equal(gkey[:, rank], key) and torch.equal(ggid[:, rank], gid) and torch.equal(glab[:, rank], lab)
        # merge the gathered lists the way the device kernel does (key desc, id asc) and compare to unsharded
        full = orc.FlatIndexOracle(D, orc.METRIC_IP)
        full.add(xb)
        Df, If = full.search(xq, k)
        mk = gkey.reshape(Q, -1).numpy()
        mi = ggid.reshape(Q, -1).numpy()
        for r in range(Q):
            order = np.lexsort((mi[r], -mk[r]))[:k]
            assert (mi[r][order] == If[r]).all()
            assert np.allclose(mk[r][order], Df[r], rtol=1e-6)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_exchange_two_ranks_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))
