"""torchrun worker for tests/test_gpu_multi.py: row-sharded search over NCCL == unsharded search."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("radad-retrievalaugmenteddeepfakeaudiodetection_b200")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N, Dm, Q, k = 30011, 256, 333, 10
    rng = np.random.default_rng(1234)
    xb = rng.standard_normal((N, Dm)).astype(np.float32)
    xb[20000] = xb[3]                                       # tie across shards
    xq = np.random.default_rng(5678).standard_normal((Q, Dm)).astype(np.float32)
    xq[0] = xb[3]
    labels = (np.arange(N) % 2).astype(np.float32)
    for metric, cos, store, exch in ((pkg.METRIC_L2, False, "bf16", "peer"), (pkg.METRIC_IP, True, "bf16", "peer"),
                                     (pkg.METRIC_L2, False, "f32", "peer"), (pkg.METRIC_L2, False, "bf16", "nccl"),
                                     (pkg.METRIC_IP, True, "bf16", "nccl")):
        sh = pkg.ShardedFlatIndex(Dm, metric, store, device=local, exchange=exch)
        s, e = sh.set_shard(N)
        sh.add_local(xb[s:e], normalize=cos)
        sh.set_labels_local(labels[s:e])
        q = torch.from_numpy(xq).to(dev)
        for rep in range(3):                                # repeated calls alternate the two peer half-buffers
            D, I, L = sh.search(q, k, normalize=cos)
        # host ingress: every rank uploads 1/G of the (pinned) batch, NVLink all-gather assembles it (Q = 333: ragged)
        qh = torch.from_numpy(xq).pin_memory()
        D2, I2, L2 = sh.search_from_host(qh, k, normalize=cos)
        torch.cuda.synchronize()
        assert torch.equal(I2, I) and torch.equal(D2, D) and torch.equal(L2, L), f"rank {rank}: host-ingress path differs"
        full = pkg.FlatIndex(Dm, metric, store, device=local)
        full.add(xb, normalize=cos)
        full.set_labels(labels)
        Df, If, Lf = full.search(xq, k, normalize=cos, return_labels=True)
        assert (I.cpu().numpy() == If).all(), f"rank {rank}: ids differ ({store}, metric {metric})"
        assert np.allclose(D.cpu().numpy(), Df, rtol=1e-6, atol=1e-6), f"rank {rank}: distances differ"
        assert (L.cpu().numpy() == Lf).all()
        if metric == pkg.METRIC_L2 and store == "bf16":
            assert If[0, 0] == 3 and If[0, 1] == 20000
    # odd nq * k beyond the minimum buffer size: the int64 id plane of the peer buffer must stay 8-byte aligned
    sh = pkg.ShardedFlatIndex(Dm, pkg.METRIC_L2, "bf16", device=local, exchange="peer")
    s, e = sh.set_shard(N)
    sh.add_local(xb[s:e])
    q2 = torch.from_numpy(np.random.default_rng(99).standard_normal((1001, Dm)).astype(np.float32)).to(dev)
    for rep in range(2):
        D, I, L = sh.search(q2, 5)
    torch.cuda.synchronize()
    full = pkg.FlatIndex(Dm, pkg.METRIC_L2, "bf16", device=local)
    full.add(xb)
    Df, If = full.search(q2, 5)
    assert torch.equal(I, If) and torch.equal(D, Df), f"rank {rank}: odd nq*k case differs"
    dist.barrier()
    if rank == 0:
        print(f"MULTI_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
