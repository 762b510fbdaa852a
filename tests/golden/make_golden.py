#!/usr/bin/env python
"""Generate golden fixtures by running the REFERENCE'S OWN SOURCE on seeded inputs.

Run in the build container only (needs /root/reference, which the GPU box lacks):

    python tests/golden/make_golden.py                       # everything
    python tests/golden/make_golden.py --only NAME [NAME ...]   # only these search fixtures (e.g. lattice_l2_k200)

What runs unmodified from /root/reference: ``vector_database.py::VectorDatabase`` (all
wrapper logic: normalisation, batching, metadata bookkeeping, k clamping, return types),
``pipeline.py::DeepfakeDetectionPipeline.retrieve_similar_vectors`` (rank-ordered
self-exclusion, reconstruct, padding) and ``radad_model.py::RADADModel`` (consumer).

What is stubbed: third-party modules that are absent here and cannot be installed
(``faiss``, ``librosa``, ``matplotlib``).  The ``faiss`` stub below implements
IndexFlatL2 / IndexFlatIP by their *mathematical definition* (exhaustive search, float64
arithmetic, result cast to float32, ties -> lowest id) -- deliberately independent of
``oracle/flat_oracle.py`` so the oracle can be checked against these fixtures.
FAISS's own fp32 arithmetic therefore stays unpinned (see oracle header).

Fixtures are small ``.npz`` files with inputs and outputs; tests regenerate nothing.
"""
import os
import sys
import types
import tempfile

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------ stub faiss ----------
def _make_faiss_stub():
    m = types.ModuleType("faiss")
    m.METRIC_L2 = 1
    m.METRIC_INNER_PRODUCT = 0

    class _Flat:
        metric = None

        def __init__(self, d):
            self.d = int(d)
            self.is_trained = True
            self._x = np.zeros((0, self.d), np.float32)

        @property
        def ntotal(self):
            return self._x.shape[0]

        def add(self, x):
            x = np.ascontiguousarray(x, dtype=np.float32)
            assert x.ndim == 2 and x.shape[1] == self.d
            self._x = np.concatenate([self._x, x], 0)

        def reconstruct(self, i):
            return self._x[int(i)].copy()

        def search(self, q, k):
            q64 = np.asarray(q, np.float64)
            x64 = self._x.astype(np.float64)
            nq = q64.shape[0]
            D = np.empty((nq, k), np.float32)
            I = np.empty((nq, k), np.int64)
            for r in range(nq):
                if self.metric == "L2":
                    diff = x64 - q64[r]
                    s = np.einsum("ij,ij->i", diff, diff)
                    order = np.lexsort((np.arange(len(s)), s))[:k]
                else:
                    s = x64 @ q64[r]
                    order = np.lexsort((np.arange(len(s)), -s))[:k]
                D[r] = s[order].astype(np.float32)
                I[r] = order
            return D, I

    class IndexFlatL2(_Flat):
        metric = "L2"

    class IndexFlatIP(_Flat):
        metric = "IP"

    def write_index(index, path):
        np.savez(path + ".stub.npz", x=index._x, metric=index.metric)
        open(path, "wb").write(b"stub")

    def read_index(path):
        z = np.load(path + ".stub.npz")
        idx = (IndexFlatL2 if str(z["metric"]) == "L2" else IndexFlatIP)(z["x"].shape[1])
        idx.add(z["x"])
        return idx

    def index_gpu_to_cpu(index):
        raise RuntimeError("stub: not a GPU index")

    m.IndexFlatL2, m.IndexFlatIP = IndexFlatL2, IndexFlatIP
    m.write_index, m.read_index, m.index_gpu_to_cpu = write_index, read_index, index_gpu_to_cpu
    return m


def _install_stubs():
    import importlib.machinery
    import transformers  # noqa: F401  (resolve its lazy availability probes BEFORE stubbing)
    from transformers import Wav2Vec2Model, Wav2Vec2Processor  # noqa: F401
    sys.modules["faiss"] = _make_faiss_stub()
    for name in ("librosa", "matplotlib", "matplotlib.pyplot"):
        mod = types.ModuleType(name)
        mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules.setdefault(name, mod)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)


# ------------------------------------------------------------------ input generators ----
def gaussian(n, d, seed):
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)


def lattice(n, d, seed):
    """Integer entries in {-2..2}: every dot product / norm is exact in bf16 and fp32."""
    return np.random.default_rng(seed).integers(-2, 3, size=(n, d)).astype(np.float32)


def make_cfg(Config, tmp, index_type, top_k=5, **extra):
    import torch
    cfg = Config()
    cfg.vector_db_path = tmp
    cfg.vector_db_index_type = index_type
    cfg.top_k = top_k
    cfg.device = torch.device("cpu")
    for k, v in extra.items():
        setattr(cfg, k, v)
    return cfg


def main():
    _install_stubs()
    import torch
    from config import Config                     # reference, unmodified
    from vector_database import VectorDatabase    # reference, unmodified
    import pipeline as ref_pipeline               # reference, unmodified
    from radad_model import RADADModel            # reference, unmodified

    assert VectorDatabase.__module__ == "vector_database"
    assert os.path.abspath(sys.modules["vector_database"].__file__).startswith(REF)

    # ---------------- search fixtures -------------------------------------------------
    search_cases = [
        # name,            gen,      N,    D,   Q,  k,  index_type
        ("kat_tiny_l2",    lattice,  8,    4,   3,  5,  "L2"),
        ("kat_tiny_ip",    lattice,  8,    4,   3,  5,  "IP"),
        ("lattice_l2",     lattice,  700,  64,  24, 15, "L2"),
        ("lattice_ip",     lattice,  700,  64,  24, 15, "IP"),
        ("gauss_l2",       gaussian, 3000, 96,  32, 15, "L2"),
        ("gauss_cos",      gaussian, 3000, 96,  32, 15, "IP"),
        ("gauss_l2_small", gaussian, 1500, 96,  5,  15, "L2"),   # nq < 20: FAISS direct path
        ("kclamp_l2",      gaussian, 7,    32,  4,  15, "L2"),   # k > ntotal -> clamped
        ("ref_shape_l2",   gaussian, 640,  448, 16, 15, "L2"),   # D = 7*64 (TPP 1-2-4 layout)
        ("lattice_l2_k200", lattice, 700,  64,  24, 200, "L2"),  # k beyond the fused selectors (dense keys + radix select)
        ("lattice_ip_k300", lattice, 700,  64,  24, 300, "IP"),
    ]
    models_only = "--models-only" in sys.argv      # only (re)write the traced consumer models next to the fixtures
    only = set(sys.argv[sys.argv.index("--only") + 1:]) if "--only" in sys.argv else (set() if models_only else None)
    for name, gen, N, D, Q, k, itype in search_cases:
        if only is not None and name not in only:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            cfg = make_cfg(Config, tmp, itype,
                           normalize_for_ip=(not name.startswith("lattice_ip") and name != "kat_tiny_ip"))
            xb = gen(N, D, seed=1234)
            xq = gen(Q, D, seed=5678)
            if gen is lattice:
                xb[N // 2] = xb[1]            # exact duplicate rows -> distance ties
                xq[0] = xb[1]                 # exact self match -> distance 0
            paths = [f"/data/spk{i % 13}/utt_{i:05d}.wav" for i in range(N)]
            labels = [int(v) for v in np.random.default_rng(91011).integers(0, 2, N)]
            meta = {"speaker_id": [f"spk{i % 13}" for i in range(N)]}
            vdb = VectorDatabase(cfg)
            cfg.vector_add_batch_size = 256    # several add slices
            vdb.add_vectors(xb, paths, labels, meta)
            dist, idx = vdb.search_batch(xq, k=k)
            d1, i1 = vdb.search(xq[0], k=k)
            ddef, idef = vdb.search_batch(xq[:2])          # k defaults to config.top_k
            rec = np.stack([vdb.index.reconstruct(int(i)) for i in idx[0]])
            np.savez_compressed(
                os.path.join(OUT, f"search_{name}.npz"),
                xb=xb, xq=xq, k=np.int64(k), index_type=itype,
                normalize_for_ip=bool(getattr(cfg, "normalize_for_ip", True)),
                labels=np.asarray(labels, np.int64),
                dist=dist, idx=idx, dist_single=d1, idx_single=i1,
                dist_default=ddef, idx_default=idef, recon_row0=rec,
                ntotal=np.int64(vdb.index.ntotal), cosine=bool(vdb._cosine),
                n_paths=np.int64(len(vdb.vector_paths)),
                n_meta=np.int64(len(vdb.vector_metadata["speaker_id"])))
            print(f"search_{name}: dist{dist.shape} idx{idx.shape} cosine={vdb._cosine}")

    if only is not None and not models_only:
        return

    if not models_only:
        # ---------------- wrapper-behaviour fixtures --------------------------------------
        with tempfile.TemporaryDirectory() as tmp:
            cfg = make_cfg(Config, tmp, "L2")
            vdb = VectorDatabase(cfg)
            beh = {}
            try:
                vdb.search_batch(np.zeros((1, 8), np.float32))
            except ValueError as e:
                beh["empty_search_error"] = str(e)
            vdb.add_vectors(np.zeros((0, 8), np.float32), [], [], {})
            beh["index_none_after_empty_add"] = vdb.index is None
            xb = gaussian(10, 8, 1)
            # scalar (non-indexable) metadata value is replicated per row (:145)
            vdb.add_vectors_batch(xb, [f"p{i}" for i in range(10)], list(range(10)),
                                  {"split": 7, "speaker_id": [f"s{i}" for i in range(10)]}, batch_size=4)
            beh["meta_split"] = vdb.vector_metadata["split"]
            beh["meta_speaker"] = vdb.vector_metadata["speaker_id"]
            beh["labels"] = vdb.vector_labels
            d0, i0 = vdb.search_batch(xb[:2], k=0)     # k=0 -> falsy? no: `k if k is not None` -> 0 -> empty
            beh["k0_shapes"] = [list(d0.shape), list(i0.shape)]
            beh["k0_dtypes"] = [str(d0.dtype), str(i0.dtype)]
            vdb.save()
            vdb2 = VectorDatabase(cfg)
            vdb2.load()
            beh["loaded_ntotal"] = int(vdb2.index.ntotal)
            beh["loaded_has_cosine_attr"] = hasattr(vdb2, "_cosine")
            beh["loaded_labels"] = vdb2.vector_labels
            with open(vdb.metadata_path, "rb") as f:
                import pickle
                beh["pickle_keys"] = sorted(pickle.load(f).keys())
            cfg_bad = make_cfg(Config, tmp, "HNSW")
            try:
                VectorDatabase(cfg_bad).create_index(8)
            except ValueError as e:
                beh["bad_type_error"] = str(e)
            import json
            with open(os.path.join(OUT, "wrapper_behaviour.json"), "w") as f:
                json.dump(beh, f, indent=1, sort_keys=True)
            print("wrapper_behaviour:", beh)

        # cosine-after-load quirk: IP index, queries NOT normalised after load()
        with tempfile.TemporaryDirectory() as tmp:
            cfg = make_cfg(Config, tmp, "IP")
            xb = gaussian(200, 32, 7)
            xq = gaussian(6, 32, 8) * 3.0
            vdb = VectorDatabase(cfg)
            vdb.add_vectors(xb, [f"p{i}" for i in range(200)], [0] * 200, {})
            d_before, i_before = vdb.search_batch(xq, k=5)
            vdb.save()
            vdb2 = VectorDatabase(cfg)
            vdb2.load()
            d_after, i_after = vdb2.search_batch(xq, k=5)
            np.savez_compressed(os.path.join(OUT, "quirk_cosine_after_load.npz"), xb=xb, xq=xq,
                                d_before=d_before, i_before=i_before, d_after=d_after, i_after=i_after)
            print("quirk: max |d_after/d_before| =", float(np.abs(d_after / d_before).max()))

    # ---------------- caller fixtures: retrieve_similar_vectors + RADADModel -----------
    for name, itype in (("retrieve_l2", "L2"), ("retrieve_cos", "IP")):
        with tempfile.TemporaryDirectory() as tmp:
            N, D, B, K = 400, 56, 12, 5            # D = 7 * feature_dim(8)
            cfg = make_cfg(Config, tmp, itype, top_k=K)
            cfg.feature_dim = 8
            cfg.use_batch_norm = False
            cfg.use_layer_norm = True               # main.py:65-66
            xb = gaussian(N, D, 21)
            paths = [f"/train/spk{i % 9}/clip_{i % 150:04d}.wav" for i in range(N)]  # repeated basenames
            labels = [torch.tensor(int(v)) for v in np.random.default_rng(5).integers(0, 2, N)]  # 0-d tensors (pipeline.py:436-441)
            vdb = VectorDatabase(cfg)
            vdb.add_vectors(xb, paths, labels, {"speaker_id": [f"spk{i % 9}" for i in range(N)]})
            # queries: half are (perturbed) DB rows carrying the DB row's own path (self-match)
            rng = np.random.default_rng(33)
            q = gaussian(B, D, 22)
            qpaths = [f"/val/x/none_{i}.wav" for i in range(B)]
            for j in range(0, B, 2):
                src = int(rng.integers(0, N))
                q[j] = xb[src] + 0.01 * rng.standard_normal(D).astype(np.float32)
                qpaths[j] = paths[src]
            fake_self = types.SimpleNamespace(
                config=cfg, device=torch.device("cpu"), vector_db=vdb,
                tpp=types.SimpleNamespace(get_output_dim=lambda: D),
                training_file_ids={os.path.basename(p) for p in paths[:40]})
            fn = ref_pipeline.DeepfakeDetectionPipeline.retrieve_similar_vectors
            out = {}
            for tag, kw in (("excl_paths", dict(query_paths=qpaths, exclude_self=True)),
                            ("excl_train", dict(query_paths=None, exclude_self=True)),
                            ("noexcl", dict(query_paths=qpaths, exclude_self=False))):
                vec, lbl, pth, dst = fn(fake_self, torch.from_numpy(q), return_info=True,
                                        return_distances=True, **kw)
                out[f"{tag}_vec"] = vec.numpy()
                out[f"{tag}_lbl"] = lbl.numpy()
                out[f"{tag}_dist"] = dst.numpy()
                out[f"{tag}_paths"] = np.array(pth, dtype=object).astype(str)
            # consumer: reference RADADModel on the retrieved neighbours (eval mode, seeded init)
            torch.manual_seed(0)
            model = RADADModel(cfg, D).eval()
            with torch.no_grad():
                logits = model(torch.from_numpy(out["excl_paths_vec"]), torch.from_numpy(q))
            out["logits_excl_paths"] = logits.numpy()
            # The consumer itself as a fixture: a torch.jit trace of the seeded reference model (a graph of aten ops with
            # its weights, not source), so the GPU box -- where /root/reference does not exist -- can run the unmodified
            # RADADModel arithmetic on the neighbours OUR retrieval returns and compare the logits with the ones above.
            traced = torch.jit.trace(model, (torch.from_numpy(out["excl_paths_vec"]), torch.from_numpy(q)))
            with torch.no_grad():
                assert torch.equal(traced(torch.from_numpy(out["excl_paths_vec"]), torch.from_numpy(q)), logits)
            torch.jit.save(traced, os.path.join(OUT, f"radad_model_{name}.pt"))
            if "--models-only" in sys.argv:
                old = np.load(os.path.join(OUT, f"{name}.npz"), allow_pickle=True)
                assert np.array_equal(old["logits_excl_paths"], out["logits_excl_paths"]), "fixture drifted"
                assert np.array_equal(old["excl_paths_vec"], out["excl_paths_vec"]), "fixture drifted"
                print(name, "traced model written; logits equal the committed fixture")
                continue
            np.savez_compressed(os.path.join(OUT, f"{name}.npz"), xb=xb, q=q,
                                paths=np.array(paths), qpaths=np.array(qpaths),
                                labels=np.array([int(l) for l in labels], np.int64),
                                train_ids=np.array(sorted(fake_self.training_file_ids)),
                                K=np.int64(K), index_type=itype, **out)
            print(name, "vec", out["excl_paths_vec"].shape, "logits", logits.numpy().ravel()[:3])


if __name__ == "__main__":
    main()
