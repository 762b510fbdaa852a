"""Worker of tests/test_dropin_import.py: run in a fresh interpreter with the PACKAGE DIRECTORY on sys.path, importing the
product exactly as the reference does (``from vector_database import VectorDatabase``, pipeline.py:11) and driving it
with the reference caller's loop (pipeline.py:449-532: search_batch + one index.reconstruct per neighbour) on the
fixtures that the reference's own code produced (tests/golden/make_golden.py).

    python dropin_worker.py <repo root> <fixture name> <tmp dir>
"""
import os
import sys

root, name, tmp = sys.argv[1], sys.argv[2], sys.argv[3]
pkg_dir = os.path.join(root, "radad-retrievalaugmenteddeepfakeaudiodetection_b200")
sys.path.insert(0, pkg_dir)                      # INTEGRATION.md route A
from vector_database import VectorDatabase      # noqa: E402  -- the reference's import line, verbatim

assert VectorDatabase.__module__ == "vector_database"
assert os.path.dirname(os.path.abspath(sys.modules["vector_database"].__file__)) == pkg_dir
sys.path.append(root)                            # checker only: the oracle's restatement of the caller loop
import numpy as np                               # noqa: E402
import torch                                     # noqa: E402
from oracle.flat_oracle import retrieve_similar_vectors_oracle   # noqa: E402

g = np.load(os.path.join(root, "tests", "golden", f"{name}.npz"), allow_pickle=True)


class Config:                                    # the keys the reference Config carries for this path (config.py)
    vector_db_path = tmp
    vector_db_index_type = str(g["index_type"])
    top_k = int(g["K"])
    use_float16 = False
    normalize_for_ip = True
    vector_add_batch_size = 128                  # several add slices


K, D = int(g["K"]), g["xb"].shape[1]
vdb = VectorDatabase(Config())                   # pipeline.py:90
paths = [str(p) for p in g["paths"]]
labels = [torch.tensor(int(v)) for v in g["labels"]]             # 0-d tensors (pipeline.py:436-441)
vdb.add_vectors(g["xb"], paths, labels, {"speaker_id": ["s"] * len(paths)})     # pipeline.py:444-445
assert vdb.index.ntotal == len(paths) and vdb.index.d == D
qpaths = [str(p) for p in g["qpaths"]]
train_ids = {str(s) for s in g["train_ids"]}
for tag, kw in (("excl_paths", dict(query_paths=qpaths, exclude_self=True)),
                ("excl_train", dict(query_paths=None, exclude_self=True, training_file_ids=train_ids)),
                ("noexcl", dict(query_paths=qpaths, exclude_self=False))):
    vec, lbl, pth, dst = retrieve_similar_vectors_oracle(vdb, g["q"], K, D, **kw)
    assert [list(r) for r in pth] == [list(map(str, r)) for r in g[f"{tag}_paths"]], tag
    np.testing.assert_array_equal(lbl, g[f"{tag}_lbl"])
    # fp32 store: index.reconstruct returns the stored row bit for bit (cosine: the device-side x / (|x| + 1e-12) equals
    # numpy's), hence identical neighbour tensors ...
    np.testing.assert_array_equal(vec, g[f"{tag}_vec"])
    np.testing.assert_allclose(dst, g[f"{tag}_dist"], rtol=1e-4, atol=2e-4, equal_nan=True)
    if tag == "excl_paths":
        # ... and identical downstream predictions: the seeded reference RADADModel (a torch.jit trace written by
        # make_golden.py next to the fixture) on OUR neighbours gives the reference's logits bit for bit
        model = torch.jit.load(os.path.join(root, "tests", "golden", f"radad_model_{name}.pt")).eval()
        with torch.no_grad():
            logits = model(torch.from_numpy(vec), torch.from_numpy(g["q"]))
        np.testing.assert_array_equal(logits.numpy(), g["logits_excl_paths"])
# save / load through the same import path (pipeline.py / app.py / main.py call load())
vdb.save()
v2 = VectorDatabase(Config())
v2.load()
assert v2.index.ntotal == len(paths) and v2.vector_paths == paths
np.testing.assert_array_equal(v2.index.reconstruct(7), vdb.index.reconstruct(7))
vdb.cleanup_gpu_resources()
print("DROPIN_OK", name)
