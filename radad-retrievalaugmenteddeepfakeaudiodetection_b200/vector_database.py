"""Drop-in replacement for the reference's ``vector_database.py`` (module name kept on purpose).

``VectorDatabase`` keeps the reference's constructor, methods, defaults, return types, attributes,
error behaviour and quirks (reference ``vector_database.py:8-273``), but every arithmetic step runs in
hand-written sm_100a CUDA behind the C ABI of ``include/radad_flat.h``:

* ``add_vectors*``  -> fused normalise / |y|^2 / dtype-convert ingest kernel   (was numpy + faiss ``add``)
* ``search*``       -> tcgen05 or exact-fp32 score+select kernel + on-device merge (was faiss ``search``)
* ``save`` / ``load`` -> faiss ``IndexFlat`` file layout + the same ``metadata.pkl`` keys

There is NO CPU fallback: constructing a ``VectorDatabase`` without a CUDA device or without the built
``libradad_flat.so`` raises (the reference silently falls back to CPU faiss, ``:51-53,89-91``).

Config keys read (all via ``getattr`` so an unmodified reference ``Config`` works -- SURVEY section 5):
``vector_db_path``, ``vector_db_index_type`` (L2 | IP | IVF), ``use_float16``, ``normalize_for_ip``,
``vector_add_batch_size``, ``top_k``, ``vector_db_nprobe``.  New optional keys: ``db_dtype``
("f32" | "bf16" | "f16"; default f32, or f16 when ``use_float16``), ``db_keep_f32_master`` (bool),
``db_device`` (int), ``db_devices`` (list of ints or "all": row-shard the database over several GPUs of the box from
this one process -- ``MultiGpuFlatIndex``), ``restore_cosine_on_load`` (bool, opt-in fix of the ``load()`` quirk).
"""
from __future__ import annotations

import logging
import os
import pickle
from typing import Dict, List, Tuple

import numpy as np

if __package__:
    from . import _cabi
    from .flat_index import FlatIndex, _is_cuda_tensor
    from .multi_gpu import MultiGpuFlatIndex
    from ._cabi import METRIC_IP, METRIC_L2
else:
    # Flat layout, exactly how the reference imports this module (`from vector_database import VectorDatabase`,
    # pipeline.py:11): this directory is on sys.path and there is no parent package.
    import _cabi
    from flat_index import FlatIndex, _is_cuda_tensor
    from multi_gpu import MultiGpuFlatIndex
    from _cabi import METRIC_IP, METRIC_L2


class VectorDatabase:
    """GPU vector database for storing and retrieving feature vectors (reference ``:8-9``)."""

    def __init__(self, config):
        self.config = config
        self.index = None
        self.gpu_index = None
        self.vector_paths = []
        self.vector_labels = []
        self.vector_metadata = {}
        self.db_path = os.path.join(config.vector_db_path, "faiss_index.bin")
        self.metadata_path = os.path.join(config.vector_db_path, "metadata.pkl")

        self.gpu_resources = None
        self.device_id = 0
        self._labels_synced = -1

        os.makedirs(config.vector_db_path, exist_ok=True)
        self._initialize_gpu_resources()

    # ---- reference :31-53.  No CPU fallback here: failure raises. --------------------------------
    def _initialize_gpu_resources(self):
        lib = _cabi.load()                      # raises NativeLibraryMissing when the .so is absent
        dev = getattr(self.config, "db_device", None)
        if dev is None:
            try:
                import torch
                if torch.cuda.is_available():
                    dev = torch.cuda.current_device()
            except Exception:  # noqa: BLE001
                dev = None
        self.device_id = int(dev) if dev is not None else 0
        self._devices = None
        devs = getattr(self.config, "db_devices", None)
        if devs is not None:
            if isinstance(devs, str):
                import torch
                devs = list(range(torch.cuda.device_count())) if devs == "all" else [int(v) for v in devs.split(",")]
            self._devices = [int(v) for v in devs]
            self.device_id = self._devices[0]
        # probe the device through the library itself (torch is plumbing, not a requirement)
        probe = FlatIndex(1, METRIC_L2, "f32", device=self.device_id)
        probe.close()
        self.gpu_resources = lib
        logging.info(f"radad_flat GPU resources ready on device {self.device_id}")

    def _store_dtype(self) -> str:
        dt = getattr(self.config, "db_dtype", None)
        if dt is not None:
            return str(dt)
        return "f16" if bool(getattr(self.config, "use_float16", False)) else "f32"   # reference :80

    # ---- reference :56-97 ----------------------------------------------------------------------------
    def create_index(self, dimension: int):
        index_type = self.config.vector_db_index_type.upper()
        if index_type == "L2":
            metric = METRIC_L2
        elif index_type == "IP":
            metric = METRIC_IP
        elif index_type == "IVF":
            # The reference builds IndexIVFFlat(L2) here (:65-70).  The exact flat index returns what IVF
            # returns with nprobe == nlist, so it is served exactly; `nprobe` is accepted and ignored.
            logging.info("IVF requested: served by the exact flat L2 index (superset of IVF accuracy)")
            metric = METRIC_L2
        else:
            raise ValueError(f"Unsupported index type: {index_type}")
        keep = bool(getattr(self.config, "db_keep_f32_master", False))
        if self._devices is not None and len(self._devices) > 1:
            self.index = MultiGpuFlatIndex(dimension, metric, self._store_dtype(), self._devices, keep_f32_master=keep)
        else:
            self.index = FlatIndex(dimension, metric, self._store_dtype(), device=self.device_id, keep_f32_master=keep)
        logging.info(f"Created radad_flat index on GPU ({type(self.index).__name__}/{self.index.store}) dim={dimension}")
        self._cosine = (index_type == "IP") and bool(getattr(self.config, "normalize_for_ip", True))

    # ---- reference :100-105 (kept for API parity; the hot path fuses this into the ingest kernel) -----
    def _maybe_normalize(self, arr: np.ndarray) -> np.ndarray:
        if getattr(self, "_cosine", False):
            norms = np.linalg.norm(arr, axis=1, keepdims=True) + 1e-12
            arr = arr / norms
        return arr

    # ---- reference :108-151 ----------------------------------------------------------------------------
    def add_vectors_batch(self, vectors, paths: List[str], labels: List[int], metadata: Dict,
                          batch_size: int = 10000):
        if vectors.shape[0] == 0:
            logging.warning("No vectors to add to database")
            return

        if self.index is None:
            self.create_index(vectors.shape[1])

        cosine = bool(getattr(self, "_cosine", False))
        if not _is_cuda_tensor(vectors):
            vectors = np.ascontiguousarray(vectors.astype(np.float32, copy=False))

        total_vectors = vectors.shape[0]
        added = 0
        for start in range(0, total_vectors, batch_size):
            end = min(start + batch_size, total_vectors)
            batch = vectors[start:end]
            try:
                self.index.add(batch, normalize=cosine)          # fused normalise + convert + |y|^2
                added += batch.shape[0]
                self.vector_paths.extend(paths[start:end])
                self.vector_labels.extend(labels[start:end])
                for key, values in metadata.items():
                    self.vector_metadata.setdefault(key, [])
                    vals = values[start:end] if hasattr(values, '__getitem__') else [values] * len(batch)
                    self.vector_metadata[key].extend(vals)
            except Exception as e:  # noqa: BLE001 - reference :147-149 logs and skips the slice
                logging.error(f"Error adding batch {start}-{end}: {e}")
                continue

        logging.info(f"Added {added}/{total_vectors} vectors. Index ntotal={self.index.ntotal}")

    def add_vectors(self, vectors, paths: List[str], labels: List[int], metadata: Dict):
        """Add vectors to the database with automatic batching (reference :154-157)."""
        batch_size = getattr(self.config, 'vector_add_batch_size', 10000)
        self.add_vectors_batch(vectors, paths, labels, metadata, batch_size)

    # ---- reference :159-182 ----------------------------------------------------------------------------
    def search_batch(self, query_vectors, k: int = None) -> Tuple[np.ndarray, np.ndarray]:
        if self.index is None:
            raise ValueError("Vector database is empty. Build the database first.")

        k = int(k if k is not None else getattr(self.config, 'top_k', 5))
        cuda_in = _is_cuda_tensor(query_vectors)
        if query_vectors.ndim == 1:
            query_vectors = query_vectors.reshape(1, -1)
        if not cuda_in:
            query_vectors = np.ascontiguousarray(query_vectors.astype(np.float32, copy=False))

        k = min(k, self.index.ntotal)
        if k <= 0:
            logging.warning("No vectors available for search")
            return (np.zeros((len(query_vectors), 0), dtype=np.float32),
                    np.zeros((len(query_vectors), 0), dtype=np.int64))

        try:
            if hasattr(self.index, 'nprobe') and hasattr(self.config, 'vector_db_nprobe'):
                self.index.nprobe = int(self.config.vector_db_nprobe)
        except Exception:  # noqa: BLE001
            pass

        distances, indices = self.index.search(query_vectors, k, normalize=bool(getattr(self, "_cosine", False)))
        return distances, indices

    def search(self, query_vector, k: int = None) -> Tuple[np.ndarray, np.ndarray]:
        """Search for similar vectors (single query) (reference :185-188)."""
        distances, indices = self.search_batch(query_vector.reshape(1, -1), k)
        return (distances[0] if len(distances) > 0 else np.array([]),
                indices[0] if len(indices) > 0 else np.array([]))

    # ---- additive: neighbour labels + kNN label vote (north star; reference has no counterpart) -------
    def _sync_labels(self):
        n = len(self.vector_labels)
        if self.index is None or n != self.index.ntotal:
            return False
        if self._labels_synced != n:
            lab = np.empty((n,), dtype=np.float32)
            for i, v in enumerate(self.vector_labels):
                try:
                    lab[i] = float(v)            # ints, numpy scalars, 0-d torch tensors (pipeline.py:436-441)
                except Exception:  # noqa: BLE001
                    lab[i] = 0.0
            self.index.set_labels(lab)
            self._labels_synced = n
        return True

    def search_batch_with_labels(self, query_vectors, k: int = None):
        """(distances, indices, labels[nq,k]) -- labels gathered on the device in the merge kernel."""
        if self.index is None:
            raise ValueError("Vector database is empty. Build the database first.")
        k = int(k if k is not None else getattr(self.config, 'top_k', 5))
        if query_vectors.ndim == 1:
            query_vectors = query_vectors.reshape(1, -1)
        k = min(k, self.index.ntotal)
        if k <= 0:
            z = np.zeros((len(query_vectors), 0), dtype=np.float32)
            return z, np.zeros((len(query_vectors), 0), dtype=np.int64), z.copy()
        synced = self._sync_labels()
        if not _is_cuda_tensor(query_vectors):
            query_vectors = np.ascontiguousarray(query_vectors.astype(np.float32, copy=False))
        if synced:
            return self.index.search(query_vectors, k, normalize=bool(getattr(self, "_cosine", False)),
                                     return_labels=True)
        # len(vector_labels) != index.ntotal (a metadata.pkl that does not belong to the index file, or a partially failed
        # add): the device copy cannot be trusted, so index vector_labels by the returned ids on the host exactly like the
        # reference caller (pipeline.py:504) -- an id without a label raises IndexError there and here, never label 0.
        logging.error(f"labels out of sync with the index ({len(self.vector_labels)} labels, {self.index.ntotal} rows); "
                      "gathering labels on the host")
        dist, idx = self.index.search(query_vectors, k, normalize=bool(getattr(self, "_cosine", False)))
        host_idx = idx.cpu().numpy() if _is_cuda_tensor(idx) else idx
        lab = np.array([[float(self.vector_labels[int(i)]) if i >= 0 else 0.0 for i in row] for row in host_idx],
                       dtype=np.float32).reshape(host_idx.shape)
        if _is_cuda_tensor(idx):
            import torch
            return dist, idx, torch.from_numpy(lab).to(idx.device)
        return dist, idx, lab

    def label_vote(self, neighbour_labels, kvote: int = None):
        """Sum of the first ``kvote`` neighbour labels per query (spoof = 1, bona-fide = 0: dataset.py:36-44)."""
        kvote = int(kvote if kvote is not None else getattr(self.config, 'top_k', 5))
        return self.index.label_vote(neighbour_labels, kvote)

    # ---- reference :190-216 (never raises) ----------------------------------------------------------------
    def save(self):
        try:
            if self.index is None:
                logging.warning("No index to save.")
                return
            self.index.save(self.db_path)                      # faiss IndexFlat on-disk layout
            meta = {
                'paths': self.vector_paths,
                'labels': self.vector_labels,
                'metadata': self.vector_metadata,
                'index_type': self.config.vector_db_index_type,
                'dimension': self.index.d if hasattr(self.index, 'd') else None
            }
            with open(self.metadata_path, 'wb') as f:
                pickle.dump(meta, f)
            logging.info(f"Saved FAISS-format index to {self.db_path} with {self.index.ntotal} vectors")
        except Exception as e:  # noqa: BLE001
            logging.error(f"Error saving vector database: {e}")

    # ---- reference :218-242 (never raises; does NOT restore _cosine) ----------------------------------------
    def load(self):
        try:
            if not (os.path.exists(self.db_path) and os.path.exists(self.metadata_path)):
                logging.warning("No saved vector database found")
                return
            with open(self.metadata_path, 'rb') as f:
                meta = pickle.load(f)
            self.vector_paths = meta['paths']
            self.vector_labels = meta['labels']
            self.vector_metadata = meta['metadata']

            keep = bool(getattr(self.config, "db_keep_f32_master", False))
            if self._devices is not None and len(self._devices) > 1:
                self.index = MultiGpuFlatIndex.load(self.db_path, self._store_dtype(), self._devices, keep_f32_master=keep)
            else:
                self.index = FlatIndex.load(self.db_path, self._store_dtype(), device=self.device_id, keep_f32_master=keep)
            self._labels_synced = -1
            if bool(getattr(self.config, "restore_cosine_on_load", False)):
                self._cosine = (str(meta.get('index_type', '')).upper() == "IP") and \
                    bool(getattr(self.config, "normalize_for_ip", True))
            logging.info(f"Loaded FAISS-format index to GPU (FlatIndex); ntotal={self.index.ntotal}")
        except Exception as e:  # noqa: BLE001
            logging.error(f"Error loading vector database: {e}")

    # ---- reference :245-256 ------------------------------------------------------------------------------
    def get_gpu_memory_usage(self):
        try:
            idx = self.index if self.index is not None else FlatIndex(1, METRIC_L2, "f32", device=self.device_id)
            info = idx.mem_info()
            used = info["total"] - info["free"]
            return {'used': int(used), 'total': int(info["total"]), 'utilization': float(used / info["total"])}
        except Exception:  # noqa: BLE001
            pass
        return None

    # ---- reference :259-273 ------------------------------------------------------------------------------
    def cleanup_gpu_resources(self, release_index: bool = True):
        """Clean up GPU resources to prevent memory leaks (reference :259-268).

        In the reference this drops the GPU index handle and leaves the rest to destructors, and its only caller is
        ``__del__`` (:270-273).  Here it actually returns the device memory: the search scratch (the counterpart of
        ``faiss.StandardGpuResources``' temporary memory) and -- unless ``release_index=False`` -- the stored rows
        (``rdb_destroy``); ``self.index`` then becomes ``None`` exactly as before the first add, so a later search raises
        the reference's "Vector database is empty" error and ``load()`` brings a saved database back.  Idempotent."""
        if self.gpu_resources is not None:
            try:
                del self.gpu_index
                self.gpu_index = None
                idx = self.index
                if idx is not None:
                    if release_index:
                        self.index = None
                        self._labels_synced = -1
                        idx.close()
                    else:
                        idx.release_scratch()
                logging.info("GPU resources cleaned up")
            except Exception as e:  # noqa: BLE001
                logging.warning(f"Error during GPU cleanup: {e}")

    def __del__(self):
        """Destructor to ensure GPU cleanup."""
        try:
            self.cleanup_gpu_resources()
        except Exception:  # noqa: BLE001
            pass
