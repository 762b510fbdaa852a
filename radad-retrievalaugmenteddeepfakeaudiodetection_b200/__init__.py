"""radad_b200 -- B200-native retrieval hot path of RADAD (exact flat nearest-neighbour search).

The directory name contains a hyphen (it mirrors the reference repository name), so import it with
``importlib.import_module("radad-retrievalaugmenteddeepfakeaudiodetection_b200")`` or put this
directory on ``sys.path`` and ``import vector_database`` exactly as the reference's flat layout does.

Public surface (drop-in for the reference's ``vector_database.py``):
    VectorDatabase        add_vectors / add_vectors_batch / search / search_batch / save / load / ...
    FlatIndex             the object behind ``VectorDatabase.index`` (faiss flat-index duck type)
    retrieve_similar_vectors   device-resident version of pipeline.py:449-532
    ShardedFlatIndex      row-sharded multi-GPU search (one process per GPU, NCCL all-gather + device merge)
    MultiGpuFlatIndex     the same row shards driven by ONE process (``config.db_devices``): single-process drop-in
"""
from ._cabi import (ALGO_AUTO, ALGO_SIMT, ALGO_TC, METRIC_IP, METRIC_L2, STORE_BF16, STORE_F16, STORE_F32,
                    NativeLibraryMissing, load as load_native)
from .flat_index import FlatIndex
from .vector_database import VectorDatabase
from .retrieval import retrieve_similar_vectors
from .sharded import ShardedFlatIndex, shard_bounds
from .multi_gpu import MultiGpuFlatIndex

__all__ = ["VectorDatabase", "FlatIndex", "retrieve_similar_vectors", "ShardedFlatIndex", "MultiGpuFlatIndex", "shard_bounds",
           "load_native", "NativeLibraryMissing", "METRIC_L2", "METRIC_IP", "STORE_F32", "STORE_BF16",
           "STORE_F16", "ALGO_AUTO", "ALGO_SIMT", "ALGO_TC"]
