"""The drop-in boundary, tested the way the reference uses it: ``pipeline.py:11`` does
``from vector_database import VectorDatabase`` with the module's directory on ``sys.path`` (flat layout, no parent
package).  INTEGRATION.md route A tells a maintainer to put the package directory first on ``sys.path``; these tests do
exactly that in a fresh interpreter."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "radad-retrievalaugmenteddeepfakeaudiodetection_b200")


def _run(code_or_args, cwd):
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    return subprocess.run([sys.executable] + code_or_args, cwd=cwd, env=env, capture_output=True, text=True, timeout=600)


def test_flat_layout_import_resolves(tmp_path):
    """No GPU needed: every module of the product imports with the package directory on sys.path and no parent
    package (round 1 raised `ImportError: attempted relative import with no known parent package` here)."""
    code = (f"import sys; sys.path.insert(0, {PKG_DIR!r})\n"
            "from vector_database import VectorDatabase\n"
            "import flat_index, multi_gpu, sharded, retrieval, _cabi\n"
            "assert VectorDatabase.__module__ == 'vector_database'\n"
            "assert flat_index.FlatIndex is sys.modules['vector_database'].FlatIndex\n"
            "lib = _cabi.load(); assert lib.rdb_abi_version() == _cabi.ABI_VERSION\n"
            "print('FLAT_OK')\n")
    r = _run(["-c", code], str(tmp_path))
    assert r.returncode == 0 and "FLAT_OK" in r.stdout, r.stderr[-2000:]


def test_package_import_still_works(tmp_path):
    code = (f"import sys, importlib; sys.path.insert(0, {ROOT!r})\n"
            "pkg = importlib.import_module('radad-retrievalaugmenteddeepfakeaudiodetection_b200')\n"
            "assert pkg.VectorDatabase.__module__.endswith('.vector_database')\n"
            "print('PKG_OK')\n")
    r = _run(["-c", code], str(tmp_path))
    assert r.returncode == 0 and "PKG_OK" in r.stdout, r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["retrieve_l2", "retrieve_cos"])
def test_dropin_through_the_reference_import_path(name, tmp_path):
    """`from vector_database import VectorDatabase` + the reference caller loop (search_batch + index.reconstruct per
    neighbour, pipeline.py:449-532) == the outputs the reference's own code produced: neighbour tensors bit-equal,
    labels / paths equal, and the seeded reference RADADModel's logits on our neighbours bit-equal."""
    r = _run([os.path.join(ROOT, "tests", "dropin_worker.py"), ROOT, name, str(tmp_path / "db")], str(tmp_path))
    assert r.returncode == 0 and f"DROPIN_OK {name}" in r.stdout, (r.stdout[-1000:], r.stderr[-3000:])
