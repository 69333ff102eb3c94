"""Runs tests/test_gpu_parity.py::test_random_models_and_batches_across_paths for many seeds (a one-off soak, not a test)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distgcn_b200 import engine as E
from tests import test_gpu_parity as T


class Env:   # what the test suite's monkeypatch fixture does: the library re-reads its options after every change
    def setenv(self, k, v):
        os.environ[k] = v
        E.reload_env()

    def delenv(self, k, raising=True):
        if k in os.environ:
            del os.environ[k]
        elif raising:
            raise KeyError(k)
        E.reload_env()


ctx = E.Context(0)
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = []
for seed in range(lo, hi):
    try:
        T.test_random_models_and_batches_across_paths.__wrapped__(ctx, seed, Env()) if hasattr(
            T.test_random_models_and_batches_across_paths, "__wrapped__") else T.test_random_models_and_batches_across_paths(ctx, seed, Env())
    except AssertionError as e:
        bad.append((seed, str(e)[:200]))
        print("seed", seed, "FAILED", str(e)[:200], flush=True)
print("seeds %d..%d: %d failures %s" % (lo, hi - 1, len(bad), bad))
