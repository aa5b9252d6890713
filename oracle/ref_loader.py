"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference env for golden-vector generation.

Only usable in the build container, where `/root/reference` exists.  Puts
`oracle/ref_shims/` (stand-ins for gym/pygame/shapely/qpsolvers/matplotlib, all
absent from the image) plus `/root/reference` and `/root/reference/scripts` on
`sys.path`, then imports `merging_gym` exactly as the reference scripts do
(scripts/main.py:1-2,20).  Never imported by the product package, by `-m gpu`
tests, by `smoke()` or by `bench.py`.
"""
import contextlib
import io
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get("MERGING_GYM_REFERENCE", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "merging_gym", "envs", "merging_env.py"))


def load_reference_env():
    """Return a fresh instance of the reference's `MergeEnv` via `gym.make` (shimmed)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (os.path.join(REFERENCE_ROOT, "scripts"), REFERENCE_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)   # `is "ego"`, `is 1` under py3.12
        import gym  # the shim
        import merging_gym  # noqa: F401  registers merging_env-v0 (merging_gym/__init__.py:3-6)
    with contextlib.redirect_stdout(io.StringIO()):
        env = gym.make("merging_env-v0")
    return env.unwrapped


@contextlib.contextmanager
def quiet():
    """The reference prints on every collision (merging_env.py:204)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
