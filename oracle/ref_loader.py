"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference env for golden-vector generation.

Uses `/root/reference` where it exists (the build container), else the byte-identical copy that
`baseline/install_ref.py` left in the git-ignored `baseline/_ref/` (the GPU box).  Puts
`oracle/ref_shims/` (stand-ins for gym/pygame/shapely/qpsolvers/matplotlib, all
absent from the image) plus `/root/reference` and `/root/reference/scripts` on
`sys.path`, then imports `merging_gym` exactly as the reference scripts do
(scripts/main.py:1-2,20).  Never imported by the product package; `bench.py` uses it only in its CPU-baseline legs
(to time the reference's own `MergeEnv.step` on the box's host cores).
"""
import contextlib
import io
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "ref_shims")
# The unmodified copy made by baseline/install_ref.py (git-ignored; it travels to the GPU box, /root/reference does not)
TRAVELLING_COPY = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")


def _has_env(root: str) -> bool:
    return os.path.isfile(os.path.join(root, "merging_gym", "envs", "merging_env.py")) and \
        os.path.isfile(os.path.join(root, "scripts", "helper.py"))


def _pick_root() -> str:
    env = os.environ.get("MERGING_GYM_REFERENCE")
    if env:
        return env
    return "/root/reference" if _has_env("/root/reference") else TRAVELLING_COPY


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return _has_env(REFERENCE_ROOT)


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
