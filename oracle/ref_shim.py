"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference (dobraczka/kiez
v0.5.0, mounted read-only at /root/reference) in the authoring container.

The reference cannot be imported as shipped here:

1. ``class_resolver`` is not installed (kiez/kiez.py:6,
   kiez/neighbors/__init__.py:1, kiez/hubness_reduction/__init__.py:1).
2. ``importlib.metadata.version("kiez")`` raises because kiez is not
   pip-installed (kiez/__init__.py:5).
3. scikit-learn 1.9 ``check_is_fitted`` rejects plain objects
   (kiez/neighbors/neighbor_algorithm_base.py:117, csls.py:85,
   local_scaling.py:129, mutual_proximity.py:154-163, dis_sim.py:139-142).

The three shims below are applied at *import time only*; no file under
/root/reference is touched.  This module is used solely

* by ``oracle/make_golden.py`` to generate the committed fixtures under
  ``tests/golden/`` from the real reference, and
* by ``tests/test_oracle_vs_reference.py`` and the plugin tests (skipped automatically when
  no copy of the reference is present), and
* by ``bench.py --impl reference``, which drives the reference's own ``Kiez`` from the
  unmodified copy under ``baseline/_ref`` (tools/vendor_reference.sh)

so that the numpy restatement in ``oracle/kiez_oracle.py`` is *pinned* against
outputs of the reference itself.  Nothing in the product package imports it.
"""
from __future__ import annotations

import importlib
import importlib.metadata
import os
import sys
import types

_VENDORED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                         "baseline", "_ref")


def _find_reference_root() -> str:
    """/root/reference in the authoring container; on the GPU box the unmodified copy that
    tools/vendor_reference.sh placed under baseline/_ref (git-ignored, travels with gpurun)."""
    env = os.environ.get("KIEZ_REFERENCE_ROOT")
    if env:
        return env
    for root in ("/root/reference", _VENDORED):
        if os.path.isdir(os.path.join(root, "kiez")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "kiez"))


# --------------------------------------------------------------------------
# shim 1: a stand-in for the tiny part of class-resolver 0.4.x that kiez uses
# --------------------------------------------------------------------------
def _normalize(name: str, suffix: str) -> str:
    name = name.lower().replace("_", "").replace("-", "").replace(" ", "")
    suffix = suffix.lower()
    if suffix and name != suffix and name.endswith(suffix):
        name = name[: -len(suffix)]
    return name


class _ClassResolver:
    """``ClassResolver.from_subclasses(base, skip=, default=)`` + ``make`` /
    ``lookup`` / ``options`` / ``register`` as used by kiez/kiez.py:118-129 and
    kiez/neighbors/util.py:28-38."""

    def __init__(self, classes, base, default=None, suffix=None):
        self.base = base
        self.default = default
        self.suffix = base.__name__ if suffix is None else suffix
        self.lookup_dict = {}
        for cls in classes:
            self.register(cls)

    @classmethod
    def from_subclasses(cls, base, skip=None, default=None, **kwargs):
        skip = set(skip or ())
        found = []
        stack = list(base.__subclasses__())
        while stack:
            sub = stack.pop()
            stack.extend(sub.__subclasses__())
            if sub in skip or sub in found:
                continue
            found.append(sub)
        return cls(found, base=base, default=default, **kwargs)

    def normalize_cls(self, cls) -> str:
        return _normalize(cls.__name__, self.suffix)

    def register(self, cls, synonyms=None, raise_on_conflict=True):
        self.lookup_dict[self.normalize_cls(cls)] = cls
        for syn in synonyms or ():
            self.lookup_dict[_normalize(syn, self.suffix)] = cls

    @property
    def options(self):
        return set(self.lookup_dict)

    def lookup(self, query, default=None):
        if query is None:
            default = default or self.default
            if default is None:
                raise ValueError("no default given")
            return default
        if isinstance(query, type):
            return query
        if isinstance(query, str):
            key = _normalize(query, self.suffix)
            if key not in self.lookup_dict:
                raise KeyError(f"{query} is an invalid. Try one of: {sorted(self.options)}")
            return self.lookup_dict[key]
        raise TypeError(f"Invalid query type: {type(query)}")

    def make(self, query, pos_kwargs=None, **kwargs):
        if query is not None and not isinstance(query, (str, type)):
            return query  # an instance is returned as-is
        cls = self.lookup(query)
        return cls(**(pos_kwargs or {}), **kwargs)


def _install_class_resolver_shim():
    try:
        import class_resolver  # noqa: F401

        return
    except ImportError:
        pass
    mod = types.ModuleType("class_resolver")
    mod.ClassResolver = _ClassResolver

    class _HintOrType:
        def __class_getitem__(cls, item):
            return object

    mod.HintOrType = _HintOrType
    mod.__kiez_b200_shim__ = True
    sys.modules["class_resolver"] = mod


# --------------------------------------------------------------------------
# shim 3: attribute-presence check_is_fitted (what sklearn 1.3.2 did)
# --------------------------------------------------------------------------
def _check_is_fitted(estimator, attributes=None, *, msg=None, all_or_any=all):
    from sklearn.exceptions import NotFittedError

    if attributes is None:
        fitted = any(v.endswith("_") and not v.startswith("__") for v in vars(estimator))
    else:
        if not isinstance(attributes, (list, tuple)):
            attributes = [attributes]
        fitted = all_or_any([hasattr(estimator, a) for a in attributes])
    if not fitted:
        raise NotFittedError(
            msg or f"This {type(estimator).__name__} instance is not fitted yet."
        )


_REF = None


def load_reference():
    """Import the reference package under the three shims; returns the module."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise ImportError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_class_resolver_shim()
    # shim 2: version("kiez")
    orig_version = importlib.metadata.version

    def _version(name):
        if name == "kiez":
            return "0.5.0"
        return orig_version(name)

    importlib.metadata.version = _version
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        kiez = importlib.import_module("kiez")
    finally:
        importlib.metadata.version = orig_version
        sys.path.remove(REFERENCE_ROOT)
    for modname in (
        "kiez.neighbors.neighbor_algorithm_base",
        "kiez.hubness_reduction.csls",
        "kiez.hubness_reduction.local_scaling",
        "kiez.hubness_reduction.mutual_proximity",
        "kiez.hubness_reduction.dis_sim",
    ):
        importlib.import_module(modname).check_is_fitted = _check_is_fitted
    _REF = kiez
    return kiez


def reference_kneighbors(source, target, *, metric="euclidean", hubness=None,
                         hubness_kwargs=None, n_candidates=10, k=None,
                         n_jobs=None):
    """Run the reference's own ``Kiez(SklearnNN brute).fit().kneighbors()``."""
    kiez = load_reference()
    from kiez.neighbors import SklearnNN

    algo = SklearnNN(n_candidates=n_candidates, metric=metric, algorithm="brute",
                     n_jobs=n_jobs)
    inst = kiez.Kiez(n_candidates=n_candidates, algorithm=algo, hubness=hubness,
                     hubness_kwargs=dict(hubness_kwargs or {}))
    inst.fit(source, target)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return inst.kneighbors(k)
