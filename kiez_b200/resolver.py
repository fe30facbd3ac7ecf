"""Name -> class resolution for ``Kiez(algorithm=..., hubness=...)``.

The reference delegates this to the third-party ``class_resolver`` package
(kiez/neighbors/__init__.py:20-26, kiez/hubness_reduction/__init__.py:9-12);
only the behaviour kiez relies on is provided here: a query may be ``None``
(default class), a string (case-insensitive, base-class-name suffix optional:
"NoHubnessReduction" -> "no"), a class, or an instance (returned as-is).
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, Optional, Type


class Resolver:
    def __init__(self, base: Type, classes: Iterable[Type], default: Optional[Type] = None):
        self.base = base
        self.default = default
        self._suffix = base.__name__.lower()
        self._lookup: Dict[str, Type] = {}
        for cls in classes:
            self.register(cls)

    def _norm(self, name: str) -> str:
        key = "".join(ch for ch in name.lower() if ch not in "_- ")
        if key != self._suffix and key.endswith(self._suffix):
            key = key[: -len(self._suffix)]
        return key

    def register(self, cls: Type, synonyms: Iterable[str] = ()) -> None:
        self._lookup[self._norm(cls.__name__)] = cls
        for s in synonyms:
            self._lookup[self._norm(s)] = cls

    @property
    def options(self):
        return set(self._lookup)

    def lookup(self, query) -> Type:
        if query is None:
            if self.default is None:
                raise ValueError("No default available")
            return self.default
        if isinstance(query, type):
            return query
        if isinstance(query, str):
            key = self._norm(query)
            if key not in self._lookup:
                raise KeyError(f"Invalid query: {query}. Try one of: {sorted(self._lookup)}")
            return self._lookup[key]
        raise TypeError(f"Invalid query type: {type(query)}")

    def make(self, query, pos_kwargs: Optional[Dict[str, Any]] = None, **kwargs):
        if query is not None and not isinstance(query, (str, type)):
            return query
        return self.lookup(query)(**(pos_kwargs or {}), **kwargs)
