"""Named registry used for the noise schedules (interface of generative/utils/component_store.py:27-117)."""
from __future__ import annotations

from keyword import iskeyword
from typing import Any, Callable, Iterable, NamedTuple


class _Entry(NamedTuple):
    description: str
    value: Any


class ComponentStore:
    """``store.add_def(name, desc)`` decorates a function; ``store[name]`` / ``store.name`` fetches it."""

    def __init__(self, name: str, description: str) -> None:
        self.components: dict[str, _Entry] = {}
        self.name, self.description = name, description
        self.__doc__ = f"Component Store '{name}': {description}\n{self.__doc__ or ''}".strip()

    def add(self, name: str, desc: str, value: Any) -> Any:
        if not (name.isidentifier() and not iskeyword(name)):
            raise ValueError("Name of component must be valid Python identifier")
        self.components[name] = _Entry(desc, value)
        return value

    def add_def(self, name: str, desc: str) -> Callable:
        return lambda func: self.add(name, desc, func)

    def __contains__(self, name: str) -> bool:
        return name in self.components

    def __len__(self) -> int:
        return len(self.components)

    def __iter__(self) -> Iterable:
        yield from self.components

    def __str__(self) -> str:
        rows = "\n".join(f"* {k}:\n    {v.description}" for k, v in self.components.items())
        return f"Component Store '{self.name}': {self.description}\nAvailable components:\n{rows}"

    def __getattr__(self, name: str) -> Any:
        comps = self.__dict__.get("components", {})
        if name in comps:
            return comps[name].value
        raise AttributeError(name)

    def __getitem__(self, name: str) -> Any:
        if name in self.components:
            return self.components[name].value
        raise ValueError(f"Component '{name}' not found")
