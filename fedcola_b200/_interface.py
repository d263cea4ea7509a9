"""Declarative abstract interfaces.

The reference resolves its server / client / optimizer classes by name and only relies on a handful of method and
attribute names.  Instead of hand-writing one stub per name, the base classes of this package are
generated from a table: `abstract_interface(name, attributes, required)` returns an ABC whose instances start with
the listed attributes set to their defaults and whose listed methods are abstract (instantiating a subclass that
misses one raises TypeError, calling one through super() raises NotImplementedError)."""
import sys
from abc import ABCMeta, abstractmethod


def _stub(method_name):
    def method(self, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__}.{method_name}() is part of the interface and must be overridden")
    method.__name__ = method_name
    return abstractmethod(method)


def abstract_interface(name, doc, attributes, required):
    """attributes: {attribute name: default}; required: iterable of method names every concrete class must define."""
    defaults = dict(attributes)

    def __init__(self, **kwargs):
        for key, value in defaults.items():
            setattr(self, key, value)

    namespace = {"__doc__": doc, "__init__": __init__, "interface_attributes": tuple(defaults),
                 "interface_methods": tuple(required), "__module__": sys._getframe(1).f_globals.get("__name__", __name__)}
    namespace.update({m: _stub(m) for m in required})
    return ABCMeta(name, (object,), namespace)
