"""Stand-in for fastcore.basics.patch (test infrastructure only).

`@patch def f(self: Cls, ...)` attaches `f` to `Cls`, the class named in the annotation of
the first parameter.  Used by /root/reference/diffdrr/drr.py:155 and detector.py:97,144.
"""
import sys
import typing


def patch(fn):
    first = next(iter(fn.__code__.co_varnames[: fn.__code__.co_argcount]))
    ann = fn.__annotations__[first]
    if isinstance(ann, str):  # `from __future__ import annotations`
        ann = eval(ann, sys.modules[fn.__module__].__dict__)
    targets = typing.get_args(ann) or (ann,)
    for cls in targets:
        setattr(cls, fn.__name__, fn)
    return fn
