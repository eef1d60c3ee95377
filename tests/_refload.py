"""Load the vendored reference modules by file path (build container only).

The reference package ``__init__`` pulls in matplotlib and ``plum`` which are
absent, so ``unet.py`` / ``nn.py`` / ``sde_diffusion.py`` are loaded under a stub
package.  Returns None when /root/reference is absent (the GPU box).
"""
import importlib.util
import os
import sys
import types

REF_DIR = "/root/reference/amortised diffusion/image_diffusion"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "unet.py"))


def load_reference():
    if not reference_available():
        return None
    name = "_ref_image_diffusion"
    if name in sys.modules:
        return sys.modules[name]
    pkg = types.ModuleType(name)
    pkg.__path__ = [REF_DIR]
    sys.modules[name] = pkg
    if "functorch" not in sys.modules:
        try:
            import functorch  # noqa: F401
        except Exception:
            import torch.func as tf
            stub = types.ModuleType("functorch"); stub.vmap = tf.vmap; stub.grad = tf.grad
            sys.modules["functorch"] = stub
    for sub in ("nn", "unet", "sde_diffusion"):
        spec = importlib.util.spec_from_file_location(f"{name}.{sub}", os.path.join(REF_DIR, f"{sub}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"{name}.{sub}"] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, sub, mod)
    return pkg


def _install_plum_stub():
    """Minimal stand-in for ``plum.dispatch`` / ``plum.Union`` (absent here): overloads are tried in definition order and
    the first whose annotated parameters all match by ``isinstance`` wins (``Callable`` annotations match any callable)."""
    if "plum" in sys.modules:
        return
    import collections.abc
    import inspect
    import typing

    registry = {}

    def matches(value, ann):
        if ann is inspect.Parameter.empty:
            return True
        origin = typing.get_origin(ann)
        if origin is typing.Union:
            return any(matches(value, a) for a in typing.get_args(ann))
        if origin is collections.abc.Callable or ann is typing.Callable:
            return callable(value)
        return isinstance(value, ann)

    def dispatch(fn):
        key = (fn.__module__, fn.__qualname__)
        registry.setdefault(key, []).append((inspect.signature(fn), fn))

        def call(*args, **kwargs):
            # the overloads in the files loaded here are disjoint, so the order does not matter
            for sig, impl in reversed(registry[key]):
                try:
                    bound = sig.bind(*args, **kwargs)
                except TypeError:
                    continue
                if all(matches(v, sig.parameters[k].annotation) for k, v in bound.arguments.items()):
                    return impl(*args, **kwargs)
            raise TypeError(f"no overload of {fn.__qualname__} matches")
        call.__name__ = fn.__name__
        return call

    plum = types.ModuleType("plum")
    plum.dispatch = dispatch
    plum.Union = typing.Union
    sys.modules["plum"] = plum


def load_reference_samplers():
    """``conditioning`` / ``likelihoods`` / ``sampling`` of the reference on top of :func:`load_reference` (``plum`` and
    ``matplotlib`` are stubbed: the first only routes overloads, the second is only used by plotting helpers)."""
    pkg = load_reference()
    if pkg is None:
        return None
    if hasattr(pkg, "sampling"):
        return pkg
    _install_plum_stub()
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    name = pkg.__name__
    for sub in ("plotting_utils", "conditioning", "likelihoods", "sampling"):
        spec = importlib.util.spec_from_file_location(f"{name}.{sub}", os.path.join(REF_DIR, f"{sub}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"{name}.{sub}"] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, sub, mod)
    return pkg
