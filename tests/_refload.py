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
