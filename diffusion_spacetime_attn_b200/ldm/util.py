"""`ldm.util` helpers the entrypoints use (reference ldm/util.py:78-93)."""
import importlib


def get_obj_from_str(string, reload=False):
    module, cls = string.rsplit(".", 1)
    if module.startswith("ldm."):  # the mirror lives inside the package
        module = "diffusion_spacetime_attn_b200." + module
    return getattr(importlib.import_module(module), cls)


def instantiate_from_config(config):
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**config.get("params", dict()))
