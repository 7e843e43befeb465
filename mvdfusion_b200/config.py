"""yaml -> object factory and checkpoint loader with the semantics of the reference's utils/load_model.py:18-110.

`target:` strings written for the reference (`mvdfusion.viewfusion_zero_depth_rgb.ViewFusion`, `mvdfusion.unet.UNetModel`,
…) resolve to this package's mirror modules, so the reference's configs/*.yaml load unchanged.
`external.sd1.ldm.models.autoencoder.AutoencoderKL` resolves to this package's VAE (encode / decode).  Other targets outside the hot
path (CLIP, ...) resolve to None unless the reference package itself is importable: those components are out of scope
 and the facade reports them as absent.
"""
import importlib
from collections import OrderedDict

import torch
import yaml

_ALIASES = {"mvdfusion.": "mvdfusion_b200.mvdfusion.",
            # the VAE: configs/mvd_gso.yaml:53-54
            "external.sd1.ldm.models.autoencoder.": "mvdfusion_b200.mvdfusion.autoencoder.",
            # the CLIP image embedder: viewfusion_zero_depth_rgb.py:103-105
            "external.sd1.ldm.modules.encoders.modules.": "mvdfusion_b200.mvdfusion.clip_encoder."}
_OUT_OF_SCOPE_PREFIXES = ("external.sd1.",)


def load_yaml(path):
    with open(path, "r") as f:
        return yaml.safe_load(f)


def get_obj_from_str(string):
    module, cls = string.rsplit(".", 1)
    for src, dst in _ALIASES.items():
        if module.startswith(src) or module + "." == src:
            module = (dst + module[len(src):]).rstrip(".")
            break
    return getattr(importlib.import_module(module), cls)


def instantiate_from_config(config):
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    target = config["target"]
    if target.startswith(_OUT_OF_SCOPE_PREFIXES) and not any(target.startswith(a) for a in _ALIASES):
        try:
            module, cls = target.rsplit(".", 1)
            return getattr(importlib.import_module(module), cls)(**(config.get("params") or {}))
        except ImportError:
            return None
    return get_obj_from_str(target)(**(config.get("params") or {}))


def load_model_from_config(config, ckpt=None, verbose=False, replace_key=None, ignore_keys=(), param_mapper=None, remove_keys=()):
    """utils/load_model.py:28-110: instantiate, optionally load `ckpt['state_dict']` after key replace / remap / removal
    (strict=False; missing core keys are reported, not fatal), return the model in eval mode."""
    model = instantiate_from_config(config)
    if model is None:
        return None
    param_mapper = param_mapper or {}
    if ckpt:
        print(f"Loading {config['target']} from {ckpt}")
        pl_sd = torch.load(ckpt, map_location="cpu")
        sd = pl_sd["state_dict"]
        renamed = OrderedDict()
        for k, v in sd.items():
            name = k.replace(replace_key[0], replace_key[1]) if replace_key is not None else k
            name = param_mapper.get(name, name)
            if name in remove_keys:
                print("REMOVING WEIGHT", name)
                continue
            renamed[name] = v
        missing, unexpected = model.load_state_dict(renamed, strict=False)
        core = [m for m in missing if not any(ik in m for ik in ignore_keys)]
        for m in core:
            print("missing core:", m)
        if core:
            print(f"***\n***CRITICAL WARNING\nMissing core parameters while loading {config['target']} from {ckpt}")
        if verbose and unexpected:
            print("unexpected keys:", len(unexpected))
    model.eval()
    return model
