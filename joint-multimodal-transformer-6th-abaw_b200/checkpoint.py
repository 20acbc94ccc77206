"""Checkpoint I/O with the reference's file names and key conventions (main.py:105-177 writes,
main.py:54-70 / 515-585 reads), plus what the reference lacks: optimizer / scaler / epoch state for a true resume
(the reference's `start_epoch` only offsets the loop counter).

The drop-in modules keep the reference's state_dict keys, so `fusion_w.pt` etc. written here load into the
reference modules with strict=True and vice versa (a leading `module.` from nn.DataParallel is stripped on load,
as load_clean_weights does).
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, Optional

import torch

# role -> file name (main.py:116-176)
FILES = {
    "fusion_model": "fusion_w.pt",
    "backbone_pretrainer": "backbone_pretrainer_w.pt",
    "fc_layer_for_audio_concat": "fc_layer_for_audio_concat.pt",
    "transformer_audio_modality_fusion": "transformer_audio_modality_fusion.pt",
    "fc_layer_for_video_concat": "fc_layer_for_video_concat.pt",
    "transformer_visio_modality_fusion": "transformer_visio_modality_fusion.pt",
    "temporal": "temporal_tcn.pt",            # the TCN head of I3D_WSDDA (inside vision_i3d.pt in the reference)
}
RESUME_FILE = "resume_state.pt"


def _cpu_state(module: torch.nn.Module) -> "OrderedDict[str, torch.Tensor]":
    sd = module.module.state_dict() if hasattr(module, "module") else module.state_dict()      # tools.get_state_dict
    return OrderedDict((k, v.detach().to("cpu")) for k, v in sd.items())                      # tools.state_dict_to_cpu


def dump_models_into_disk(path: str, modules: Dict[str, Optional[torch.nn.Module]], epoch: int = 0, optimizer=None,
                          scaler=None, extra: Optional[dict] = None):
    """Write one `<role file>.pt` CPU state_dict per non-None module (roles = keys of FILES) and, when an optimizer is
    given, RESUME_FILE with optimizer / GradScaler / epoch state."""
    os.makedirs(path, exist_ok=True)
    for role, mod in modules.items():
        if mod is None:
            continue
        if role not in FILES:
            raise KeyError(f"unknown checkpoint role {role!r}; known: {sorted(FILES)}")
        torch.save(_cpu_state(mod), os.path.join(path, FILES[role]))
    if optimizer is not None:
        torch.save({"epoch": int(epoch), "optimizer": optimizer.state_dict(),
                    "scaler": scaler.state_dict() if scaler is not None else None, "extra": extra or {}},
                   os.path.join(path, RESUME_FILE))


def load_clean_weights(w_path: str, map_location="cpu") -> "OrderedDict[str, torch.Tensor]":
    """main.py:54-70: load a state_dict, strip nn.DataParallel's `module.` prefix, move to `map_location`."""
    sd = torch.load(w_path, map_location="cpu", weights_only=True)
    if any(k.startswith("module.") for k in sd):
        sd = OrderedDict((k.replace("module.", ""), v) for k, v in sd.items())
    return OrderedDict((k, v.to(map_location)) for k, v in sd.items())


def load_models_from_disk(path: str, modules: Dict[str, Optional[torch.nn.Module]], optimizer=None, scaler=None,
                          map_location="cpu") -> dict:
    """strict=True load of every given module (main.py:515-585) and, if present and requested, the resume state.
    Returns {'epoch': ..., 'extra': ...} (epoch 0 when there is no resume file)."""
    for role, mod in modules.items():
        if mod is None:
            continue
        target = mod.module if hasattr(mod, "module") else mod
        target.load_state_dict(load_clean_weights(os.path.join(path, FILES[role]), map_location), strict=True)
    info = {"epoch": 0, "extra": {}}
    rp = os.path.join(path, RESUME_FILE)
    if optimizer is not None and os.path.exists(rp):
        st = torch.load(rp, map_location="cpu", weights_only=False)
        optimizer.load_state_dict(st["optimizer"])
        if scaler is not None and st.get("scaler") is not None:
            scaler.load_state_dict(st["scaler"])
        info = {"epoch": st["epoch"], "extra": st.get("extra", {})}
    return info


def extract_submodule_state(state_dict: "Dict[str, torch.Tensor]", prefix: str) -> "OrderedDict[str, torch.Tensor]":
    """Keys under `prefix` with the prefix removed -- e.g. the TCN head out of a reference `vision_i3d.pt`
    (main.py:157-159 saves the whole I3D_WSDDA, whose TemporalConvNet lives under `temporal.`, I3DWSDDA.py:26-28):
        tcn.load_state_dict(extract_submodule_state(load_clean_weights(".../vision_i3d.pt"), "temporal."), strict=True)"""
    out = OrderedDict((k[len(prefix):], v) for k, v in state_dict.items() if k.startswith(prefix))
    if not out:
        raise KeyError(f"no key starts with {prefix!r}")
    return out
