"""Checkpoint ingestion for the drop-in: the reference's ``init_from_ckpt`` rules (fit/utils/eval_utils.py:12-71)
in front of ``fitv2_b200.FiT``.

``model_ema.safetensors`` (or a ``torch.save`` file) -> ``load_state_dict(strict=False)`` with the same key handling
as the reference: the ``_orig_mod.`` prefix of ``torch.compile`` checkpoints is added / removed to match the model,
``ignore_keys`` are regular expressions matched with ``re.match``.  The parameter names and shapes of
``fitv2_b200.FiT`` are the reference's, so released FiTv2 checkpoints load unchanged; the kernel-side
layouts (stacked bf16 block weights, interleaved fc1_g / fc1_x rows, fp32 conditioning weights) are rebuilt lazily
from the loaded parameters at the next forward (``FiT.pack_weights``).
"""
from __future__ import annotations

import re
from typing import Iterable, Optional, Tuple

import torch


def load_checkpoint_file(checkpoint_dir: str) -> dict:
    """eval_utils.py:15-22: safetensors first for ``*.safetensors`` (falling back to torch.load), torch.load otherwise."""
    if checkpoint_dir.endswith(".safetensors"):
        try:
            from safetensors.torch import load_file
            return dict(load_file(checkpoint_dir))
        except Exception:
            return dict(torch.load(checkpoint_dir, map_location="cpu"))
    return dict(torch.load(checkpoint_dir, map_location="cpu"))


def match_checkpoint_keys(model_keys: Iterable[str], ckpt: dict, verbose: bool = False) -> dict:
    """eval_utils.py:26-52: reconcile the ``_orig_mod.`` prefix between the model and the checkpoint."""
    model_keys = set(model_keys)
    if set(ckpt.keys()) == model_keys:
        return ckpt
    model_has = any(k.startswith("_orig_mod.") for k in model_keys)
    ckpt_has = any(k.startswith("_orig_mod.") for k in ckpt.keys())
    if model_has and not ckpt_has:
        if verbose:
            print("Added '_orig_mod.' prefix to checkpoint keys to match compiled model.")
        return {f"_orig_mod.{k}": v for k, v in ckpt.items()}
    if not model_has and ckpt_has:
        if verbose:
            print("Removed '_orig_mod.' prefix from checkpoint keys to match non-compiled model.")
        return {(k[len("_orig_mod."):] if k.startswith("_orig_mod.") else k): v for k, v in ckpt.items()}
    return ckpt


def init_from_ckpt(model, checkpoint_dir: str, ignore_keys: Optional[Iterable[str]] = None, verbose: bool = False
                   ) -> Tuple[list, list]:
    """Same call as the reference (eval_utils.py:12-71); additionally returns (missing, unexpected)."""
    ckpt = match_checkpoint_keys(model.state_dict().keys(), load_checkpoint_file(checkpoint_dir), verbose)
    for k in list(ckpt.keys()):
        if ignore_keys:
            for ik in ignore_keys:
                if re.match(ik, k):
                    print("Deleting key {} from state_dict.".format(k))
                    ckpt.pop(k, None)
    missing, unexpected = model.load_state_dict(ckpt, strict=False)
    if verbose:
        print(f"Restored with {len(missing)} missing and {len(unexpected)} unexpected keys")
        if len(missing) > 0:
            print(f"Missing Keys: {missing}")
        if len(unexpected) > 0:
            print(f"Unexpected Keys: {unexpected}")
        print("")
    return list(missing), list(unexpected)
