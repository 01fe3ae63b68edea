"""LoRA save / load round trip for the fused-layout denoisers (second half).

The reference saves adapters with `get_peft_model_state_dict(unwrapped_model)` (helpers/training/save_hooks.py:862-891:
keys `transformer_blocks.0.attn.to_q.lora_A.weight`, adapter name stripped) and hands them to the pipeline's
`save_lora_weights(output_dir, transformer_lora_layers=...)`, which prefixes the component name and writes
`pytorch_lora_weights.safetensors`.  The libstb200 denoisers keep PERMANENTLY fused q|k|v weights internally
(blocks.AttnPlan.w_qkv) but their adapter parameters stay un-fused and PEFT-named, so the saved file is the same one the
reference (and ComfyUI-style loaders) expect.  For checkpoints written by a reference run with `--fuse_qkv_projections`
(helpers/training/diffusers_overrides.py:133-466 targets `to_qkv` / `add_qkv_proj`), `unfuse_qkv_lora` / `fuse_qkv_lora`
convert between the two key layouts (block-diagonal B, stacked A).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch

WEIGHT_NAME = "pytorch_lora_weights.safetensors"
_GROUPS: Tuple[Tuple[str, Tuple[str, str, str]], ...] = (("to_qkv", ("to_q", "to_k", "to_v")),
                                                        ("add_qkv_proj", ("add_q_proj", "add_k_proj", "add_v_proj")))


def get_peft_model_state_dict(model: torch.nn.Module, adapter_name: str = "default") -> Dict[str, torch.Tensor]:
    """peft.utils.get_peft_model_state_dict for LoRA: every `lora_A / lora_B` tensor of `adapter_name`, name stripped."""
    out = {}
    tag = f".{adapter_name}."
    for k, v in model.state_dict().items():
        if ".lora_A." in k or ".lora_B." in k:
            if tag in k:
                out[k.replace(tag, ".")] = v.detach()
    return out


def set_peft_model_state_dict(model: torch.nn.Module, state: Dict[str, torch.Tensor], adapter_name: str = "default") -> None:
    own = dict(model.named_parameters())
    missing = []
    with torch.no_grad():
        for k, v in state.items():
            full = k.replace(".lora_A.", f".lora_A.{adapter_name}.").replace(".lora_B.", f".lora_B.{adapter_name}.")
            if full not in own:
                missing.append(k)
                continue
            if own[full].shape != v.shape:
                raise ValueError(f"{k}: shape {tuple(v.shape)} does not match the adapter's {tuple(own[full].shape)}")
            own[full].copy_(v.to(device=own[full].device, dtype=own[full].dtype))
    if missing:
        raise KeyError(f"LoRA keys without a matching adapter tensor: {missing[:4]}{' ...' if len(missing) > 4 else ''}")


def save_lora_weights(output_dir, transformer_lora_layers: Dict[str, torch.Tensor], component: str = "transformer",
                      metadata: Optional[Dict[str, str]] = None) -> str:
    """diffusers `*LoraLoaderMixin.save_lora_weights`: `<component>.<key>` -> pytorch_lora_weights.safetensors."""
    import os

    from safetensors.torch import save_file

    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(output_dir, WEIGHT_NAME)
    save_file({f"{component}.{k}": v.detach().to("cpu").contiguous() for k, v in transformer_lora_layers.items()}, path,
              metadata={"format": "pt", **(metadata or {})})
    return path


def load_lora_weights(path_or_dir, component: str = "transformer") -> Dict[str, torch.Tensor]:
    import os

    from safetensors.torch import load_file

    path = os.path.join(path_or_dir, WEIGHT_NAME) if os.path.isdir(path_or_dir) else path_or_dir
    pre = component + "."
    return {k[len(pre):]: v for k, v in load_file(path).items() if k.startswith(pre)}


def fuse_qkv_lora(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Un-fused (to_q / to_k / to_v) adapter keys -> the fused `to_qkv` / `add_qkv_proj` layout: A stacked [3r, K],
    B block-diagonal [3N, 3r] (exactly equivalent: x A_f^T B_f^T = cat(x A_q^T B_q^T, ...))."""
    out = dict(state)
    for fused, members in _GROUPS:
        for key in [k for k in state if k.endswith(f".{members[0]}.lora_A.weight")]:
            base = key[: -len(f"{members[0]}.lora_A.weight")]
            a = [state.get(f"{base}{m}.lora_A.weight") for m in members]
            b = [state.get(f"{base}{m}.lora_B.weight") for m in members]
            if any(t is None for t in a + b):
                continue
            r, n = a[0].shape[0], b[0].shape[0]
            bf = torch.zeros((3 * n, 3 * r), dtype=b[0].dtype)
            for i in range(3):
                bf[i * n:(i + 1) * n, i * r:(i + 1) * r] = b[i]
            out[f"{base}{fused}.lora_A.weight"] = torch.cat(a, 0)
            out[f"{base}{fused}.lora_B.weight"] = bf
            for m in members:
                out.pop(f"{base}{m}.lora_A.weight"), out.pop(f"{base}{m}.lora_B.weight")
    return out


def unfuse_qkv_lora(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Inverse of `fuse_qkv_lora`.  A fused adapter trained as ONE rank-R LoRA on the fused projection has a dense B and no
    exact un-fused equivalent of the same rank; it is split as three rank-R adapters sharing A (B rows of each member)."""
    out = dict(state)
    for fused, members in _GROUPS:
        for key in [k for k in state if k.endswith(f".{fused}.lora_A.weight")]:
            base = key[: -len(f"{fused}.lora_A.weight")]
            a, b = state[key], state[f"{base}{fused}.lora_B.weight"]
            n, R = b.shape[0] // 3, a.shape[0]
            r = R // 3
            block_diag = R % 3 == 0
            if block_diag:
                off = b.clone()
                for i in range(3):
                    off[i * n:(i + 1) * n, i * r:(i + 1) * r] = 0
                block_diag = not bool(off.any())
            for i, m in enumerate(members):
                if block_diag:
                    out[f"{base}{m}.lora_A.weight"] = a[i * r:(i + 1) * r].clone()
                    out[f"{base}{m}.lora_B.weight"] = b[i * n:(i + 1) * n, i * r:(i + 1) * r].clone()
                else:
                    out[f"{base}{m}.lora_A.weight"] = a.clone()
                    out[f"{base}{m}.lora_B.weight"] = b[i * n:(i + 1) * n].clone()
            out.pop(key), out.pop(f"{base}{fused}.lora_B.weight")
    return out
