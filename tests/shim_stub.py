"""A stand-in for the reference's `ModelFoundation` family class + `Trainer` glue, reproducing exactly the call ORDER the
shims rely on (so the shim code is exercised without importing the reference, which needs diffusers / accelerate / peft):

  * `load_model` (helpers/models/common.py:3400-3548): `self.model = MODEL_CLASS.from_pretrained(...)`, `.to(device)`, then
    `configure_chunked_feed_forward()`, `apply_gradient_checkpointing_settings()`, `fuse_qkv_projections()`,
    `post_model_load_setup()` in this order (:3543-3548);
  * `add_lora_adapter` (:1049-1117): builds a LoraConfig-shaped object and calls `self.model.add_adapter(self.lora_config)`;
  * `Trainer.model_predict` (helpers/training/trainer.py:6051-6107) incl. the x-prediction fix-up (:6099-6105), and
    `_compute_model_prediction_loss` (:6109-6117).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch


class StubFoundation:
    """Reference family base: methods the shim falls back to are recorded in `self.calls`."""

    MODEL_CLASS = None            # set by the test: a callable returning the "diffusers" module (same state-dict names)
    LORA_TARGETS = None

    def __init__(self, config, device):
        self.config = config
        self.accelerator = SimpleNamespace(device=torch.device(device))
        self.model = None
        self.calls = []

    # -- common.py:3400-3548
    def load_model(self, move_to_device=True):
        self.model = self.MODEL_CLASS()
        if move_to_device:
            self.model.to(self.accelerator.device)
        self.calls.append("configure_chunked_feed_forward")
        self.apply_gradient_checkpointing_settings()
        self.fuse_qkv_projections()
        self.post_model_load_setup()

    def apply_gradient_checkpointing_settings(self):
        self.calls.append("apply_gradient_checkpointing_settings")
        if getattr(self.config, "gradient_checkpointing", False):
            self.model.gradient_checkpointing = True
            interval = getattr(self.config, "gradient_checkpointing_interval", None)
            if interval:
                self.model.gradient_checkpointing_interval = interval

    def fuse_qkv_projections(self):
        self.calls.append("fuse_qkv_projections")

    def post_model_load_setup(self):
        self.calls.append("post_model_load_setup")

    # -- common.py:1049-1117
    def add_lora_adapter(self):
        c = self.config
        rank = c.lora_rank
        alpha = c.lora_alpha if c.lora_alpha is not None else rank
        self.lora_config = SimpleNamespace(r=rank, lora_alpha=alpha, lora_dropout=c.lora_dropout,
                                           init_lora_weights=True, target_modules=self.LORA_TARGETS)
        hook = getattr(self.model, "register_lora_custom_modules", None)
        if callable(hook):
            hook(self.lora_config)
        self.model.add_adapter(self.lora_config)

    # -- reference step methods (the fallback path)
    def prepare_batch(self, batch, state):
        self.calls.append("ref.prepare_batch")
        return batch

    def model_predict(self, prepared_batch, **kw):
        self.calls.append("ref.model_predict")
        return {"model_prediction": None}

    def loss(self, prepared_batch, model_output, apply_conditioning_mask=True):
        self.calls.append("ref.loss")
        return torch.zeros(())

    def loss_with_logs(self, prepared_batch, model_output, apply_conditioning_mask=True):
        self.calls.append("ref.loss_with_logs")
        return torch.zeros(()), None


class StubTrainer:
    """trainer.py:6051-6132 reduced to the calls that cross seam B9."""

    def __init__(self, model, noise_scheduler=None):
        self.model = model
        self.noise_scheduler = noise_scheduler
        self.config = SimpleNamespace(disable_accelerator=False, controlnet=False)

    def model_predict(self, prepared_batch):
        model_pred = self.model.model_predict(prepared_batch=prepared_batch)
        if (hasattr(self.noise_scheduler, "config") and hasattr(self.noise_scheduler.config, "prediction_type")
                and self.noise_scheduler.config.prediction_type == "sample"):
            # trainer.py:6099-6105 applies `model_pred - noise`; with the dict contract of :6085-6087 that is the entry
            model_pred = dict(model_pred)
            model_pred["model_prediction"] = model_pred["model_prediction"] - prepared_batch["noise"]
        return model_pred

    def compute_model_prediction_loss(self, prepared_batch):
        model_pred = self.model_predict(prepared_batch)
        loss, logs = self.model.loss_with_logs(prepared_batch=prepared_batch, model_output=model_pred,
                                               apply_conditioning_mask=True)
        return loss, logs, model_pred
