"""GPU (-m gpu): the SDXL step plumbing (BASELINE configs[0] geometry: rank-irrelevant, 2 x 256^2 images -> latents
[2, 4, 32, 32]) around a stand-in UNet with the reference's call signature (sdxl/model.py:306-373).  The UNet itself stays the
reference's module in a real run; what is checked is everything libstb200 owns around it: integer timestep selection,
the fused fp32 add_noise kernel (bit-exact vs the eager DDPMScheduler chain), the time-ids / text-embeds conditioning dict,
the weighted epsilon / v-prediction loss kernel on the NCHW output and the gradient it sends back into the UNet."""
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu


class TinyUNet(torch.nn.Module):
    """UNet2DConditionModel-shaped callable: (sample, timestep, encoder_hidden_states, class_labels, added_cond_kwargs=...)."""

    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(4, 4, 3, padding=1).to(torch.bfloat16)
        self.txt = torch.nn.Linear(64, 4).to(torch.bfloat16)
        self.tid = torch.nn.Linear(6, 4).to(torch.bfloat16)
        self.seen = {}

    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, added_cond_kwargs=None,
                cross_attention_kwargs=None, return_dict=True):
        self.seen = {"timestep": timestep, "class_labels": class_labels, "added": added_cond_kwargs, "enc": encoder_hidden_states}
        cond = self.txt(encoder_hidden_states).mean(1) + self.tid(added_cond_kwargs["time_ids"] / 1024.0)
        t = (timestep.float() / 1000.0).to(sample.dtype).view(-1, 1, 1, 1)
        return (self.conv(sample) * (1 + t) + cond.view(-1, 4, 1, 1),)


@pytest.mark.parametrize("pred_type,gamma", [("epsilon", None), ("epsilon", 5.0), ("v_prediction", 5.0)])
def test_sdxl_step_plumbing(pred_type, gamma):
    from simpletuner_b200.sdxl.model import SDXL, default_config
    from simpletuner_b200.training import noise as N

    dev = torch.device("cuda")
    unet = TinyUNet().to(dev)
    w = SDXL(default_config(prediction_type=pred_type, snr_gamma=gamma), unet=unet, device=dev)
    g = torch.Generator().manual_seed(0)
    examples = [{"intermediary_size": (256, 256), "crop_coordinates": [0, 0], "drop_conditioning": False},
                {"intermediary_size": (300, 256), "crop_coordinates": [0, 22], "drop_conditioning": False}]
    batch = {"latent_batch": torch.randn(2, 4, 32, 32, generator=g).bfloat16(), "prompt_embeds": torch.randn(2, 77, 64, generator=g).bfloat16(),
             "add_text_embeds": torch.randn(2, 32, generator=g).bfloat16(), "examples": examples}
    torch.manual_seed(11); torch.cuda.manual_seed(11)
    pb = w.prepare_batch(dict(batch), {"global_step": 0})
    assert pb["timesteps"].dtype == torch.int64 and pb["timesteps"].shape == (2,)
    # fused add_noise == DDPMScheduler.add_noise in fp32 (common.py:5998-6002), bit for bit
    ref_noisy = N.add_noise(w.noise_schedule, pb["latents"].float().cpu(), pb["input_noise"].float().cpu(), pb["timesteps"].cpu()).bfloat16()
    assert torch.equal(pb["noisy_latents"].cpu(), ref_noisy)
    assert pb["added_cond_kwargs"]["time_ids"].shape == (2, 6) and pb["added_cond_kwargs"]["time_ids"][1].tolist() == [256.0, 300.0, 0.0, 22.0, 256.0, 256.0]
    out = w.model_predict(pb)
    assert set(out) == {"model_prediction", "hidden_states_buffer", "urepa_hidden_states"} and out["model_prediction"].shape == (2, 4, 32, 32)
    assert unet.seen["class_labels"] is pb["add_text_embeds"] or torch.equal(unet.seen["class_labels"], pb["add_text_embeds"])
    assert unet.seen["added"]["text_embeds"].shape == (2, 32)
    loss = w.loss(pb, out)
    loss.backward()
    got = {n: p.grad.float().clone() for n, p in unet.named_parameters()}
    # eager reference of the loss (common.py:6376-6398) and its autograd
    for p in unet.parameters():
        p.grad = None
    pred = w.model_predict(pb)["model_prediction"]
    target = pb["noise"] if pred_type == "epsilon" else N.get_velocity(w.noise_schedule, pb["latents"], pb["noise"], pb["timesteps"])
    l = torch.nn.functional.mse_loss(pred.float(), target.float(), reduction="none")
    if gamma:
        wts = N.min_snr_loss_weights(pb["timesteps"], w.noise_schedule, gamma, pred_type).float().view(-1, 1, 1, 1)
        l = l * wts
    ref = l.mean(dim=(1, 2, 3)).mean()
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 2e-5 * abs(float(ref)), (float(loss), float(ref))
    cos = torch.nn.functional.cosine_similarity
    for n, p in unet.named_parameters():
        assert float(cos(got[n].flatten(), p.grad.float().flatten(), dim=0)) >= 0.9999, n
