"""Generates tests/golden/text_golden.pt by running the REAL `transformers` classes the reference calls
(T5EncoderModel / CLIPTextModel, simpletuner/helpers/models/flux/pipeline.py:1085, 1127-1130) on tiny random-weight configs.
The weights are the oracle's own seeded initialisation, rounded to bf16-representable values (the CUDA path stores bf16
weights; the fp32 reference then sees exactly the same numbers), loaded into the transformers modules (strict), so the fixture
only needs to hold the seed, the token ids and the outputs.  Run here (transformers is importable in this container):
    python -m oracle.make_golden_text
"""
from pathlib import Path

import torch

from oracle import text_oracle as TO

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "text_golden.pt"


def main():
    import transformers
    from transformers import CLIPTextConfig, CLIPTextModel, T5Config, T5EncoderModel

    torch.manual_seed(0)
    fx = {"transformers_version": transformers.__version__}
    # ---- T5 (v1.1 style: gated-gelu, no bias, un-tied), two sequence lengths incl. one beyond max_distance
    tc = TO.T5Config(vocab_size=200, d_model=128, d_kv=64, d_ff=256, num_layers=2, num_heads=4)
    hf = T5Config(vocab_size=tc.vocab_size, d_model=tc.d_model, d_kv=tc.d_kv, d_ff=tc.d_ff, num_layers=tc.num_layers,
                  num_heads=tc.num_heads, relative_attention_num_buckets=32, relative_attention_max_distance=128,
                  feed_forward_proj="gated-gelu", dropout_rate=0.0, layer_norm_epsilon=1e-6, is_encoder_decoder=False,
                  use_cache=False, tie_word_embeddings=False)
    m = T5EncoderModel(hf).eval()
    P = {k: v.bfloat16().float() for k, v in TO.init_params(TO.t5_param_shapes(tc), seed=11).items()}   # bf16-representable
    sd = dict(P)
    sd["encoder.embed_tokens.weight"] = P["shared.weight"]
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("embed_tokens" in k or "shared" in k for k in missing), (missing, unexpected)
    fx["t5_cfg"] = tc.__dict__.copy()
    fx["t5_seed"] = 11
    for S in (40, 300):
        ids = torch.randint(0, tc.vocab_size, (2, S))
        with torch.no_grad():
            out = m(ids, output_hidden_states=False)[0]
        fx[f"t5_ids_{S}"], fx[f"t5_out_{S}"] = ids, out
    # ---- CLIP text model (quick_gelu, causal), both pooling rules
    for eos in (2, 199):
        cc = TO.CLIPTextConfig(vocab_size=200, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                               max_position_embeddings=77, eos_token_id=eos)
        hf = CLIPTextConfig(vocab_size=cc.vocab_size, hidden_size=cc.hidden_size, intermediate_size=cc.intermediate_size,
                            num_hidden_layers=cc.num_hidden_layers, num_attention_heads=cc.num_attention_heads,
                            max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=eos,
                            bos_token_id=0, pad_token_id=1, attention_dropout=0.0)
        m = CLIPTextModel(hf).eval()
        P = {k: v.bfloat16().float() for k, v in TO.init_params(TO.clip_param_shapes(cc), seed=12).items()}
        missing, unexpected = m.load_state_dict(P, strict=False)
        assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
        ids = torch.randint(3, 190, (3, 77))
        ids[:, 0] = 0
        for b, n in enumerate((9, 40, 77)):      # an end-of-text token (highest id) at position n - 1, padding after
            ids[b, n - 1] = 199
            ids[b, n:] = 199 if eos == 2 else 1
        with torch.no_grad():
            o = m(ids, output_hidden_states=False)
        fx[f"clip_cfg_eos{eos}"] = cc.__dict__.copy()
        fx[f"clip_ids_eos{eos}"], fx[f"clip_last_eos{eos}"], fx[f"clip_pooled_eos{eos}"] = ids, o.last_hidden_state, o.pooler_output
    fx["clip_seed"] = 12
    torch.save(fx, OUT)
    print("wrote", OUT, {k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in fx.items()})


if __name__ == "__main__":
    main()
