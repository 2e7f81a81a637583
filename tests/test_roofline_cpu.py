"""CPU-only: the byte / FLOP model behind bench.py's roofline numbers against SURVEY section 8(d)'s table
(B = 256, S = 512, T = 128, 2-byte elements)."""
import pytest

from anyscale_workshop_nyc_2023_b200 import roofline
from anyscale_workshop_nyc_2023_b200.synth import SPECS

# model: W_step (M elements), cross-KV GB/step, decode GB over 128 steps, encoder + cross-KV projection TFLOP
TABLE = {"flan-t5-small": (38.5, 1.611, 242.4, 6.59), "flan-t5-base": (123.8, 4.832, 729.3, 28.45),
         "flan-t5-large": (391.5, 12.885, 1960.5, 100.6)}


@pytest.mark.parametrize("name", list(TABLE))
def test_model_matches_survey_table(name):
    spec = SPECS[name]
    w_m, kv_gb, dec_gb, enc_tf = TABLE[name]
    assert roofline.step_weight_elements(spec) / 1e6 == pytest.approx(w_m, abs=0.06)
    assert roofline.cross_attention_bytes_per_launch(spec, [512] * 256) * spec.num_decoder_layers / 1e9 == pytest.approx(kv_gb, abs=0.001)
    assert roofline.decode_bytes(spec, 256, 128, seq=512) / 1e9 == pytest.approx(dec_gb, abs=0.06)
    assert roofline.encoder_flops(spec, 256, seq=512) / 1e12 == pytest.approx(enc_tf, rel=2e-3)


def test_extents_and_the_fp16_contract_change_the_model_as_documented():
    spec = SPECS["flan-t5-base"]
    full = roofline.decode_bytes(spec, 4, 10, seq=512)
    ragged = roofline.decode_bytes(spec, 4, 10, extents=[512, 100, 7, 512])
    assert ragged < full
    assert full - ragged == pytest.approx(10 * 2.0 * spec.num_decoder_layers * 2 * spec.inner_dim * (2048 - 1131))
    extra = roofline.decode_bytes(spec, 4, 10, seq=512, fp32_wo=True) - full
    assert extra == pytest.approx(10 * 2.0 * spec.num_decoder_layers * spec.d_model * spec.d_ff)
    assert roofline.encoder_flops(spec, 2, extents=[512, 512]) == pytest.approx(roofline.encoder_flops(spec, 2, seq=512))
