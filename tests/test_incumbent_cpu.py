"""The GPU incumbent's control flow (oracle/gpu_incumbent.py: flash_attn Block modules inside the oracle's forward) must be the
oracle's own, which tests/test_oracle.py pins to the unmodified reference: with the attention kernel swapped for
flash_attn's pure-torch SelfAttention it runs on CPU in fp32 and has to agree with the oracle to the last bit."""
import sys

import pytest
import torch

pytest.importorskip("flash_attn")


def test_incumbent_equals_oracle_on_cpu():
    sys.modules.setdefault("flash_attn.ops.triton.layer_norm", None)   # driver-less container
    from flash_attn.modules.mha import SelfAttention
    from oracle import gpu_incumbent as G
    from oracle import mae3d_oracle as O
    from oracle.gen_golden import TOY, toy_inputs
    sd, vol, noise = toy_inputs()
    m = G.IncumbentMAE(TOY)
    m.load_state_dict(sd, strict=True)                                 # the reference's parameter names and shapes
    for blk in list(m.blocks) + list(m.decoder_blocks):
        blk.mixer.inner_attn = SelfAttention()
        blk.mixer.use_flash_attn = False
    (loss, fl), pred, mask = m(vol, 0.9, noise, True)
    loss.backward()
    (ref, grads) = O.forward_backward(TOY, sd, vol, 0.9, noise, True)
    assert float(loss) == float(ref[0][0]) and torch.equal(mask, ref[2]) and torch.equal(pred, ref[1])
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(grads)                                      # quirk Q13
    for k in got:
        assert torch.equal(got[k], grads[k]), k
