"""Golden fixture for the MVAE mixture-of-experts decoder: the reference's own `MixedDecoder.forward`
(vid2player/motion_vae/model.py:186-252) EXECUTED on CPU in fp32 with parameters drawn from numpy (tests/helpers.py
`mixed_decoder_params`, so the GPU test rebuilds the same parameters instead of storing 5.7 MB of weights).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_nn.py
Output: tests/golden/nn_mixed_decoder.npz (inputs z, c; outputs; gate coefficients; per-layer activations).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refenv  # noqa: E402

_refenv.setup("vid2player")
from motion_vae.model import MixedDecoder  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from helpers import mixed_decoder_params  # noqa: E402


def main():
    torch.set_num_threads(1)
    frame, latent, hidden, E = 288, 32, 256, 6
    # MixedDecoder(frame_size_in, frame_size_out, latent_size, hidden_size, num_condition_frames, num_future_predictions, num_experts)
    dec = MixedDecoder(frame, frame + 2, latent, hidden, 1, 1, E)
    ws, bs, gate = mixed_decoder_params(seed=5, frame=frame, latent=latent, hidden=hidden, experts=E)
    with torch.no_grad():
        for i in range(3):
            getattr(dec, f"w{i}").copy_(torch.from_numpy(ws[i]))
            getattr(dec, f"b{i}").copy_(torch.from_numpy(bs[i]))
        for j, li in enumerate((0, 2, 4)):
            dec.gate[li].weight.copy_(torch.from_numpy(gate[j][0]))
            dec.gate[li].bias.copy_(torch.from_numpy(gate[j][1]))
    rng = np.random.default_rng(6)
    n = 200                                   # not a multiple of the 128-row tile on purpose
    z = np.clip(rng.normal(size=(n, latent)), -5, 5).astype(np.float32)
    c = rng.normal(size=(n, frame)).astype(np.float32)
    with torch.no_grad():
        zt, ct = torch.from_numpy(z), torch.from_numpy(c)
        out = dec(zt, ct)
        coef = F.softmax(dec.gate(torch.cat((zt, ct), dim=1)), dim=1)
    np.savez_compressed(os.path.join(HERE, "nn_mixed_decoder.npz"), z=z, c=c, out=out.numpy(), coef=coef.numpy(), seed=np.array(5))
    print("nn_mixed_decoder.npz", out.shape, float(out.abs().max()), coef.numpy().max(1).mean())


if __name__ == "__main__":
    main()
