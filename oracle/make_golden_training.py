"""Generate ``tests/golden/training_iter_d6.npz``: ONE training iteration of the reference, computed by the reference's own
``model.py`` classes (``EncoderLSTM``, ``RocheExpertDecoder``, ``VariationalInference.loss`` with ``elbo=False`` so that no
random sample enters) imported unmodified, with ``oracle.odeint`` standing in for the missing torchdiffeq.  Stored: the
mini-batch, both ``state_dict``s, the loss, the encoder output and every parameter gradient after ``loss.backward()``
(``training_utils.py:41-50``).  Build container only:

    python -m oracle.make_golden_training
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refload  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    from _util import make_cohort

    M = refload.load("model")
    cpu = torch.device("cpu")
    D, obs, B = 6, 20, 12
    torch.manual_seed(666)
    enc = M.EncoderLSTM(obs + 1, 2 * obs, D, device=cpu, normalize=True)
    dec = M.RocheExpertDecoder(obs, D, 1, 14, 1, roche=True, method="dopri5", device=cpu)
    vi = M.VariationalInference(enc, dec, prior_log_pdf=None, elbo=False)
    _, a, x, mask = make_cohort(B, D, obs=obs, seed=77, dose_max=1.0)
    data = {"measurements": x, "actions": a, "masks": mask}
    loss = vi.loss(data)
    loss.backward()
    out = {"x": x.numpy(), "a": a.numpy(), "mask": mask.numpy(), "loss": np.float64(loss.item()), "mu": vi.mu.detach().numpy(),
           "x_hat": vi.x_hat.detach().numpy()}
    for prefix, mod in (("enc", enc), ("dec", dec)):
        for k, v in mod.state_dict().items():
            out["{}__sd__{}".format(prefix, k)] = v.detach().numpy().copy()
        for k, p in mod.named_parameters():
            if p.grad is not None:
                out["{}__grad__{}".format(prefix, k)] = p.grad.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "training_iter_d6.npz"), **out)
    print("loss", loss.item(), "keys", len(out), "grads", sorted(k for k in out if "__grad__" in k))


if __name__ == "__main__":
    main()
