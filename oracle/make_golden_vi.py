"""Generate ``tests/golden/vi_elbo_d6.npz``: ``VariationalInference.loss`` of the reference (imported unmodified; only the
missing torchdiffeq is restated) WITH the stochastic terms -- ``elbo=True``: one re-parameterised sample for ``z`` and the
Monte-Carlo KL against ``ExponentialPrior`` (``mc_size`` samples, model.py:1186-1214) -- under ``torch.manual_seed(4321)`` on
the CPU, plus the closed-form-KL variant (``prior_log_pdf=None``).  Stored: inputs, state_dicts, posterior parameters, the
sample ``z``, both losses and the encoder / decoder gradients of the Monte-Carlo variant.  Build container only:

    python -m oracle.make_golden_vi
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
SEED, MC = 4321, 7


def main():
    from _util import make_cohort

    M = refload.load("model")
    cpu = torch.device("cpu")
    D, obs, B = 6, 20, 10
    torch.manual_seed(666)
    enc = M.EncoderLSTM(obs + 1, 2 * obs, D, device=cpu, normalize=True)
    dec = M.RocheExpertDecoder(obs, D, 1, 14, 1, roche=True, method="dopri5", device=cpu)
    _, a, x, mask = make_cohort(B, D, obs=obs, seed=78, dose_max=1.0)
    data = {"measurements": x, "actions": a, "masks": mask}
    out = {"x": x.numpy(), "a": a.numpy(), "mask": mask.numpy(), "seed": np.int64(SEED), "mc_size": np.int64(MC)}
    vi = M.VariationalInference(enc, dec, prior_log_pdf=M.ExponentialPrior.log_density, elbo=True, mc_size=MC)
    torch.manual_seed(SEED)
    loss = vi.loss(data)
    loss.backward()
    out.update(loss_mc=np.float64(loss.item()), mu=vi.mu.detach().numpy(), log_var=vi.log_var.detach().numpy(),
               z=vi.z.detach().numpy(), model_name=np.array(vi.model_name))
    for prefix, mod in (("enc", enc), ("dec", dec)):
        for k, v in mod.state_dict().items():
            out["{}__sd__{}".format(prefix, k)] = v.detach().numpy().copy()
        for k, p in mod.named_parameters():
            if p.grad is not None:
                out["{}__grad__{}".format(prefix, k)] = p.grad.detach().numpy().copy()
    vi2 = M.VariationalInference(enc, dec, prior_log_pdf=None, elbo=True)
    torch.manual_seed(SEED)
    out["loss_closed_form"] = np.float64(vi2.loss(data).item())
    np.savez_compressed(os.path.join(OUT, "vi_elbo_d6.npz"), **out)
    print("loss_mc", out["loss_mc"], "loss_closed_form", out["loss_closed_form"], vi.model_name)


if __name__ == "__main__":
    main()
