"""Timing probe for the GPU cohort generator (development aid): python scripts/kbench_datagen.py [--patients N] [--D 8]"""
import argparse
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import hybrid_ode_neurips_2021_b200 as H  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--patients", type=int, default=1 << 20)
ap.add_argument("--D", type=int, default=8)
ap.add_argument("--obs", type=int, default=40)
args = ap.parse_args()
np.random.seed(666)
torch.manual_seed(666)
dg = H.DataGeneratorRoche(args.patients, args.obs, 14, 1, H.RochConfig(kel=1), 0.2, 10, args.D, 0.5, p_remove=0.5,
                          output_sparsity=0.625, device=torch.device("cuda:0"), val_size=100, test_size=1000, exact_rng=False)
dg.generate_data()
torch.cuda.synchronize()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); dg.generate_data(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
init = dg.latents[0].contiguous()
dt, da = torch.as_tensor(dg.dose_time), torch.as_tensor(dg.dose_amount)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); dg.solve_latents(init, dt, da); e.record(); torch.cuda.synchronize()
print(json.dumps({"patients": args.patients, "D": args.D, "generate_data_s": min(ts), "solve_ms": s.elapsed_time(e),
                  "patients_per_s": args.patients / min(ts), "reference_lsoda_ms_per_patient": 22.0}))
