"""Generate tests/golden/plms_golden.npz by executing the UNMODIFIED reference p_sample_plms loop (TEST INFRASTRUCTURE ONLY).

Run in the build container (needs /root/reference):  python oracle/make_golden_plms.py
usr/diff/shallow_diffusion_tts.py:168-201 (p_sample_plms), :258-264 (the loop).  The reference only runs for B = 1 (it takes
max() of a [B] tensor), so the fixtures are B = 1; the sampler is deterministic after the start noise,
which comes from oracle/synth.py like everything else."""
from __future__ import annotations

import os
import sys
import warnings
from collections import deque

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import synth  # noqa: E402
from make_golden import K_STEP, build_reference_sampler  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
PLMS_CASES = [dict(seed=61, B=1, T=60, interval=5), dict(seed=62, B=1, T=33, interval=10), dict(seed=63, B=1, T=20, interval=1)]


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(os.cpu_count() or 1)
    ns = ref_shim.load()
    net, model = build_reference_sampler(ns)
    out = {}
    with torch.no_grad():
        for i, c in enumerate(PLMS_CASES):
            inp = synth.kernel_inputs(c["seed"], c["B"], c["T"], 1)
            cond = inp["cond"].transpose(1, 2)
            xs = model.norm_spec(inp["fs2_mel"]).transpose(1, 2)[:, None]
            x = model.q_sample(xs, torch.tensor([K_STEP - 1]), noise=inp["start_noise"])
            model.noise_list = deque(maxlen=4)                                   # :259
            for t in reversed(range(0, K_STEP, c["interval"])):                  # :261-264
                x = model.p_sample_plms(x, torch.full((c["B"],), t, dtype=torch.long), c["interval"], cond)
            out[f"mel.{i}"] = model.denorm_spec(x[:, 0].transpose(1, 2)).numpy()
            out[f"x0.{i}"] = x.numpy()
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "plms_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
