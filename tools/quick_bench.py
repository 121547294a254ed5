"""Developer timing probe (not the contract bench): per-entry-point device time of
one train step and one generation timestep."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import music_generator_b200  # noqa
from music_generator_b200.config import ModelConfig
from music_generator_b200.engine import Engine
from music_generator_b200.sampler import generate_events
import dataset


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    out = {}
    for prec in ("fp32", "bf16"):
        e = Engine(ModelConfig(), precision=prec)
        e.init_params(0)
        x, y = dataset.synthetic_all(B)
        dev = [torch.tensor(a).cuda() for a in x] + [torch.tensor(y[0]).cuda()]
        for i in range(2):
            e.train_step(*dev, seed=i)
        torch.cuda.synchronize()
        t0 = time.time()
        for i in range(3):
            e.train_step(*dev, seed=10 + i)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / 3
        e.profile = []
        e.train_step(*dev, seed=99)
        agg = e.profile_summary()
        e.profile = None
        out[prec] = dict(ms_per_step=dt * 1e3, seqs_per_s=B / dt,
                         kernels={k: [v[0], round(v[1], 3)] for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])})
        print(prec, "B", B, "ms/step", round(dt * 1e3, 2), "seq/s", round(B / dt, 1), flush=True)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"   {k:28s} x{v[0]:3d} {v[1]:9.3f} ms")
        del e
        torch.cuda.empty_cache()
    e = Engine(ModelConfig(), precision="fp32")
    e.init_params(0)
    sty = np.mean([np.eye(23)[i] for i in (0, 5, 12)], axis=0)
    steps = 16
    u = np.random.RandomState(42).random_sample(2 * 48 * steps)
    generate_events(e, [sty], 2, u)
    torch.cuda.synchronize()
    t0 = time.time()
    generate_events(e, [sty], steps, u)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / steps
    print("generation G=1: ms/timestep", round(dt * 1e3, 3), "timesteps/s", round(1 / dt, 1))
    e.profile = []
    generate_events(e, [sty], 4, u)
    agg = e.profile_summary()
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:28s} x{v[0]:3d} {v[1] / 4:9.3f} ms/timestep")
    out["gen_g1_ms_per_timestep"] = dt * 1e3
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/quick_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
