"""Per-kernel device time of batched generation (one predict-chunk of 32 sequences)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import music_generator_b200  # noqa
from music_generator_b200.config import ModelConfig
from music_generator_b200.engine import Engine
from music_generator_b200.sampler import generate_events

G = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = 4
e = Engine(ModelConfig(), precision="fp32"); e.init_params(0)
stys = [np.eye(23)[i % 23] for i in range(G)]
u = np.random.RandomState(7).random_sample((steps, G, 48, 2))
generate_events(e, stys, 1, u[:1], stream_mode=1); torch.cuda.synchronize()
import time
t0 = time.time(); generate_events(e, stys, steps, u, stream_mode=1); torch.cuda.synchronize()
dt = time.time() - t0
print(f"G={G}: {G * steps / dt:.1f} timesteps/s  ({1e3 * dt / steps:.2f} ms per timestep of {G} sequences)")
e.profile = []
generate_events(e, stys, steps, u, stream_mode=1)
for k, v in sorted(e.profile_summary().items(), key=lambda kv: -kv[1][1]):
    print(f"   {k:28s} x{v[0]:3d} {v[1] / steps:9.3f} ms/timestep")
