"""Per-kernel device time of one train step of the scaled model (BASELINE configs[4]: 512/256 units, T = 512, B = 16)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import music_generator_b200  # noqa
from music_generator_b200.config import ModelConfig
from music_generator_b200.engine import Engine
import dataset

B, T = 16, 512
e = Engine(ModelConfig(time_axis_units=512, note_axis_units=256, seq_len=512), precision="bf16"); e.init_params(0)
x, y = dataset.synthetic_all(B, T)
dev = [torch.tensor(a).cuda() for a in x] + [torch.tensor(y[0]).cuda()]
for i in range(2):
    e.train_step(*dev, seed=i)
torch.cuda.synchronize()
e.profile = []
e.train_step(*dev, seed=9)
agg = e.profile_summary()
tot = sum(v[1] for v in agg.values())
print(f"scaled model, serialised kernel sum {tot:.2f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"   {k:32s} x{v[0]:2d} {v[1]:8.3f} ms")
