"""Two (or more) ranks: the fused peer-memory exchange + Nadam kernel against NCCL all-reduce + Nadam kernel.
  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_check.py [--steps 3] [--time 20]
Every rank trains two engines from the same initial weights on its own synthetic shard, one per exchange path, then
checks (a) the peer path's weights are bit-identical on all ranks, (b) they agree with the NCCL path to the
rounding noise of the gradient kernels' fp32 atomics, (c) the exact kernel: both paths fed the same gradient give
bit-identical weights at world 2 (a+b is order-free), and prints the time per exchange of either path."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import music_generator_b200  # noqa: F401,E402
import dataset  # noqa: E402
from music_generator_b200 import parallel  # noqa: E402
from music_generator_b200.config import ModelConfig  # noqa: E402
from music_generator_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--time", type=int, default=20)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=16)
    args = ap.parse_args()
    rank, world, local = parallel.env_world()
    import faulthandler
    faulthandler.dump_traceback_later(90, exit=True)      # a stuck rank says where, then leaves

    def stage(msg):
        print(f"[rank {rank}] {msg}", file=sys.stderr, flush=True)
    torch.cuda.set_device(local)
    parallel.init_distributed("nccl")
    a, b = Engine(ModelConfig(), precision="bf16"), Engine(ModelConfig(), precision="bf16")
    a.init_params(0)
    b.init_params(0)
    stage("engines ready")
    peer = parallel.PeerNadam(b)
    stage("peer buffers mapped")
    res = {"world": world}

    # (c) same gradient into both exchanges
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    exact = True
    for it in range(3):
        grad = torch.randn(a.flat_size, device="cuda", generator=g)
        a.gflat.copy_(grad)
        b.gflat.copy_(grad)
        parallel.allreduce_flat(a.gflat)
        a.nadam_step(1.0 / world)
        peer.step(b, b._nadam_scalars(), None)
        torch.cuda.synchronize()
        peer.raise_if_timed_out()
        stage(f"exchange {it} done")
        # two ranks: a+b has one rounding whatever the order, so the paths agree bitwise; more ranks: NCCL's
        # reduction order differs from the kernel's rank order, and Nadam (g / sqrt(v)) magnifies the last-bit
        # differences of near-zero gradients, so compare the bulk
        d = (a.flat - b.flat).abs()
        same = torch.equal(a.flat, b.flat) if world <= 2 else (float(d.median()) < 1e-7 and
                                                                float((d > 1e-4).float().mean()) < 2e-3)
        exact = exact and bool(same)
    res["same_gradient_same_weights"] = exact

    stage(f"same-gradient check: {exact}")
    # (a)+(b) whole training steps on per-rank shards
    x, y = dataset.synthetic_all(args.batch, args.seq, seed=1234 + rank)
    dev = [torch.from_numpy(np.ascontiguousarray(t)).cuda() for t in (x[0], x[1], x[2], x[3], y[0])]
    a.init_params(0)
    b.init_params(0)
    for i in range(args.steps):
        a.train_step(*dev, seed=i, allreduce=parallel.allreduce_flat, world=world)
        b.train_step(*dev, seed=i, world=world)
    torch.cuda.synchronize()
    peer.raise_if_timed_out()
    res["max_abs_diff_vs_nccl"] = float((a.flat - b.flat).abs().max())
    res["tracks_nccl"] = bool(torch.allclose(a.flat, b.flat, rtol=1e-3, atol=2e-5))
    gathered = [torch.empty_like(b.flat) for _ in range(world)]
    dist.all_gather(gathered, b.flat.clone())
    res["ranks_bit_identical"] = all(bool(torch.equal(gathered[0], t)) for t in gathered)

    stage("training steps compared")
    # time per exchange (device events, max over ranks)
    def timed(fn):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.time):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return parallel.max_over_ranks(e0.elapsed_time(e1) / args.time * 1e3, "cuda")

    def nccl_path():
        parallel.allreduce_flat(a.gflat)
        a.nadam_step(1.0 / world)

    res["us_nccl_plus_nadam"] = round(timed(nccl_path), 1)
    res["us_peer_fused"] = round(timed(lambda: peer.step(b, b._nadam_scalars(), None)), 1)
    peer.raise_if_timed_out()
    res["ok"] = bool(res["same_gradient_same_weights"] and res["tracks_nccl"] and res["ranks_bit_identical"])
    stage("timed")
    peer.close(b)
    stage("closed")
    oks = [None] * world
    dist.all_gather_object(oks, res["ok"])
    res["ok"] = all(oks)
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
