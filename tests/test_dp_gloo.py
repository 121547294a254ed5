"""N>1 host logic on CPU: gloo, world size 2.  The compute inside each rank is the
CPU oracle (tests may use it); what is under test is the sharding, the single
flat-buffer all-reduce, the 1/world scaling and rank-0-only reporting."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import music_generator_b200  # noqa: F401
    from music_generator_b200 import parallel
    from oracle import deepj_oracle as O
    torch.set_num_threads(2)
    r, w, _ = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world)
    cfg = O.Config()
    p = O.init_params(cfg, 0, torch.float64)
    full = O.synthetic_batch(cfg, 4, 2, 1234, torch.float64)
    idx = parallel.shard_indices(4, rank, world)
    shard = [t[idx] for t in full]
    loss, _, grads = O.loss_and_grads(p, cfg, *shard)
    flat = torch.cat([grads[k].reshape(-1) for k in p])
    parallel.allreduce_flat(flat)
    flat /= world
    lmax = parallel.max_over_ranks(float(loss), "cpu")
    # epoch statistics as TrainModel.fit forms them: [sum of per-sequence losses, sequences] summed over ranks
    tot_n = parallel.sum_over_ranks(torch.tensor([float(loss) * len(idx), float(len(idx))], dtype=torch.float64))
    torch.save({"flat": flat, "loss": float(loss), "lmax": lmax, "idx": idx, "epoch_loss": float(tot_n[0] / tot_n[1])},
               os.path.join(out_dir, f"r{rank}.pt"))
    torch.distributed.destroy_process_group()


def test_dp_allreduce_equals_mean_of_shard_gradients(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f"r{r}.pt"), weights_only=False) for r in range(world)]
    assert torch.equal(res[0]["flat"], res[1]["flat"])                      # every rank holds the same average
    assert sorted(np.concatenate([r["idx"] for r in res]).tolist()) == [0, 1, 2, 3]
    assert res[0]["lmax"] == res[1]["lmax"] == max(r["loss"] for r in res)
    # every rank reports the same epoch loss = the mean over all sequences of all ranks (early stopping stays in step)
    assert res[0]["epoch_loss"] == res[1]["epoch_loss"]
    assert abs(res[0]["epoch_loss"] - np.mean([r["loss"] for r in res])) < 1e-12
    # single-process check: average of the two shard gradients (pitch_bins is scoped to the LOCAL batch)
    sys.path.insert(0, ROOT)
    from oracle import deepj_oracle as O
    cfg = O.Config()
    p = O.init_params(cfg, 0, torch.float64)
    full = O.synthetic_batch(cfg, 4, 2, 1234, torch.float64)
    acc = None
    for r in range(world):
        _, _, g = O.loss_and_grads(p, cfg, *[t[r::world] for t in full])
        f = torch.cat([g[k].reshape(-1) for k in p])
        acc = f if acc is None else acc + f
    assert torch.allclose(res[0]["flat"], acc / world, rtol=0, atol=1e-15)


def test_reference_arm_prints_one_line_from_rank0_only():
    port = _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus",
           "2", "--steps", "1", "--warmup", "0"]
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["n_gpus"] == 2


def test_step_exchange_selection_without_gpus(monkeypatch):
    """World 1 needs no exchange; on a host without CUDA a multi-rank job gets the all-reduce (the fused
    peer-memory kernel is a GPU path and is never substituted by host code)."""
    sys.path.insert(0, ROOT)
    import music_generator_b200  # noqa: F401
    from music_generator_b200 import parallel
    assert parallel.make_step_exchange(None, 1) == (None, None)
    if not torch.cuda.is_available():
        assert parallel.make_step_exchange(None, 2) == (parallel.allreduce_flat, None)
    monkeypatch.setenv("DJ_PEER_NADAM", "0")
    assert parallel.make_step_exchange(None, 4) == (parallel.allreduce_flat, None)


def test_training_shards_are_equal_and_disjoint():
    """train.py gives every rank n // world sequences (the remainder is dropped) so that all ranks run the same
    number of steps per epoch."""
    sys.path.insert(0, ROOT)
    import music_generator_b200  # noqa: F401
    from music_generator_b200 import parallel
    for n, world in ((41, 2), (256, 8), (7, 4), (64, 1)):
        shards = [parallel.shard_indices(n, r, world)[:n // world] for r in range(world)]
        assert all(len(s) == n // world for s in shards)
        allidx = np.concatenate(shards)
        assert len(set(allidx.tolist())) == len(allidx) and allidx.max(initial=-1) < n


def test_generation_shards_are_whole_predict_chunks():
    """Generation sharding (no collective): every sequence on exactly one rank, whole chunks of 32 together."""
    from music_generator_b200 import parallel
    for G in (1, 31, 32, 70, 1024):
        for world in (1, 2, 8):
            shards = [parallel.shard_sequence_chunks(G, r, world) for r in range(world)]
            assert sorted(sum(shards, [])) == list(range(G))
            for sh in shards:
                for g in sh:
                    assert all(h in sh for h in range(32 * (g // 32), min(32 * (g // 32) + 32, G)))
    assert len(parallel.shard_sequence_chunks(1024, 3, 8)) == 128
